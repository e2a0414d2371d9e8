// gp.cuh -- Gaussian-process mean prediction (pyRSD/rsd/simulation.py:16-256), see gp.cu
#pragma once
#include "common.h"
#include <cmath>

namespace prsd {

struct GPModel {
    Context* ctx;
    int n, nd;
    double amp, y_mean, y_scale;
    double mean[8], inv[8];
    DevBuf<double> Z, alpha;
    // x [n][nd] raw training inputs; x_mean / x_scale: StandardScaler of the inputs; metric [nd]: ExpSquaredKernel metric
    // (in scaled units); alpha [n]: (K + diag(yerr^2))^-1 y_scaled
    GPModel(Context& c, int n, int nd, const double* x, const double* x_mean, const double* x_scale, const double* metric,
            double amp, const double* alpha, double y_mean, double y_scale);
    void predict(int npred, const double* x /*host [npred][nd]*/, double* out /*host [npred]*/);
    void predict_device(int npred, const double* d_x, double* d_out);      // no synchronisation
};

} // namespace prsd
