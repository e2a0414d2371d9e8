// discrete.cuh -- see discrete.cu
#pragma once
#include "common.h"

namespace prsd {

// method: 1 SimpsIntegrate, 2 TrapzIntegrate (IntegrationMethods of the reference); y [batch][n] on shared x [n]
void discrete_integrate_dev(Context& ctx, int method, int batch, int n, const double* d_x, const double* d_y, double* d_out);

// ComputeXiLM with the SIMPS / TRAPZ methods: out [batch][nr] = scale * \int dk/(2 pi^2) k^m j_l(k r) P(k) e^{-(k R)^2}
void xilm_discrete_dev(Context& ctx, int method, int batch, int l, int m, int nk, const double* d_k, const double* d_pk,
                       int nr, const double* d_r, double smoothing, double scale, double* d_out);

// LinearSpline: y [batch][n] on shared abscissae x [n], queries xq [nq] -> out [batch][nq]; deriv != 0: the slope
void linear_spline_dev(Context& ctx, int batch, int n, const double* d_x, const double* d_y, int nq, const double* d_xq, int deriv,
                       double* d_out);

} // namespace prsd
