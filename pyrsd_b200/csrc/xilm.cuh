// xilm.cuh -- see xilm.cu
#pragma once
#include "common.h"
#include "fftlog.cuh"

namespace prsd {

// ComputeXiLM_fftlog: log-spaced input (d_k device + h_k host copy, d_pk [batch][N]); fills r_host and
// d_xi [batch][N]
void compute_xilm_fftlog_dev(Context& ctx, int batch, int l, int m, int N, const double* d_k, const double* h_k,
                             const double* d_pk, double q, double smoothing, std::vector<double>& r_host,
                             double* d_xi, double* kr_out);

// ComputeXiLM (FFTLOG method): arbitrary increasing h_k; output scale*xi_l^m at h_r, d_out [batch][nr]
void compute_xilm_dev(Context& ctx, int batch, int l, int m, int nk, const double* h_k, const double* d_pk,
                      int nr, const double* h_r, double smoothing, double scale, double* d_out);

// netlib cubic spline of `batch` rows y [batch][n] on shared abscissae h_x, evaluated at h_xq -> d_out [batch][nq]
void spline_eval_host_x(Context& ctx, int batch, int n, const double* h_x, const double* d_y, int nq, const double* h_xq, double* d_out);

} // namespace prsd
