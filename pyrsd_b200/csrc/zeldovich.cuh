// zeldovich.cuh -- Zel'dovich P00 / P01 / P11 for batches of cosmologies.
//
// Reference: ZeldovichPS::fftlog_compute (pyRSD/_gcl/cpp/ZeldovichPS.cpp:112-151): for every k
// and every Bessel order n = 0..15 it builds a_n(r; k) on the 1024-point r grid, constructs a
// FortranFFTLog (complex log-Gamma set-up included), transforms, and then keeps ONE linearly
// interpolated sample of the 1024 outputs.  Because the discrete transform is a circular
// convolution with a kernel g_n that depends only on (N, dlnr, mu = n + 1/2), that one sample is
//
//      out_n(k) = sum_l a_n(r_l; k) * [ w_a g_n(N-1-j_a-l) + w_b g_n(N-1-j_b-l) ]
//
// with (j_a, j_b, w_a, w_b) the reference's nearest_interp_1d stencil (ZeldovichPS.cpp:58-82).
// g_n, the stencils and the low-ringing kr are precomputed once per k list; a cosmology then
// costs 16 x 1024 fused multiply-adds per (k, spectrum) with the three spectra sharing every
// exponential -- no FFT at all on the per-cosmology path.
#pragma once
#include "common.h"
#include "fftlog.cuh"

namespace prsd {

static const int ZEL_N = 1024;      // NUM_PTS
static const int ZEL_NMAX = 15;     // orders 0..15

struct ZelPlan {
    int nk = 0;
    std::vector<double> k;
    DevBuf<double> dk;
    DevBuf<double> g;           // [16][1024]
    DevBuf<double> G;           // [nk][16][1024] g folded with the per-k output stencil
    DevBuf<int> ja, jb;         // [nk][16]
    DevBuf<double> wa, wb;      // [nk][16]
    DevBuf<double> r;           // [1024]
    void build(Context& ctx, const double* k, int nk);
};

// X0, XX(=X0+2 sigma^2 is formed inside), YY : device [ncosmo][1024] at the normalisation of the
// spectra they were computed from; `ratio[c]` (device, may be NULL) rescales them by
// (sigma8_new/sigma8_old)^2 as ZeldovichPS::SetSigma8AtZ does (ZeldovichPS.cpp:95-109).
// lowk: use the SPT low-k forms below k0_low (ZeldovichPS.cpp:205-208, 229-232, 262-265); then
// plin [ncosmo][nk] (linear power at the CURRENT normalisation) and q3coef [ncosmo]
// (= Q3_Zel(k)/k^4 at the normalisation of X0) must be given.
// out: device [ncosmo][3][nk] = P00, P01, P11 (Zel'dovich parts only).
void zeldovich_eval(Context& ctx, const ZelPlan& plan, int ncosmo, const double* X0, const double* YY,
                    const double* sigmasq, int sigmasq_stride, const double* ratio, bool lowk, double k0_low,
                    const double* plin, const double* q3coef, int q3_stride, double* out);

// same with an explicit XX array (the 8-argument ZeldovichPS state constructor,
// ZeldovichPS.cpp:32-39, takes X0 and XX separately); XX == NULL means X0 + 2 sigma^2
void zeldovich_eval_xx(Context& ctx, const ZelPlan& plan, int ncosmo, const double* X0, const double* XX, const double* YY,
                       const double* sigmasq, int sigmasq_stride, const double* ratio, bool lowk, double k0_low,
                       const double* plin, const double* q3coef, int q3_stride, double* out);

// general form: output row w reads the inputs of row cmap[w] (cmap == NULL: identity); ratio, plin
// and out are indexed by w
void zeldovich_eval_map(Context& ctx, const ZelPlan& plan, int nrow, const int* cmap, const double* X0, const double* XX,
                        const double* YY, const double* sigmasq, int sigmasq_stride, const double* ratio, bool lowk,
                        double k0_low, const double* plin, const double* q3coef, int q3_stride, double* out, bool only_p00 = false);

} // namespace prsd
