// fftlog.cuh -- batched discrete log-space Hankel transform (FFTLog) on the device.
//
// Replaces FortranFFTLog + fhti/fht/fhtq + FFTPACK (pyRSD/_gcl/cpp/FortranFFTLog.cpp,
// pyRSD/_gcl/extern/fftlog/fftlog.f:233-784, drfftf/drfftb.f).  One CTA transforms TWO real
// sequences at once (packed as one complex sequence: the operator is real-linear), entirely in
// shared memory: power-law bias -> radix-2 FFT -> multiply by u_m -> inverse FFT -> reversal,
// 1/N and output bias, i.e. the whole of fht() fused in one kernel (N >= 1024: register-tiled Stockham passes of
// radix 16, k_fftlog_r16; smaller powers of two: radix-2 passes).  Lengths that are not a
// power of two use the equivalent real-space circular kernel g (an O(N^2) contraction).
#pragma once
#include "common.h"

#include <memory>
#include <vector>

namespace prsd {

struct FFTLogPlan {
    int N = 0;
    double mu = 0, q = 0, dlnr = 0, kr = 1;     // kr AFTER the optional low-ringing adjustment; q = the exponent of the
                                                // bias factors (-q of the multipliers for a backward plan)
    bool pow2 = false;
    DevBuf<double> ure, uim;     // multipliers, m = 0 .. N/2
    DevBuf<double> g;            // real-space circular kernel g[d], d = 0 .. N-1 (built on demand)
    DevBuf<double> tw;           // exp(-2 pi i j / N), j < N/2, interleaved (re, im): register-tiled kernel
    void build_g(Context& ctx);
};

// cached per (device, N, mu, q, dlnr, kr, kropt, dir); dir = -1: the backward transform (FFTLog.Transform(a, -1))
std::shared_ptr<FFTLogPlan> fftlog_plan(Context& ctx, int N, double mu, double q, double dlnr, double kr, int kropt, int dir = 1);

// plan with caller-supplied multipliers U_m, m = 0 .. N/2 (U_{N/2} real), no bias: out[i] = (1/N) sum_m F_m U_m
// exp(-2 pi i m (i + 1) / N) with F = DFT of the input -- the mcfit-style transforms of transfer.cu; not cached
std::shared_ptr<FFTLogPlan> fftlog_plan_custom(Context& ctx, int N, double dlnr, const std::vector<double>& re,
                                               const std::vector<double>& im);

// optional windowing of a zero-padded sequence (N in [1024, 8192] only): input element i of row b is
// pre[i] * in[b*in_stride + i] placed at slot off_in + i, i < n_io; output i = post[i] * slot (off_out + i)
struct FFTLogWindow {
    const double* pre = nullptr;     // device [n_io]
    const double* post = nullptr;    // device [n_io]
    int n_io = 0, off_in = 0, off_out = 0;
    long long in_stride = 0, out_stride = 0;
};

// forward transform (dir = +1) of `batch` sequences; d_in / d_out are device pointers [batch][N]
// (may alias).
void fftlog_apply(Context& ctx, FFTLogPlan& plan, int batch, const double* d_in, double* d_out,
                  const FFTLogWindow* win = nullptr);

} // namespace prsd
