// transfer.cuh -- theory-to-data transfer functions on the device.
//
//   * GridOp: the (k, mu)-grid reductions of rsd/transfers/grid.py -- GriddedWedgeTransfer.average (:165-187) and
//     GriddedMultipoleTransfer.__call__ (:300-319) are both "sum over mu of a fixed weight x P(k, mu)" once the
//     mode counts, the mu-bin membership and the (2 ell + 1) L_ell(mu) factors are folded into one weight array;
//   * WindowOp: WindowFunctionTransfer.__call__ (rsd/transfers/window.py:51-132): the multipoles, zero-padded in k,
//     go to configuration space with the vendored mcfit P2xi (extern/mcfit/mcfit.py:86-176, cosmology.py:15-23),
//     are mixed with the window-function kernels (rsd/window.py:165-213) and come back with xi2P
//     (cosmology.py:26-35).  mcfit's transform is rfft -> multiply by u_m -> hfft; the register-tiled FFTLog
//     kernel (fftlog.cu) does the same with the multipliers U_m = u_m exp(2 pi i m / N), its window mode supplying
//     the zero padding, the power-law prefactors and the cropping.
#pragma once
#include "common.h"
#include "fftlog.cuh"

#include <memory>
#include <vector>

namespace prsd {

struct GridOp {
    Context* ctx;
    int nbin, nk, nmu;
    DevBuf<double> w;                 // [nbin][nk][nmu]
    GridOp(Context& c, int nbin_, int nk_, int nmu_, const double* h_w);
    // d_power [nw][nk][nmu] -> d_out [nw][nbin][nk]   (device pointers)
    void apply(int nw, const double* d_power, double* d_out);
};

struct WindowOp {
    Context* ctx;
    int nk, nell, N, off_in, off_out;
    double q, delta;
    std::vector<int> ells;
    std::vector<double> k, rr, lnxy;                   // rr: the separation grid of the LAST ell (window.py:100-102)
    std::vector<std::shared_ptr<FFTLogPlan>> plan;     // one per ell; P2xi and xi2P share the multipliers
    DevBuf<double> pre1, post1, pre2, post2;           // [nell][nk], [nell][nk], [nk], [nell][nk]
    DevBuf<double> kern;                               // [nell][nell][nk]
    bool has_kernel = false;
    DevBuf<double> xi, xic;                            // scratch [nw][nell][nk]
    int last_nw = 0;

    WindowOp(Context& c, int nk_, const double* k_, int nell_, const int* ells_, double q_, int N_);
    void set_kernel(const double* h_kern);
    // mode 0: convolution; 1: dry run (P -> xi -> P without the window).  d_in, d_out [nw][nell][nk] device
    void apply(int nw, const double* d_in, double* d_out, int mode);
};

} // namespace prsd
