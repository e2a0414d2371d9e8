// linear.cuh -- every 1-D integral of the hot path as a precomputed linear functional of a
// tabulated spectrum ("product integration"):
//
//     O[o] = \int_{lo}^{hi} S(q) K_o(q) dq  =  sum_j  S(q_j) * Omega[o][j]
//
// where S is read with the same 4-point Lagrange interpolation the 2-D operators use and
// Omega[o][j] = \int phi_j(q) K_o(q) dq is integrated ONCE on the device with enough
// Gauss-Legendre sub-panels to resolve the kernel's own oscillations (j0/j1(q r) for the
// Zel'dovich X(r), Y(r); the top-hat window for sigma(R)).  Covers
//   Jmn             pyRSD/_gcl/cpp/Jmn.cpp:96-110            (limits 1e-5 .. 1e5, log kink at q = k)
//   sigma_v^2       PowerSpectrum.cpp:53-62                  (1e-5 .. 100  or  0.5 k)
//   X_Zel, Y_Zel    PowerSpectrum.cpp:90-127                 (1e-5 .. 100)
//   Q3_Zel          PowerSpectrum.cpp:129-136                (on the P_L^2 table)
//   VelocityKurtosis OneLoopPS.cpp:183-193                   (on the P22bar table, 1e-5 .. 10)
//   Sigma(R)^2      PowerSpectrum.cpp:28-41 ; sigma3_squared PowerSpectrum.cpp:146-161
#pragma once
#include "common.h"
#include "bilinear.cuh"

namespace prsd {

enum LinKernel {
    LK_JMN = 0,      // sub = 3 m + n, param = k
    LK_SIGV,         // plain \int S dq / (6 pi^2)
    LK_XZEL,         // param = r
    LK_YZEL,         // param = r
    LK_INVQ2,        // \int S / q^2 dq  (Q3 and the velocity kurtosis; scale carries the prefactor)
    LK_SIGMA_R,      // param = R : \int q^2 S W(qR)^2 dq / (2 pi^2)
    LK_SIGMA3,       // param = k : \int q^2 S I_R(q/k) dq / (2 pi^2)
    LK_XI,           // param = r : \int q^2 S j0(q r) dq / (2 pi^2)  (CorrelationFunction::Evaluate)
};

struct LinOutSpec {
    int type;
    int sub;
    double param;
    double lo, hi;     // integration limits in q
    double scale;      // overall prefactor multiplying the kernel
};

struct LinearOp {
    int nout = 0;
    DevBuf<double> Omega;      // [nout][TAB_MPAD]
    void build(Context& ctx, const std::vector<LinOutSpec>& outs, bool sync = true);
    // out[row][nout] = sum_j tables[tab_index[row]][j] * Omega[o][j]   (all device pointers)
    void apply(Context& ctx, const double* tables, const int* tab_index, int nrow, double* out) const;
};

} // namespace prsd
