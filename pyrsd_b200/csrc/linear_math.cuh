// linear_math.cuh -- host/device arithmetic of the 1-D product-integration weights
// (shared by linear.cu and the CPU-only test emulation).
#pragma once
#include "pt_math.cuh"
#include "linear.cuh"

namespace prsd {

static const int GLN = 16;

struct ZonesDev {
    int nz;
    double x0[8], h[8];
    int n[8], start[8];
};

struct LinOutDev { int type, sub; double param, lo, hi, scale; };

// kernel per dq (q = integration variable), without the overall scale
PRSD_HD double lin_kernel(int type, int sub, double p, double q)
{
    const double inv2pi2 = 1.0/(2.0*M_PI*M_PI);
    switch (type) {
        case LK_JMN:  return inv2pi2*jmn_kernel(sub, q/p);
        case LK_SIGV: return 1.0/(6.0*M_PI*M_PI);
        case LK_XZEL: { const double x = q*p; return inv2pi2*(-2.0*sph_j1(x)/x); }
        case LK_YZEL: { const double x = q*p; return inv2pi2*(-2.0*sph_j0(x) + 6.0*sph_j1(x)/x); }
        case LK_INVQ2: return 1.0/(q*q);
        case LK_SIGMA_R: {
            // top-hat window of PowerSpectrum.cpp:28-33 (including its small-x branch)
            const double x = q*p;
            const double W = (x < 1e-5) ? 1.0 - (1.0/30.0)*x*x : 3.0/(x*x*x)*(sin(x) - x*cos(x));
            return inv2pi2*q*q*W*W;
        }
        case LK_SIGMA3: {
            // I_R of PowerSpectrum.cpp:139-149
            double r = q/p;
            if (r == 1.0) r = 1.0 - 1e-10;
            const double r2 = r*r, r4 = r2*r2, w = r2 - 1.0, w2 = w*w;
            const double lg = log(((r - 1.0)*(r - 1.0))/((r + 1.0)*(r + 1.0)));
            const double IR = (5.0/(512.0*r4*r))*(-4.0*r*(1.0 + r2)*(3.0 - 14.0*r2 + 3.0*r4) - 3.0*w2*w2*lg);
            return inv2pi2*q*q*IR;
        }
        case LK_XI: return inv2pi2*q*q*sph_j0(q*p);      // = q sin(q r) / (2 pi^2 r), CorrelationFunction.cpp:151-157
        default: return 0.0;
    }
}

// angular frequency of the kernel in q (for sub-panelling)
PRSD_HD double lin_omega(int type, double p)
{
    if (type == LK_XZEL || type == LK_YZEL || type == LK_XI) return p;
    if (type == LK_SIGMA_R) return 2.0*p;
    return 0.0;
}

// Omega[o][j]: integral of the j-th cardinal function of the 4-point Lagrange read against
// the kernel of output S, over the cells whose stencil contains knot j.
PRSD_HD double linear_weight(const LinOutDev& S, const ZonesDev& Z, int j, const double* glx, const double* glw)
{
    double total = 0.0;
    int z = 0;
    while (z < Z.nz - 1 && j >= Z.start[z + 1]) ++z;
    const int jl = j - Z.start[z];
    const int n = Z.n[z];
    const double x0 = Z.x0[z], h = Z.h[z];
    const double xlo = log(S.lo), xhi = log(S.hi);
    const bool has_break = (S.type == LK_JMN || S.type == LK_SIGMA3);
    const double omega = lin_omega(S.type, S.param);
    const int c_lo = jl - 3 > 0 ? jl - 3 : 0;
    const int c_hi = jl + 3 < n - 1 ? jl + 3 : n - 1;
    for (int c = c_lo; c <= c_hi; c++) {
        int s0 = c - 1;
        if (s0 < 0) s0 = 0;
        if (s0 > n - 3) s0 = n - 3;
        const int l = jl - s0;
        if (l < 0 || l > 3) continue;
        double xa = x0 + h*c, xb = x0 + h*(c + 1);
        if (xa < xlo) xa = xlo;
        if (xb > xhi) xb = xhi;
        if (!(xb > xa)) continue;
        const double qa = exp(xa), qb = exp(xb);
        // pieces: optionally split at the kernel's kink q = param
        double pe[3];
        int np = 1;
        pe[0] = qa;
        if (has_break && S.param > qa && S.param < qb) { pe[1] = S.param; np = 2; }
        pe[np] = qb;
        for (int ip = 0; ip < np; ip++) {
            const double p0 = pe[ip], p1 = pe[ip + 1];
            long long nsub = 1;
            if (omega > 0.0) {
                const double ph = omega*(p1 - p0);
                if (ph > 10.0) nsub = (long long)ceil(ph/10.0);
            }
            const double dq = (p1 - p0)/(double)nsub;
            double acc = 0.0;
            for (long long is = 0; is < nsub; is++) {
                const double mid = p0 + dq*((double)is + 0.5), hw = 0.5*dq;
                double part = 0.0;
                for (int g = 0; g < GLN; g++) {
                    const double q = mid + hw*glx[g];
                    const double s = (log(q) - x0)/h - (double)s0;
                    double L[4];
                    lagrange4(s, L);
                    part += glw[g]*L[l]*lin_kernel(S.type, S.sub, S.param, q);
                }
                acc += part*hw;
            }
            total += acc;
        }
    }
    return total*S.scale;
}

} // namespace prsd
