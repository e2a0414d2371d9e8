// pt_math.cuh -- host/device math shared by every kernel of the PT hot path.
//
//  * the mode-coupling kernels of the reference's Imn / Kmn / ImnOneLoop drivers
//    (pyRSD/_gcl/cpp/Imn.cpp:25-111, Kmn.cpp:24-107, ImnOneLoop.cpp:36-72),
//    written in terms of S = u^2, T = v^2, P = u v so that the common
//    denominators are formed once;
//  * the propagator kernels g_mn(r) of Jmn (pyRSD/_gcl/cpp/Jmn.cpp:26-94);
//  * the linear power spectrum of LinearPS::Evaluate (LinearPS.cpp:11-15) on top of
//    Cosmology::GetSplineTransfer / GetNoWiggleTransfer (Cosmology.cpp:187-246) and
//    the netlib cubic spline evaluation (Spline.cpp:11-41, 268-273);
//  * 4-point Lagrange weights used to read the tabulated spectra.
//
// Everything here is `__host__ __device__` so the CPU-only test-suite can exercise
// exactly the arithmetic the CUDA kernels run (tests/host_emul), but nothing in the
// shipped library calls these on the host for results.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PRSD_HD __host__ __device__ __forceinline__
#else
#define PRSD_HD inline
#endif

namespace prsd {

// ---------------------------------------------------------------- kernel ids
enum KernelId {
    // I_mn : id = 4 m + n
    KID_F00 = 0, KID_F01, KID_F02, KID_F03,
    KID_F10, KID_F11, KID_F12, KID_F13,
    KID_F20, KID_F21, KID_F22, KID_F23,
    KID_F30, KID_F31, KID_F32, KID_F33,
    // K_mn
    KID_K00 = 16, KID_K01, KID_K00S, KID_K01S, KID_K02S,
    KID_K10, KID_K11, KID_K10S, KID_K11S,
    KID_K20A, KID_K20B, KID_K20SA, KID_K20SB,
    // ImnOneLoop-only kernels (h03/h04 are K20A/K20B, f23/f32/f33 are the I ones)
    KID_H01 = 29, KID_H02,
    KID_COUNT
};

// Kernels flagged non-symmetric under v -> -v by the reference
// (Imn.cpp:159-180 pass symmetric=false; Kmn.cpp:164-169).
PRSD_HD bool kernel_is_symmetric(int id)
{
    switch (id) {
        case KID_F03: case KID_F10: case KID_F13: case KID_F22: case KID_F30: case KID_F31:
        case KID_K11: case KID_K11S:
            return false;
        default:
            return true;
    }
}

// f(u, v) for any kernel id.  um = u - v and up = u + v are passed separately because the
// quadrature knows them exactly (um = 2 q / k, up = 2 r / k): the singular denominators
// (u - v)^4 (u + v)^2 and (u^2 - v^2)^n are then free of cancellation.
PRSD_HD double pt_kernel(int id, double u, double v, double um, double up)
{
    const double S = u*u, T = v*v, P = u*v;
    const double d  = um*up;            // u^2 - v^2
    const double i1 = 1.0/d;
    const double i2 = i1*i1;            // 1/(u^2-v^2)^2
    const double i4 = i2*i2;            // 1/(u^2-v^2)^4
    const double ium2 = 1.0/(um*um);
    const double ia = ium2*ium2/(up*up);   // 1/((u-v)^4 (u+v)^2)
    // building blocks
    const double A  = 4.0 + 3.0*T + S*(3.0 - 10.0*T);         // 7/2 (u^2-v^2)^2 F2 numerator / 2
    const double B  = -8.0 + T + S*(1.0 + 6.0*T);              // appears in f01, f21, f22
    const double Cc = -8.0 + S + T + 6.0*S*T;                  // G2-like numerator
    const double sm = (S - 1.0)*(T - 1.0);                     // (u^2-1)(v^2-1)
    const double s2 = (2.0*(6.0 + S*S - 6.0*T + T*T + S*(4.0*T - 6.0)))*i2*(1.0/3.0);  // S2
    switch (id) {
        case KID_F00: return 4.0*A*A*i4*(1.0/49.0);
        case KID_F01: return 4.0*B*(-A)*i4*(1.0/49.0);
        case KID_F02: return -8.0*sm*(-A)*i4*(1.0/7.0);
        case KID_F03: return 8.0*(S - 1.0)*(3.0*P - 1.0)*(T - 1.0)*ia;
        case KID_F10: return 4.0*(P - 1.0)*(-A)*ia*(1.0/7.0);
        case KID_F11: return 4.0*Cc*Cc*i4*(1.0/49.0);
        case KID_F12: return -8.0*sm*Cc*i4*(1.0/7.0);
        case KID_F13: return 8.0*(T + S*(1.0 - 2.0*T) + S*P*(3.0*T - 2.0) + P*(1.0 - 2.0*T))*ia;
        case KID_F20: return 8.0*(-1.0 - T + S*(3.0*T - 1.0))*(-A)*i4*(1.0/7.0);
        case KID_F21: return 8.0*(-1.0 - T + S*(3.0*T - 1.0))*B*i4*(1.0/7.0);
        case KID_F22: return 4.0*(P - 1.0)*B*ia*(1.0/7.0);
        case KID_F23: return 48.0*sm*sm*i4;
        case KID_F30: return -8.0*(1.0 + T + S*(1.0 - 3.0*T) + S*P*(5.0*T - 3.0) + P*(1.0 - 3.0*T))*ia;
        case KID_F31: return -8.0*P*sm*ia;
        case KID_F32: return -16.0*sm*(-1.0 - 3.0*T + 3.0*S*(5.0*T - 1.0))*i4;
        case KID_F33: return 16.0*(3.0 + 2.0*T + 3.0*T*T + S*(2.0 + 12.0*T - 30.0*T*T)
                                   + S*S*(3.0 - 30.0*T + 35.0*T*T))*i4;
        case KID_K00:  return 2.0*A*i2*(1.0/7.0);                 // F2
        case KID_K01:  return 1.0;
        case KID_K00S: return 2.0*A*i2*(1.0/7.0)*s2;
        case KID_K01S: return s2*s2;
        case KID_K02S: return s2;
        case KID_K10:  return -2.0*Cc*i2*(1.0/7.0);               // G2
        case KID_K11:  return (2.0 - 2.0*P)*ium2;
        case KID_K10S: return -2.0*Cc*i2*(1.0/7.0)*s2;
        case KID_K11S: return (2.0 - 2.0*P)*ium2*s2;
        case KID_K20A: return 2.0*sm*i2;                          // h03
        case KID_K20B: return 2.0*(1.0 + T + S*(1.0 - 3.0*T))*i2; // h04
        case KID_K20SA: return 2.0*sm*i2*s2;
        case KID_K20SB: return 2.0*(1.0 + T + S*(1.0 - 3.0*T))*i2*s2;
        case KID_H01: return -2.0*sm*ium2*ium2;
        case KID_H02: return (6.0 - 8.0*P - 2.0*T + S*(6.0*T - 2.0))*ium2*ium2;
        default: return 0.0;
    }
}

// ------------------------------------------------------------ J_mn kernels
// g_mn(r), r = q/k.  id = 3 m + n as in Jmn::Evaluate (Jmn.cpp:122-139): valid 0,1,2,3,4,6.
// The reference switches to a large-r series at r >= 60 and nudges r == 1 to 1 - 1e-10.
PRSD_HD double jmn_kernel(int id, double r)
{
    if (r == 1.0) r = 1.0 - 1e-10;
    const double r2 = r*r, r4 = r2*r2, r6 = r4*r2, r8 = r4*r4;
    if (r < 60.0) {
        const double L = log((r + 1.0)/fabs(r - 1.0));
        const double w = r2 - 1.0, w3 = w*w*w, ir2 = 1.0/r2, ir3 = 1.0/(r2*r);
        switch (id) {
            case 0: return (1.0/3024.0)*(12.0*ir2 - 158.0 + 100.0*r2 - 42.0*r4 + 3.0*ir3*w3*(7.0*r2 + 2.0)*L);
            case 1: return (1.0/3024.0)*(24.0*ir2 - 202.0 + 56.0*r2 - 30.0*r4 + 3.0*ir3*w3*(5.0*r2 + 4.0)*L);
            case 2: return (1.0/224.0)*(2.0*ir2*(r2 + 1.0)*(3.0*r4 - 14.0*r2 + 3.0) - 3.0*ir3*w3*w*L);
            case 3: return (1.0/1008.0)*(-38.0 + 48.0*r2 - 18.0*r4 + 9.0/r*w3*L);
            case 4: return (1.0/1008.0)*(12.0*ir2 - 82.0 + 4.0*r2 - 6.0*r4 + 3.0*ir3*w3*(r2 + 2.0)*L);
            case 6: return (1.0/672.0)*(2.0*ir2*(9.0 - 109.0*r2 + 63.0*r4 - 27.0*r6) + 9.0*ir3*w3*(3.0*r2 + 1.0)*L);
            default: return 0.0;
        }
    } else {
        const double r10 = r8*r2, r12 = r6*r6;
        switch (id) {
            case 0: return (-2.0/3024.0)*(70.0 + 125.0*r2 - 354.0*r4 + 263.0*r6 + 400.0*r8 - 1008.0*r10 + 5124.0*r12)/(105.0*r12);
            case 1: return (-2.0/3024.0)*(140.0 - 65.0*r2 - 168.0*r4 + 229.0*r6 + 656.0*r8 - 3312.0*r10 + 10500.0*r12)/(105.0*r12);
            case 2: return (-2.0/224.0)*(35.0 - 95.0*r2 + 93.0*r4 - 17.0*r6 + 128.0*r8 - 1152.0*r10 + 2688.0*r12)/(105.0*r12);
            case 3: return (8.0/1008.0)*(-28.0 - 60.0*r2 - 156.0*r4 - 572.0*r6 - 5148.0*r8 + 1001.0*r10)/(5005.0*r10);
            case 4: return (-2.0/1008.0)*(70.0 - 85.0*r2 + 6.0*r4 + 65.0*r6 + 304.0*r8 - 1872.0*r10 + 5292.0*r12)/(105.0*r12);
            case 6: return (-2.0/672.0)*(35.0 + 45.0*r2 - 147.0*r4 + 115.0*r6 + 192.0*r8 - 576.0*r10 + 2576.0*r12)/(35.0*r12);
            default: return 0.0;
        }
    }
}

// spherical Bessel j0, j1 with the reference's small-argument series
// (SpecialFunctions.cpp:12-24)
PRSD_HD double sph_j0(double x)
{
    if (fabs(x) < 1e-4) { double x2 = x*x; return 1.0 - x2/6.0 + x2*x2/120.0 - x2*x2*x2/5040.0; }
    return sin(x)/x;
}
PRSD_HD double sph_j1(double x)
{
    if (fabs(x) < 1e-4) { double x2 = x*x; return x*(1.0/3.0 - x2/30.0 + x2*x2/840.0 - x2*x2*x2/45360.0); }
    return (sin(x) - x*cos(x))/(x*x);
}

// ------------------------------------------------- linear power spectrum
// Per-cosmology description of P_L(k) uploaded by the host (CLASS stays on the host).
struct PlinParams {
    double ns;          // spectral index
    double amp;         // delta_H^2 * (sigma8_z / sigma8)^2, z = 0 normalisation set by the host
    double h;           // EH no-wiggle parameters (Cosmology.cpp:233-246, 360-362)
    double omh2;
    double keq;         // k_equality
    double s_nw;        // sound-horizon fit
    double alpha_gamma;
    double k0, k1;      // first / last tabulated k
    double T0_ratio;    // T0 / T_nw(k0)
    double T1_ratio;    // T1 / T_nw(k1)
    double lnk0;        // ln k0
    double inv_dlnk;    // (n-1)/ln(k1/k0): knots are (near-)uniform in ln k
    int    n;           // number of spline knots
    int    pad_;
};

PRSD_HD double eh_nowiggle(const PlinParams& p, double k)
{
    k *= p.h;
    const double ks = k*p.s_nw/p.h;
    const double q = k/(13.41*p.keq);
    const double x = 0.43*ks, x2 = x*x;
    const double geff = p.omh2*(p.alpha_gamma + (1.0 - p.alpha_gamma)/(1.0 + x2*x2));
    const double qeff = q*p.omh2/geff;
    const double L0 = log(2.0*M_E + 1.8*qeff);
    const double C0 = 14.2 + 731.0/(1.0 + 62.5*qeff);
    return L0/(L0 + C0*qeff*qeff);
}

// spl = [X | Y | b | c | d], each of length n (netlib cubic spline of T(k) in k)
PRSD_HD double transfer_eval(const PlinParams& p, const double* spl, double k)
{
    if (k <= 0.0) return 0.0;
    if (k <= p.k0) return p.T0_ratio*eh_nowiggle(p, k);
    if (k >= p.k1) return p.T1_ratio*eh_nowiggle(p, k);
    const int n = p.n;
    int i = (int)((log(k) - p.lnk0)*p.inv_dlnk);
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i > 0 && k < spl[i]) --i;
    while (i < n - 2 && k > spl[i + 1]) ++i;
    const double du = k - spl[i];
    return spl[n + i] + du*(spl[2*n + i] + du*(spl[3*n + i] + du*spl[4*n + i]));
}

PRSD_HD double plin_eval(const PlinParams& p, const double* spl, double k)
{
    const double T = transfer_eval(p, spl, k);
    return p.amp*pow(k, p.ns)*T*T;
}

// ----------------------------------------------------------- cubic spline
// Coefficients of the reference's cubic spline (netlib fmm/spline.f end conditions: third
// derivative at both ends from divided differences; Spline.cpp:211-257).
//   y(x) = Y[i] + b[i] t + c[i] t^2 + d[i] t^3,  t = x - X[i],  X[i] <= x <= X[i+1]
// b, c, d are caller-provided work/output arrays of length n (n >= 2).
PRSD_HD void cubic_spline_build(int n, const double* X, const double* Y, double* b, double* c, double* d)
{
    if (n == 2) {
        b[0] = b[1] = (Y[1] - Y[0])/(X[1] - X[0]);
        c[0] = c[1] = 0.0;
        d[0] = d[1] = 0.0;
        return;
    }
    // tridiagonal system: b = diagonal, d = off-diagonal, c = right-hand side
    d[0] = X[1] - X[0];
    c[1] = (Y[1] - Y[0])/d[0];
    for (int i = 1; i < n - 1; i++) {
        d[i] = X[i + 1] - X[i];
        b[i] = 2.0*(d[i - 1] + d[i]);
        c[i + 1] = (Y[i + 1] - Y[i])/d[i];
        c[i] = c[i + 1] - c[i];
    }
    b[0] = -d[0];
    b[n - 1] = -d[n - 2];
    c[0] = 0.0;
    c[n - 1] = 0.0;
    if (n > 3) {
        c[0] = c[2]/(X[3] - X[1]) - c[1]/(X[2] - X[0]);
        c[n - 1] = c[n - 2]/(X[n - 1] - X[n - 3]) - c[n - 3]/(X[n - 2] - X[n - 4]);
        c[0] = c[0]*d[0]*d[0]/(X[3] - X[0]);
        c[n - 1] = -c[n - 1]*d[n - 2]*d[n - 2]/(X[n - 1] - X[n - 4]);
    }
    for (int i = 1; i < n; i++) {
        const double t = d[i - 1]/b[i - 1];
        b[i] -= t*d[i - 1];
        c[i] -= t*c[i - 1];
    }
    c[n - 1] = c[n - 1]/b[n - 1];
    for (int i = n - 2; i >= 0; i--) c[i] = (c[i] - d[i]*c[i + 1])/b[i];
    b[n - 1] = (Y[n - 1] - Y[n - 2])/d[n - 2] + d[n - 2]*(c[n - 2] + 2.0*c[n - 1]);
    for (int i = 0; i < n - 1; i++) {
        b[i] = (Y[i + 1] - Y[i])/d[i] - d[i]*(c[i + 1] + 2.0*c[i]);
        d[i] = (c[i + 1] - c[i])/d[i];
        c[i] = 3.0*c[i];
    }
    c[n - 1] = 3.0*c[n - 1];
    d[n - 1] = d[n - 2];
}

// Evaluate a cubic spline whose knots are (near-)uniform in ln x, packed as
// spl = [X | Y | b | c | d]; returns 0 outside [X[0], X[n-1]] like Spline::Evaluate
// (Spline.cpp:362-367).
PRSD_HD double cubic_spline_eval_log(int n, const double* spl, double lnx0, double inv_dlnx, double x)
{
    if (!(x >= spl[0]) || x > spl[n - 1]) return 0.0;
    int i = (int)((log(x) - lnx0)*inv_dlnx);
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i > 0 && x < spl[i]) --i;
    while (i < n - 2 && x > spl[i + 1]) ++i;
    const double t = x - spl[i];
    return spl[n + i] + t*(spl[2*n + i] + t*(spl[3*n + i] + t*spl[4*n + i]));
}

// Not-a-knot cubic spline (what SciPy's InterpolatedUnivariateSpline(k=3) / FITPACK build, used by
// the reference's RSDSpline, pyRSD/rsd/tools.py:331-388).  Output in the same piecewise layout as
// cubic_spline_build: y = Y[i] + b t + c t^2 + d t^3 on [X[i], X[i+1]].  `M` and `w` are work
// arrays of length n (n >= 4).
PRSD_HD void nak_spline_build(int n, const double* X, const double* Y, double* b, double* c, double* d,
                              double* M, double* w)
{
    // interior equations for the second derivatives M_1 .. M_{n-2}; M_0 and M_{n-1} eliminated with
    // the not-a-knot conditions (continuous third derivative at X[1] and X[n-2])
    const int m = n - 2;                // unknowns M_1..M_{n-2} stored at M[1..n-2]
    // diagonal in b[], super-diagonal in c[], sub-diagonal in d[], rhs in M[]
    for (int i = 1; i <= m; i++) {
        const double h0 = X[i] - X[i - 1], h1 = X[i + 1] - X[i];
        b[i] = 2.0*(h0 + h1);
        c[i] = h1;
        d[i] = h0;
        M[i] = 6.0*((Y[i + 1] - Y[i])/h1 - (Y[i] - Y[i - 1])/h0);
    }
    {
        const double h0 = X[1] - X[0], h1 = X[2] - X[1];
        b[1] += h0*(1.0 + h0/h1);
        c[1] -= h0*h0/h1;
        const double hn = X[n - 1] - X[n - 2], hm = X[n - 2] - X[n - 3];
        b[m] += hn*(1.0 + hn/hm);
        d[m] -= hn*hn/hm;
    }
    // Thomas algorithm
    w[1] = c[1]/b[1];
    M[1] = M[1]/b[1];
    for (int i = 2; i <= m; i++) {
        const double den = b[i] - d[i]*w[i - 1];
        w[i] = c[i]/den;
        M[i] = (M[i] - d[i]*M[i - 1])/den;
    }
    for (int i = m - 1; i >= 1; i--) M[i] -= w[i]*M[i + 1];
    {
        const double h0 = X[1] - X[0], h1 = X[2] - X[1];
        M[0] = M[1] - (h0/h1)*(M[2] - M[1]);
        const double hn = X[n - 1] - X[n - 2], hm = X[n - 2] - X[n - 3];
        M[n - 1] = M[n - 2] + (hn/hm)*(M[n - 2] - M[n - 3]);
    }
    for (int i = 0; i < n - 1; i++) {
        const double h = X[i + 1] - X[i];
        b[i] = (Y[i + 1] - Y[i])/h - h*(2.0*M[i] + M[i + 1])/6.0;
        c[i] = 0.5*M[i];
        d[i] = (M[i + 1] - M[i])/(6.0*h);
    }
    b[n - 1] = b[n - 2]; c[n - 1] = c[n - 2]; d[n - 1] = d[n - 2];
}

// piecewise-cubic evaluation on knots that are (near-)uniform in ln x; no range check
PRSD_HD double pw_cubic_eval_log(int n, const double* X, const double* Y, const double* b, const double* c,
                                 const double* d, double lnx0, double inv_dlnx, double x)
{
    int i = (int)((log(x) - lnx0)*inv_dlnx);
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i > 0 && x < X[i]) --i;
    while (i < n - 2 && x > X[i + 1]) ++i;
    const double t = x - X[i];
    return Y[i] + t*(b[i] + t*(c[i] + t*d[i]));
}

// Cubic spline on arbitrary increasing knots (binary search), 0 outside the domain like
// Spline::Evaluate (Spline.cpp:362-367); spl = [X | Y | b | c | d].
PRSD_HD double cubic_spline_eval_bsearch(int n, const double* spl, double x)
{
    if (!(x >= spl[0]) || x > spl[n - 1]) return 0.0;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (spl[mid] <= x) lo = mid; else hi = mid;
    }
    const double t = x - spl[lo];
    return spl[n + lo] + t*(spl[2*n + lo] + t*(spl[3*n + lo] + t*spl[4*n + lo]));
}

// Fill the Eisenstein-Hu no-wiggle parameters used for the extrapolation of T(k)
// (Cosmology.cpp:311-362) and the spline end-point ratios (Cosmology.cpp:288-300).
PRSD_HD void plin_params_init(PlinParams& p, int n, const double* k, const double* T, double ns,
                              double amp, double h, double Omega0_m, double Omega0_b, double Tcmb)
{
    p.ns = ns; p.amp = amp; p.h = h; p.n = n; p.pad_ = 0;
    const double h2 = h*h;
    p.omh2 = Omega0_m*h2;
    const double obh2 = Omega0_b*h2;
    const double theta = Tcmb/2.7;
    const double fb = obh2/p.omh2;
    p.keq = 0.0746*p.omh2/(theta*theta);
    p.s_nw = h*44.5*log(9.83/p.omh2)/sqrt(1.0 + 10.0*pow(obh2, 0.75));
    p.alpha_gamma = 1.0 - 0.328*log(431.0*p.omh2)*fb + 0.38*log(22.3*p.omh2)*fb*fb;
    p.k0 = k[0]; p.k1 = k[n - 1];
    p.T0_ratio = T[0]/eh_nowiggle(p, p.k0);
    p.T1_ratio = T[n - 1]/eh_nowiggle(p, p.k1);
    p.lnk0 = log(p.k0);
    p.inv_dlnk = (n - 1)/log(p.k1/p.k0);
}

// ------------------------------------------------------- table interpolation
// 4-point Lagrange weights at fractional position s (in units of the knot spacing,
// measured from the first stencil point; the stencil is s = 0, 1, 2, 3).
PRSD_HD void lagrange4(double s, double* c)
{
    const double s1 = s - 1.0, s2 = s - 2.0, s3 = s - 3.0;
    c[0] = -s1*s2*s3*(1.0/6.0);
    c[1] =  s*s2*s3*0.5;
    c[2] = -s*s1*s3*0.5;
    c[3] =  s*s1*s2*(1.0/6.0);
}

} // namespace prsd
