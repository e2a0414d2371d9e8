/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into the product library.
 *
 * Flat C-ABI over the UNMODIFIED reference C++ classes, compiled from where
 * they lie under /root/reference/pyRSD/_gcl (see oracle/Makefile).  Gives the
 * tests / bench cpu_baseline a ctypes handle on the reference's own
 *   LinearPS, Imn, Jmn, Kmn, ImnOneLoop, OneLoopP{dd,dv,vv,22Bar},
 *   PowerSpectrum::{VelocityDispersion,X_Zel,Y_Zel,Q3_Zel,Sigma},
 *   ZeldovichP{00,01,11}, ZeldovichCF, pk_to_xi / xi_to_pk / ComputeXiLM,
 *   SimpsIntegrate / TrapzIntegrate, CubicSpline.
 * Only `Cosmology` (CLASS wrapper) is replaced by oracle/stub/Cosmology_stub.h
 * and the Fortran FFTLog by oracle/fftlog_restate.c.
 */
#include <functional>
#include <cstring>

#include "LinearPS.h"
#include "Imn.h"
#include "Jmn.h"
#include "Kmn.h"
#include "ImnOneLoop.h"
#include "OneLoopPS.h"
#include "ZeldovichPS.h"
#include "ZeldovichCF.h"
#include "CorrelationFunction.h"
#include "DiscreteQuad.h"
#include "FortranFFTLog.h"
#include "Quadrature.h"
#include "SpecialFunctions.h"
#include "Spline.h"

using namespace std::placeholders;

static parray to_parray(int n, const double* x)
{
    parray p(n);
    for (int i = 0; i < n; i++) p[i] = x[i];
    return p;
}
static void from_parray(const parray& p, double* out)
{
    for (size_t i = 0; i < p.size(); i++) out[i] = p[i];
}

extern "C" {

/* ------------------------------------------------------------ cosmology */
void* ref_cosmo_new(int n, const double* k, const double* T, double sigma8, double n_s,
                    double h, double Om, double Ob, double Tcmb)
{
    return new Cosmology(n, k, T, sigma8, n_s, h, Om, Ob, Tcmb);
}
void ref_cosmo_free(void* c) { delete (Cosmology*)c; }
double ref_cosmo_delta_H(void* c) { return ((Cosmology*)c)->delta_H(); }
void ref_cosmo_set_sigma8_z(void* c, double s) { ((Cosmology*)c)->SetSigma8_z(s); }
void ref_cosmo_transfer(void* c, int n, const double* k, double* out)
{
    for (int i = 0; i < n; i++) out[i] = ((Cosmology*)c)->EvaluateTransfer(k[i]);
}

/* ------------------------------------------------------------- LinearPS */
void* ref_plin_new(void* c, double sigma8_z)
{
    LinearPS* P = new LinearPS(*(Cosmology*)c, 0.);
    P->SetSigma8AtZ(sigma8_z);
    return P;
}
void ref_plin_free(void* P) { delete (LinearPS*)P; }
void ref_ps_eval(void* P, int n, const double* k, double* out)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    for (int i = 0; i < n; i++) out[i] = ps(k[i]);
}
double ref_ps_velocity_dispersion(void* P) { return ((PowerSpectrum*)P)->VelocityDispersion(); }
void ref_ps_sigmasq_k(void* P, double factor, int n, const double* k, double* out)
{
    for (int i = 0; i < n; i++) out[i] = ((PowerSpectrum*)P)->VelocityDispersion(k[i], factor);
}
void ref_ps_sigma_R(void* P, int n, const double* R, double* out)
{
    for (int i = 0; i < n; i++) out[i] = ((PowerSpectrum*)P)->Sigma(R[i]);
}
void ref_ps_xzel(void* P, int n, const double* r, double* out)
{
    for (int i = 0; i < n; i++) out[i] = ((PowerSpectrum*)P)->X_Zel(r[i]);
}
void ref_ps_yzel(void* P, int n, const double* r, double* out)
{
    for (int i = 0; i < n; i++) out[i] = ((PowerSpectrum*)P)->Y_Zel(r[i]);
}
void ref_ps_q3zel(void* P, int n, const double* k, double* out)
{
    for (int i = 0; i < n; i++) out[i] = ((PowerSpectrum*)P)->Q3_Zel(k[i]);
}
/* CorrelationFunction(P, kmin, kmax)(r)  [CorrelationFunction.cpp:145-168] */
void ref_correlation(void* P, double kmin, double kmax, int n, const double* r, double* out)
{
    CorrelationFunction cf(*(PowerSpectrum*)P, kmin, kmax);
    for (int i = 0; i < n; i++) out[i] = cf.Evaluate(r[i]);
}
void ref_ps_sigma3sq(void* P, int n, const double* k, double* out)
{
    for (int i = 0; i < n; i++) out[i] = ((PowerSpectrum*)P)->sigma3_squared(k[i]);
}

/* X_Zel / Y_Zel with caller-chosen tolerances: same integrands and the same
 * reference quadrature template as PowerSpectrum.cpp:94-127, whose tolerances
 * (1e-4, 1e-10) are hard-coded literals there. */
static double Xint(const PowerSpectrum& P, double r, double q)
{
    return P(q)*(-2*SphericalBesselJ1(r*q)/(r*q));
}
static double Yint(const PowerSpectrum& P, double r, double q)
{
    double x = r*q;
    return P(q)*(-2*SphericalBesselJ0(x) + 6*SphericalBesselJ1(x)/x);
}
void ref_ps_xzel_tol(void* P, int n, const double* r, double epsrel, double epsabs, double* out)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    for (int i = 0; i < n; i++)
        out[i] = 1/(2*M_PI*M_PI) * Integrate(std::bind(Xint, std::cref(ps), r[i], _1), 1e-5, 1e2, epsrel, epsabs);
}
void ref_ps_yzel_tol(void* P, int n, const double* r, double epsrel, double epsabs, double* out)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    for (int i = 0; i < n; i++)
        out[i] = 1/(2*M_PI*M_PI) * Integrate(std::bind(Yint, std::cref(ps), r[i], _1), 1e-5, 1e2, epsrel, epsabs);
}

/* Same integrands and the same reference quadrature templates as PowerSpectrum.cpp:53-62,
 * 129-136 and OneLoopPS.cpp:183-193, with caller-chosen tolerances (hard-coded there). */
static double Pint(const PowerSpectrum& P, double q) { return P(q); }
double ref_ps_sigv_tol(void* P, double qmax, double epsrel, double epsabs)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    return 1/(6*M_PI*M_PI) * Integrate<ExpSub>(std::bind(Pint, std::cref(ps), _1), 1e-5, qmax, epsrel, epsabs);
}
static double Q3int(const PowerSpectrum& P, double q) { double p = P(q); return p*p/q/q; }
double ref_ps_q3_tol(void* P, double k, double epsrel, double epsabs)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    return 1/(10*M_PI*M_PI) * Common::pow4(k) * Integrate(std::bind(Q3int, std::cref(ps), _1), 1e-5, 1e2, epsrel, epsabs);
}
static double Kurtint(const OneLoopPS& P, double q) { return P.EvaluateFull(q)/(q*q); }
double ref_p22bar_kurtosis_tol(void* L, double epsrel, double epsabs)
{
    const OneLoopPS& ps = *(OneLoopPS*)L;
    return 0.25/(2*M_PI*M_PI) * Integrate<ExpSub>(std::bind(Kurtint, std::cref(ps), _1), 1e-5, 1e1, epsrel, epsabs);
}
static double SigRint(const PowerSpectrum& P, double R, double k)
{
    double x = k*R;
    double W = (x < 1e-5) ? 1 - (1./30.)*x*x : 3/Common::pow3(x) * (sin(x) - x*cos(x));
    return k*k/(2*M_PI*M_PI) * P(k) * W*W;
}
double ref_ps_sigma_R_tol(void* P, double R, double epsrel, double epsabs)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    return sqrt(Integrate<ExpSub>(std::bind(SigRint, std::cref(ps), R, _1), 1e-5, 1e2, epsrel, epsabs));
}

/* --------------------------------------------------------- Imn/Jmn/Kmn */
void ref_imn(void* P, double epsrel, int m, int n, int nk, const double* k, double* out)
{
    Imn I(*(PowerSpectrum*)P, epsrel);
    for (int i = 0; i < nk; i++) out[i] = I(k[i], m, n);
}
void ref_jmn(void* P, double epsrel, int m, int n, int nk, const double* k, double* out)
{
    Jmn J(*(PowerSpectrum*)P, epsrel);
    for (int i = 0; i < nk; i++) out[i] = J(k[i], m, n);
}
void ref_kmn(void* P, double epsrel, int m, int n, int tidal, int part, int nk, const double* k, double* out)
{
    Kmn K(*(PowerSpectrum*)P, epsrel);
    for (int i = 0; i < nk; i++) out[i] = K(k[i], m, n, (bool)tidal, part);
}

/* ------------------------------------------------------------ OneLoopPS */
/* which: 0 = Pdd, 1 = Pdv, 2 = Pvv, 3 = P22bar */
void* ref_oneloop_new(void* P, int which, double epsrel)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    switch (which) {
        case 0: return (OneLoopPS*) new OneLoopPdd(ps, epsrel);
        case 1: return (OneLoopPS*) new OneLoopPdv(ps, epsrel);
        case 2: return (OneLoopPS*) new OneLoopPvv(ps, epsrel);
        case 3: return (OneLoopPS*) new OneLoopP22Bar(ps, epsrel);
    }
    return 0;
}
void* ref_oneloop_from_table(void* P, int which, double epsrel, int n, const double* table)
{
    const PowerSpectrum& ps = *(PowerSpectrum*)P;
    parray t = to_parray(n, table);
    switch (which) {
        case 0: return (OneLoopPS*) new OneLoopPdd(ps, epsrel, t);
        case 1: return (OneLoopPS*) new OneLoopPdv(ps, epsrel, t);
        case 2: return (OneLoopPS*) new OneLoopPvv(ps, epsrel, t);
        case 3: return (OneLoopPS*) new OneLoopP22Bar(ps, epsrel, t);
    }
    return 0;
}
void ref_oneloop_free(void* L) { delete (OneLoopPS*)L; }
void ref_oneloop_table(void* L, double* out) { from_parray(((OneLoopPS*)L)->GetOneLoopPower(), out); }
void ref_oneloop_eval(void* L, int full, int nk, const double* k, double* out)
{
    OneLoopPS* ps = (OneLoopPS*)L;
    for (int i = 0; i < nk; i++) out[i] = full ? ps->EvaluateFull(k[i]) : ps->Evaluate(k[i]);
}
double ref_p22bar_kurtosis(void* L) { return ((OneLoopP22Bar*)(OneLoopPS*)L)->VelocityKurtosis(); }

/* ----------------------------------------------------------- ImnOneLoop */
/* which: 0 = EvaluateLinear, 1 = EvaluateCross, 2 = EvaluateOneLoop; P2 may be NULL */
void ref_imn_oneloop(void* P1, void* P2, double epsrel, int which, int m, int n,
                     int nk, const double* k, double* out)
{
    OneLoopPS& p1 = *(OneLoopPS*)P1;
    ImnOneLoop* I = P2 ? new ImnOneLoop(p1, *(OneLoopPS*)P2, epsrel) : new ImnOneLoop(p1, epsrel);
    for (int i = 0; i < nk; i++) {
        if (which == 0)      out[i] = I->EvaluateLinear(k[i], m, n);
        else if (which == 1) out[i] = I->EvaluateCross(k[i], m, n);
        else                 out[i] = I->EvaluateOneLoop(k[i], m, n);
    }
    delete I;
}

/* ------------------------------------------------------------ Zel'dovich */
void* ref_zel_new_full(void* c, int approx_lowk)
{
    /* full constructor: computes sigma_sq, X0, XX, YY (ZeldovichPS.cpp:15-30) */
    return new ZeldovichPS(*(Cosmology*)c, 0., (bool)approx_lowk);
}
void* ref_zel_new(void* c, int approx_lowk, double sigma8_z, double k0_low, double sigmasq,
                  int n, const double* X0, const double* XX, const double* YY)
{
    return new ZeldovichPS(*(Cosmology*)c, (bool)approx_lowk, sigma8_z, k0_low, sigmasq,
                           to_parray(n, X0), to_parray(n, XX), to_parray(n, YY));
}
void ref_zel_free(void* Z) { delete (ZeldovichPS*)Z; }
void ref_zel_state(void* Z, double* sigmasq, double* X0, double* XX, double* YY)
{
    ZeldovichPS* z = (ZeldovichPS*)Z;
    *sigmasq = z->GetSigmaSq();
    from_parray(z->GetX0Zel(), X0);
    from_parray(z->GetXZel(), XX);
    from_parray(z->GetYZel(), YY);
}
void ref_zel_set_sigma8(void* Z, double s) { ((ZeldovichPS*)Z)->SetSigma8AtZ(s); }
/* which: 0 = P00, 1 = P01, 2 = P11 ; lowk: -1 keep flag, 0/1 set */
void ref_zel_eval(void* Z, int which, int lowk, double k0_low, int nk, const double* k, double* out)
{
    ZeldovichPS* base = (ZeldovichPS*)Z;
    ZeldovichPS* z;
    if (which == 0) z = new ZeldovichP00(*base);
    else if (which == 1) z = new ZeldovichP01(*base);
    else z = new ZeldovichP11(*base);
    if (lowk >= 0) z->SetLowKApprox((bool)lowk);
    if (k0_low > 0) z->SetLowKTransition(k0_low);
    for (int i = 0; i < nk; i++) out[i] = z->Evaluate(k[i]);
    delete z;
}
void ref_zelcf_eval(void* Z, double kmin, double kmax, double smoothing, int nr, const double* r, double* out)
{
    ZeldovichCF cf(*(ZeldovichPS*)Z, kmin, kmax);
    from_parray(cf.EvaluateMany(to_parray(nr, r), smoothing), out);
}

/* --------------------------------------------- correlation-function tools */
void ref_pk_to_xi(int ell, int nk, const double* k, const double* pk, int nr, const double* r,
                  double smoothing, int method, double* out)
{
    from_parray(pk_to_xi(ell, to_parray(nk, k), to_parray(nk, pk), to_parray(nr, r), smoothing,
                         (IntegrationMethods::Type)method), out);
}
void ref_xi_to_pk(int ell, int nr, const double* r, const double* xi, int nk, const double* k,
                  double smoothing, int method, double* out)
{
    from_parray(xi_to_pk(ell, to_parray(nr, r), to_parray(nr, xi), to_parray(nk, k), smoothing,
                         (IntegrationMethods::Type)method), out);
}
void ref_compute_xilm(int l, int m, int nk, const double* k, const double* pk, int nr, const double* r,
                      double smoothing, int method, double* out)
{
    from_parray(ComputeXiLM(l, m, to_parray(nk, k), to_parray(nk, pk), to_parray(nr, r), smoothing,
                            (IntegrationMethods::Type)method), out);
}
void ref_compute_xilm_fftlog(int l, int m, int n, const double* k, const double* pk,
                             double* r, double* xi, double q, double smoothing)
{
    ComputeXiLM_fftlog(l, m, to_parray(n, k), to_parray(n, pk), r, xi, q, smoothing);
}
/* raw FFTLog through the reference's own wrapper class */
double ref_fftlog(int n, double dlnr, double mu, double q, double kr, int kropt, int dir, double* a)
{
    FortranFFTLog F(n, dlnr, mu, q, kr, kropt);
    parray p = to_parray(n, a);
    F.Transform(p, dir);
    from_parray(p, a);
    return F.KR();
}

/* --------------------------------------------------- discrete quadrature */
double ref_simps(int n, const double* x, const double* y) { return SimpsIntegrate(to_parray(n, x), to_parray(n, y)); }
double ref_trapz(int n, const double* x, const double* y) { return TrapzIntegrate(to_parray(n, x), to_parray(n, y)); }

/* ---------------------------------------------------------------- spline */
void ref_cubic_spline(int n, const double* x, const double* y, int nq, const double* xq, double* out)
{
    Spline s = CubicSpline(n, x, y);
    for (int i = 0; i < nq; i++) out[i] = s(xq[i]);
}
void ref_logspace(double a, double b, int n, double* out) { from_parray(parray::logspace(a, b, n), out); }

} /* extern "C" */
