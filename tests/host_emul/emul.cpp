// TEST INFRASTRUCTURE ONLY.  CPU emulation of the operator algebra the CUDA kernels run,
// built from the SAME headers (pt_math.cuh, quad.cpp).  It lets the CPU-only test-suite
// validate the host logic (knot grid, panel generator, kernel formulae, table reads) against
// the oracle without a GPU.  The shipped library never links or calls this.
#include "../../pyrsd_b200/csrc/pt_math.cuh"
#include "../../pyrsd_b200/csrc/quad.h"
#include <cmath>
#include <vector>

using namespace prsd;

static QuadConfig g_cfg;
static int g_w_float = 0;       // experiment: round the operator entries W = weight x kernel to binary32
static KnotGrid g_grid = make_knot_grid();

extern "C" {

void emul_set_config(double qmin, int n_outer, int n_inner, double ln_far, double ln_mid,
                     double bao_lo, double bao_hi, double bao_dq)
{
    g_cfg = QuadConfig();
    g_cfg.qmin = qmin; g_cfg.ln_far = ln_far; g_cfg.ln_mid = ln_mid;
    g_cfg.bao_lo = bao_lo; g_cfg.bao_hi = bao_hi; g_cfg.bao_dq = bao_dq;
    g_cfg.ln_thin = 0.0;
    g_cfg.set_uniform(n_outer, n_inner, n_inner);
}

void emul_set_config_default() { g_cfg = QuadConfig(); }
void emul_set_w_float(int on) { g_w_float = on; }
void emul_use_ml_variant() { g_cfg = g_cfg.ml_variant(); }

void emul_set_tier(int i, double k_below, int n_outer, int n_inner, int n_thin)
{
    g_cfg.tier[i] = {k_below, n_outer, n_inner, n_thin};
}

void emul_set_config_ext(int n_inner_thin, double ln_thin, double ln_inner, double mid_lo, double mid_hi)
{
    for (int i = 0; i < QuadConfig::NTIER; i++) g_cfg.tier[i].n_thin = n_inner_thin;
    g_cfg.ln_thin = ln_thin; g_cfg.ln_inner = ln_inner;
    g_cfg.mid_lo = mid_lo; g_cfg.mid_hi = mid_hi;
}

int emul_knot_count() { return g_grid.size(); }
void emul_knots(double* x) { for (int i = 0; i < g_grid.size(); i++) x[i] = g_grid.x[i]; }

void emul_plin_table(int n, const double* k, const double* T, double ns, double amp, double h,
                     double Om, double Ob, double Tcmb, double* tab)
{
    PlinParams p;
    plin_params_init(p, n, k, T, ns, amp, h, Om, Ob, Tcmb);
    std::vector<double> spl(5*n);
    for (int i = 0; i < n; i++) { spl[i] = k[i]; spl[n + i] = T[i]; }
    cubic_spline_build(n, &spl[0], &spl[n], &spl[2*n], &spl[3*n], &spl[4*n]);
    for (int j = 0; j < g_grid.size(); j++) tab[j] = plin_eval(p, spl.data(), std::exp(g_grid.x[j]));
}

void emul_plin_eval(int n, const double* k, const double* T, double ns, double amp, double h,
                    double Om, double Ob, double Tcmb, int nq, const double* q, double* out)
{
    PlinParams p;
    plin_params_init(p, n, k, T, ns, amp, h, Om, Ob, Tcmb);
    std::vector<double> spl(5*n);
    for (int i = 0; i < n; i++) { spl[i] = k[i]; spl[n + i] = T[i]; }
    cubic_spline_build(n, &spl[0], &spl[n], &spl[2*n], &spl[3*n], &spl[4*n]);
    for (int j = 0; j < nq; j++) out[j] = plin_eval(p, spl.data(), q[j]);
}

void emul_table_read(const double* tab, int nq, const double* q, double* out)
{
    for (int i = 0; i < nq; i++) {
        int j0; double s, c[4];
        g_grid.locate(std::log(q[i]), j0, s);
        lagrange4(s, c);
        out[i] = c[0]*tab[j0] + c[1]*tab[j0+1] + c[2]*tab[j0+2] + c[3]*tab[j0+3];
    }
}

void emul_spline(int n, const double* x, const double* y, int nq, const double* xq, double* out)
{
    std::vector<double> spl(5*n);
    for (int i = 0; i < n; i++) { spl[i] = x[i]; spl[n + i] = y[i]; }
    cubic_spline_build(n, &spl[0], &spl[n], &spl[2*n], &spl[3*n], &spl[4*n]);
    const double lnx0 = std::log(x[0]), inv = (n - 1)/std::log(x[n-1]/x[0]);
    for (int i = 0; i < nq; i++) out[i] = cubic_spline_eval_log(n, spl.data(), lnx0, inv, xq[i]);
}

// mode 0: Imn convention, 1: Kmn convention, 2: ImnOneLoop convention (see above)
void emul_ik(int kid, int nk, const double* k, double qmax, const double* tabQ, const double* tabR,
             int mode, double* out, long long* nnodes)
{
    HalfNodes H;
    build_half_nodes(k, nk, qmax, g_cfg, g_grid, H);
    *nnodes = (long long)H.n_nodes();
    const bool sym = kernel_is_symmetric(kid);
    for (int ik = 0; ik < nk; ik++) {
        const double kk = k[ik];
        double pos = 0.0, neg = 0.0;
        for (int64_t j = H.node_off[ik]; j < H.node_off[ik+1]; j++) {
            const int32_t o = H.n_outer[j];
            const double a = H.a[o], b = H.b[j];
            const double* ca = &H.a_coef[4*o];
            const double* cb = &H.b_coef[4*j];
            const int ja = H.a_j0[o], jb = H.b_j0[j];
            const double u = (a + b)/kk, vp = (b - a)/kk;
            const double Qa = ca[0]*tabQ[ja] + ca[1]*tabQ[ja+1] + ca[2]*tabQ[ja+2] + ca[3]*tabQ[ja+3];
            const double Rb = cb[0]*tabR[jb] + cb[1]*tabR[jb+1] + cb[2]*tabR[jb+2] + cb[3]*tabR[jb+3];
            { double W = H.wgt[j]*pt_kernel(kid, u, vp, 2*a/kk, 2*b/kk); if (g_w_float) W = (double)(float)W; pos += W*Qa*Rb; }
            if (mode == 2 || !sym) {
                const double Qb = cb[0]*tabQ[jb] + cb[1]*tabQ[jb+1] + cb[2]*tabQ[jb+2] + cb[3]*tabQ[jb+3];
                const double Ra = ca[0]*tabR[ja] + ca[1]*tabR[ja+1] + ca[2]*tabR[ja+2] + ca[3]*tabR[ja+3];
                { double W = H.wgt[j]*pt_kernel(kid, u, -vp, 2*b/kk, 2*a/kk); if (g_w_float) W = (double)(float)W; neg += W*Qb*Ra; }
            }
        }
        const double V8 = kk/(8*M_PI*M_PI), V4 = kk/(4*M_PI*M_PI);
        if (mode == 0)      out[ik] = sym ? 2*V8*pos : V8*(pos + neg);
        else if (mode == 1) out[ik] = sym ? V4*pos : 0.5*V4*(pos + neg);
        else                out[ik] = V8*(pos + neg);
    }
}

} // extern "C"

// ----------------------------------------------------------------- 1-D operators
#include "../../pyrsd_b200/csrc/linear_math.cuh"
extern "C" void emul_linear(int type, int sub, int n, const double* param, const double* lo, const double* hi,
                            const double* scale, const double* tab, double* out)
{
    std::vector<double> glx, glw;
    gauss_legendre(GLN, glx, glw);
    ZonesDev Z;
    Z.nz = (int)g_grid.zones.size();
    for (int i = 0; i < Z.nz; i++) {
        Z.x0[i] = g_grid.zones[i].x0; Z.h[i] = g_grid.zones[i].h; Z.n[i] = g_grid.zones[i].n; Z.start[i] = g_grid.zones[i].start;
    }
    #pragma omp parallel for schedule(dynamic)
    for (int o = 0; o < n; o++) {
        LinOutDev S = {type, sub, param[o], lo[o], hi[o], scale[o]};
        double acc = 0.0;
        for (int j = 0; j < g_grid.size(); j++) {
            if (tab[j] == 0.0) continue;
            acc += tab[j]*linear_weight(S, Z, j, glx.data(), glw.data());
        }
        out[o] = acc;
    }
}

// same as emul_ik (Imn convention, symmetric kernels only) but with P_L evaluated exactly at the
// nodes instead of read from the table: isolates the quadrature error from the table error.
extern "C" void emul_ik_exactP(int kid, int nk, const double* k, double qmax, int n, const double* ki, const double* Ti,
                               double ns, double amp, double h, double Om, double Ob, double Tcmb, double* out)
{
    PlinParams p;
    plin_params_init(p, n, ki, Ti, ns, amp, h, Om, Ob, Tcmb);
    std::vector<double> spl(5*n);
    for (int i = 0; i < n; i++) { spl[i] = ki[i]; spl[n + i] = Ti[i]; }
    cubic_spline_build(n, &spl[0], &spl[n], &spl[2*n], &spl[3*n], &spl[4*n]);
    HalfNodes H;
    build_half_nodes(k, nk, qmax, g_cfg, g_grid, H);
    for (int ik = 0; ik < nk; ik++) {
        const double kk = k[ik];
        double pos = 0.0;
        int32_t last_o = -1; double Pa = 0;
        for (int64_t j = H.node_off[ik]; j < H.node_off[ik+1]; j++) {
            const int32_t o = H.n_outer[j];
            const double a = H.a[o], b = H.b[j];
            if (o != last_o) { Pa = plin_eval(p, spl.data(), a); last_o = o; }
            if (H.wgt[j] == 0.0) continue;
            pos += H.wgt[j]*Pa*plin_eval(p, spl.data(), b)*pt_kernel(kid, (a + b)/kk, (b - a)/kk, 2*a/kk, 2*b/kk);
        }
        out[ik] = 2*kk/(8*M_PI*M_PI)*pos;
    }
}

extern "C" void emul_nak_spline(int n, const double* x, const double* y, int nq, const double* xq, double* out)
{
    std::vector<double> b(n), c(n), d(n), M(n), w(n);
    nak_spline_build(n, x, y, b.data(), c.data(), d.data(), M.data(), w.data());
    const double lnx0 = std::log(x[0]), inv = (n - 1)/std::log(x[n-1]/x[0]);
    for (int i = 0; i < nq; i++) out[i] = pw_cubic_eval_log(n, x, y, b.data(), c.data(), d.data(), lnx0, inv, xq[i]);
}

// node census for quadrature tuning: per node the global outer index and the stencil start
extern "C" long long emul_nodes(int nk, const double* k, double qmax, long long cap, int* outer, int* j0, double* wgt)
{
    HalfNodes H;
    build_half_nodes(k, nk, qmax, g_cfg, g_grid, H);
    const long long n = (long long)H.n_nodes();
    for (long long i = 0; i < n && i < cap; i++) { outer[i] = H.n_outer[i]; j0[i] = H.b_j0[i]; wgt[i] = H.wgt[i]; }
    return n;
}

// thinned k lists: keep mask and fill map of PTPlan (quad.cpp); which = 0: thin_k_list(k); 1 / 2: the one-loop grid rule with
// (max_gap_lo, max_gap_hi) = (12, 6) for the 2-D integrals / the same with every point above 1.2 for J
extern "C" int emul_fill_map(int which, int n, const double* k, char* keep_out, int* idx, double* w)
{
    std::vector<double> kk(k, k + n), ksub, mw;
    std::vector<int> midx;
    std::vector<char> keep;
    if (which == 0) keep = thin_k_list(kk);
    else {
        keep.assign(n, 1);
        int lo = 0, hi = n - 1;
        while (lo < n - 1 && kk[lo] < 0.01) lo++;
        while (hi > 0 && kk[hi] > 1.2) hi--;
        thin_graded(keep, lo, -1, 12);
        for (int i = hi + 1; i < n; i++) keep[i] = ((i - hi) % 6 == 0) ? 1 : 0;
        keep[n - 1] = 1;
        if (which == 2) for (int i = hi; i < n; i++) keep[i] = 1;
    }
    lagrange_fill_map(kk, keep, midx, mw, ksub);
    for (int i = 0; i < n; i++) keep_out[i] = keep[i];
    for (size_t i = 0; i < midx.size(); i++) { idx[i] = midx[i]; w[i] = mw[i]; }
    return (int)ksub.size();
}
