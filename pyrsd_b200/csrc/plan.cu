// plan.cu -- PT plan construction and the batched PT stage (see plan.cuh).
#include "plan.cuh"
#include "nodal.cuh"

#include <chrono>
#include <cmath>
#include <future>

namespace prsd {

// parray::logspace (pyRSD/_gcl/cpp/parray.cpp:101-121): exp(ln min + i dln), exact end points
static std::vector<double> logspace_ref(double lo, double hi, int n)
{
    std::vector<double> x(n);
    const double l0 = std::log(lo), l1 = std::log(hi);
    for (int i = 0; i < n; i++) x[i] = std::exp(l0 + i*(l1 - l0)/(n - 1));
    x[0] = lo;
    x[n - 1] = hi;
    return x;
}

std::vector<double> oneloop_kgrid() { return logspace_ref(ONELOOP_KMIN, ONELOOP_KMAX, ONELOOP_N); }

// ZeldovichPS::InitializeR (ZeldovichPS.cpp:41-53)
std::vector<double> zeldovich_rgrid()
{
    std::vector<double> r(ZEL_NR);
    const double nc = 0.5*(ZEL_NR + 1), logrmin = -2.0, logrmax = 5.0;
    const double logrc = 0.5*(logrmin + logrmax), dlogr = (logrmax - logrmin)/ZEL_NR;
    for (int i = 1; i <= ZEL_NR; i++) r[i - 1] = std::pow(10.0, logrc + (i - nc)*dlogr);
    return r;
}

void PTResult::ensure(int ncosmo_, int nk_)
{
    ncosmo = ncosmo_;
    nk = nk_;
    const size_t c = ncosmo;
    I.ensure(c*16*nk);
    J.ensure(c*6*nk);
    K.ensure(c*13*nk);
    ML.ensure(c*N_ML*3*nk);
    oneloop.ensure(c*4*ONELOOP_N);
    sigmasq_k.ensure(c*nk);
    scalars.ensure(c*4);
    X0.ensure(c*ZEL_NR);
    YY.ensure(c*ZEL_NR);
}

// (lagrange_fill_map, thin_graded, thin_k_list: the thinned k lists and their interpolation maps live in quad.cpp -- host code,
// exercised on the CPU by tests/test_host_logic.py)

static const double INV8PI2 = 1.0/(8.0*M_PI*M_PI);
static const double INV4PI2 = 1.0/(4.0*M_PI*M_PI);
static const int J_IDS[6] = {0, 1, 2, 3, 4, 6};                    // J00 J01 J02 J10 J11 J20 -> 3m+n
static const int ML_KID[N_ML] = {KID_H01, KID_H02, KID_K20A, KID_K20B, KID_F23, KID_F32, KID_F33};

PTPlan::PTPlan(Context& c, const double* k, int nk_, const QuadConfig& q, bool ml, bool zel, bool only1l)
    : ctx(&c), cfg(q), nk(nk_), with_ml(ml), with_zel(zel), only_oneloop(only1l)
{
    auto t0 = std::chrono::steady_clock::now();
    auto t_lap = t0;
    PRSD_REQUIRE(nk > 0 || only_oneloop, "PTPlan: empty k grid");
    if (nk > 0) k_interp.assign(k, k + nk);
    k_1loop = oneloop_kgrid();
    r_zel = zeldovich_rgrid();
    const KnotGrid& grid = knot_grid();

    auto lap = [&](int slot) {
        c.sync();
        const auto now = std::chrono::steady_clock::now();
        build_stage_ms[slot] += std::chrono::duration<double, std::milli>(now - t_lap).count();
        t_lap = now;
    };
    // ---- the k lists of the three 2-D operators
    // The 2-D integrals I00, I01, I11, I23, I32, I33 and J00, J01, J11 are smooth functions of k outside the BAO range:
    // below k = 0.01 (where the 2-D integrals are < 1e-3 P_L anyway) and above k = 1.2 only a subset of the reference's
    // 1000-point grid is integrated and the others are filled by 8-point Lagrange interpolation in ln k
    // (k_oneloop_combine); P_L is evaluated at every point.  The interpolation error is below the point-to-point noise
    // of the quadrature itself (measured on the dense fixtures: tables unchanged to < 4e-6 for k <= 1, < 3e-5 of |table|
    // for k > 1 where 2 I + 6 k^2 J P_L cancels 140 x).
    {
        std::vector<char> keep(ONELOOP_N, 1);
        int lo = 0, hi = ONELOOP_N - 1;
        while (lo < ONELOOP_N - 1 && k_1loop[lo] < 0.01) lo++;
        while (hi > 0 && k_1loop[hi] > 1.2) hi--;
        // down to every 12th point below k = 0.01 (0.17 apart in ln k) and every 6th above 1.2 (0.08 apart), graded
        // (thin_graded): ~420 of the 1000 points.  Measured on the dense fixtures
        // (profiles/experiments/ol_thinning_exp.sh): tables unchanged for k <= 1 against (6, 3), 1.6e-5 ... 3.1e-5 of
        // |table| above k = 1 (where 2 I + 6 k^2 J P_L cancels 140 x) against 1.7e-5 ... 3.1e-5; (12, 9) costs 3.8e-5 there
        // for no further gain.
        thin_graded(keep, lo, -1, 12);
        // (no grading at the upper end: the quasi-uniform stencils of lagrange_fill_map on every 6th point measured 2.1e-5
        // ... 3.1e-5 there, graded gaps 2.5e-5 ... 3.7e-5)
        for (int i = hi + 1; i < ONELOOP_N; i++) keep[i] = ((i - hi) % 6 == 0) ? 1 : 0;
        keep[ONELOOP_N - 1] = 1;
        std::vector<double> ksub, mw;
        std::vector<int> midx;
        lagrange_fill_map(k_1loop, keep, midx, mw, ksub);
        n_1loop_sub = (int)ksub.size();
        d_olmap_idx.upload(midx, c.stream);
        d_olmap_w.upload(mw, c.stream);
        k_1loop_sub = ksub;
        // J00, J01, J11: the same subset below k = 1.2 (J_mn(k) is a smooth function of k: the BAO structure of P_L is
        // integrated over), every point above -- 6 k^2 J P_L is one side of the 140-fold cancellation there and the fill
        // would cost a third of the margin to the 5e-5 bound (measured: 1.9e-5 -> 2.6e-5 ... 3.6e-5)
        for (int i = hi; i < ONELOOP_N; i++) keep[i] = 1;
        lagrange_fill_map(k_1loop, keep, midx, mw, k_1loop_j);
        d_jmap_idx.upload(midx, c.stream);
        d_jmap_w.upload(mw, c.stream);
    }
    // Thinned k_interp.  Below k = 0.01 every mode-coupling integral is < 1e-3 P_L(k) and a smooth function of k (no
    // BAO structure enters yet), but a k there costs as many CTA set-ups, P(a) passes and reductions as one at k = 0.3
    // for 1/15 of the nodes: on a long ascending list (rsd's 250-point k_interp has 156 points below 0.01) the operators
    // are integrated at points up to 0.29 apart in ln k there (every 6th of k_interp, graded: thin_graded) and the others are
    // filled by 8-point Lagrange interpolation in ln k (k_pack_ik, k_fill_ml).  Measured on the dense fixtures (tight
    // reference on all 250 k): the fill changes no I, K or ImnOneLoop value by more than 1.4e-7 of max(|value|, P_L).
    if (!only_oneloop) {
        std::vector<char> keep = thin_k_list(k_interp);
        std::vector<int> midx;
        std::vector<double> mw;
        lagrange_fill_map(k_interp, keep, midx, mw, k_sub);
        nk_sub = (int)k_sub.size();
        d_kmap_idx.upload(midx, c.stream);
        d_kmap_w.upload(mw, c.stream);
    }
    // ---- node geometry of the three operators on host threads (no CUDA state is touched there) ...
    const bool want_ik = !only_oneloop, want_ml = with_ml && !only_oneloop;
    const QuadConfig cfg_ml = cfg.ml_variant();
    auto f_1loop = std::async(std::launch::async, [&] { return make_geometry_host(k_1loop_sub.data(), n_1loop_sub, 1e5, cfg, grid, 16); });
    std::future<std::shared_ptr<GeometryHost>> f_ik, f_ml;
    // DMMA lane layout for I / K: 4 nodes per k-step
    if (want_ik) f_ik = std::async(std::launch::async, [&] { return make_geometry_host(k_sub.data(), nk_sub, 1e5, cfg, grid, 4); });
    if (want_ml) f_ml = std::async(std::launch::async, [&] { return make_geometry_host(k_sub.data(), nk_sub, 10.0, cfg_ml, grid, 16); });
    // ---- ... while the device integrates the 1-D operators (product integration, ~0.2 s of kernels, nothing waits on them)
    try {
        {
            // ONE operator for every 1-D functional of P_L -- [J00, J01, J11 on their subset of the one-loop grid (k_1loop_j,
            // filled to all 1000 points in k_oneloop_combine) | J_mn x 6, sigma_v^2(k), sigma_v^2, X(r), Y(r)] -- so that a
            // batch is ONE GEMM launch of 73 x (ncosmo / 64) CTAs: two launches of 26 and 48 column tiles left half of the
            // SMs idle
            std::vector<LinOutSpec> o;
            for (int id : {0, 1, 4})
                for (double kj : k_1loop_j) o.push_back({LK_JMN, id, kj, 1e-5, 1e5, 1.0});
            n_lin1 = (int)o.size();
            if (!only_oneloop) {
                // J_mn on the same thinned list as the 2-D operators (filled in k_pack_lin); sigma_v^2(k), whose upper
                // limit moves with k, on every point
                for (int j = 0; j < 6; j++)
                    for (int i = 0; i < nk_sub; i++) o.push_back({LK_JMN, J_IDS[j], k_sub[i], 1e-5, 1e5, 1.0});
                // sigma_v^2(k) integrates from 1e-5 up to 0.5 k (PowerSpectrum.cpp:60-62)
                for (int i = 0; i < nk; i++) {
                    // for 0.5 k < 1e-5 the reference integrates "backwards" and returns a negative number
                    const double hi = 0.5*k_interp[i];
                    if (hi >= 1e-5) o.push_back({LK_SIGV, 0, 0.0, 1e-5, hi, 1.0});
                    else            o.push_back({LK_SIGV, 0, 0.0, hi, 1e-5, -1.0});
                }
                o.push_back({LK_SIGV, 0, 0.0, 1e-5, 1e2, 1.0});
                if (with_zel) {
                    // (X(r), Y(r) on a thinned r grid -- every 2nd point for 1 <= r <= 1000, every 4th outside, or the same
                    // with every point in [10, 400] -- with an 8-point Lagrange fill in ln r: 1-D operators 0.29 -> 0.21 / 0.24
                    // ms, but P(k, mu) moves by 6.8e-6 ... 1.35e-5 at k = 0.05: the Zel'dovich sum amplifies 1e-10 sigma_v^2
                    // in X, Y.  Measured and reverted, round 2.)
                    for (int i = 0; i < ZEL_NR; i++) o.push_back({LK_XZEL, 0, r_zel[i], 1e-5, 1e2, 1.0});
                    for (int i = 0; i < ZEL_NR; i++) o.push_back({LK_YZEL, 0, r_zel[i], 1e-5, 1e2, 1.0});
                }
            }
            lin_all.build(c, o, false);
        }
        if (!only_oneloop) {
            lin_invq2_100.build(c, {{LK_INVQ2, 0, 0.0, 1e-5, 1e2, 1.0/(10.0*M_PI*M_PI)}}, false);
            lin_kurt.build(c, {{LK_INVQ2, 0, 0.0, 1e-5, 1e1, 0.25/(2.0*M_PI*M_PI)}}, false);
            d_kinterp.upload(k_interp, c.stream);
        }
        {
            const auto now = std::chrono::steady_clock::now();      // host time to enqueue the 1-D builds
            build_stage_ms[2] += std::chrono::duration<double, std::milli>(now - t_lap).count();
            t_lap = now;
        }
        // ---- one-loop tables: I00 I01 I11 I23 I32 I33 (all symmetric: 2 x the v > 0 half)
        {
            auto geo = upload_geometry(c, *f_1loop.get());
            std::vector<ColSpec> cols;
            for (int id : {KID_F00, KID_F01, KID_F11, KID_F23, KID_F32, KID_F33}) cols.push_back({id, +1, 2.0*INV8PI2});
            op_1loop.build(c, geo, cols, true);
            // OneLoopP22Bar only ever needs I23 + 2/3 I32 + 1/5 I33 (OneLoopPS.cpp:164-185): fold the three columns
            op_1loop.fold_columns(c, 3, {{4, 2.0/3.0}, {5, 1.0/5.0}}, 4);
            op_1loop.to_float(c);
        }
        // ---- I_mn, K_mn on k_interp (Imn.cpp:120-140, Kmn.cpp:118-138 conventions)
        if (want_ik) {
            auto geo = upload_geometry(c, *f_ik.get());
            std::vector<ColSpec> cols;
            // both spectra are P_L here, so P(a) P(b) is the same on the two halves of the v domain and the
            // non-symmetric kernels simply use f(u, v) + f(u, -v) (ColSpec sign 0): 29 columns, 4 n-tiles
            for (int id = 0; id < 16; id++) cols.push_back({id, kernel_is_symmetric(id) ? +1 : 0, kernel_is_symmetric(id) ? 2.0*INV8PI2 : INV8PI2});
            for (int id = KID_K00; id <= KID_K20SB; id++) cols.push_back({id, kernel_is_symmetric(id) ? +1 : 0, kernel_is_symmetric(id) ? INV4PI2 : 0.5*INV4PI2});
            op_ik.build(c, geo, cols);
        }
        // ---- ImnOneLoop: QMAX = 10, v in [-1, 1] (ImnOneLoop.cpp:12-13, 82-93)
        if (want_ml) {
            auto geo = upload_geometry(c, *f_ml.get());
            op_ml.build(c, geo, ml_columns(), true);
            op_ml.to_float(c);
        }
    } catch (...) {
        // the host tasks reference locals of this frame: let them finish before unwinding
        if (f_1loop.valid()) f_1loop.wait();
        if (f_ik.valid()) f_ik.wait();
        if (f_ml.valid()) f_ml.wait();
        throw;
    }
    {
        const auto now = std::chrono::steady_clock::now();          // host geometry (the longest of the three) + uploads
        build_stage_ms[0] += std::chrono::duration<double, std::milli>(now - t_lap).count();
        t_lap = now;
    }
    d_k1loop.upload(k_1loop, c.stream);
    lap(4);
    build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

size_t PTPlan::device_bytes() const
{
    size_t b = 0;
    for (const BilinearOp* op : {&op_1loop, &op_ik, &op_ml}) {
        if (!op->geo) continue;
        b += op->W.n*sizeof(double) + op->Wf.n*sizeof(float4) + op->Wh.n*sizeof(float);
    }
    if (op_1loop.geo) b += op_1loop.geo->bytes();
    if (op_ik.geo) b += op_ik.geo->bytes();
    if (op_ml.geo) b += op_ml.geo->bytes();
    for (const LinearOp* l : {&lin_all, &lin_invq2_100, &lin_kurt}) b += l->Omega.n*sizeof(double);
    return b;
}

long long PTPlan::nodes_total() const
{
    long long n = 0;
    if (op_1loop.geo) n += op_1loop.geo->n_nodes;
    if (op_ik.geo) n += op_ik.geo->n_nodes;
    if (op_ml.geo) n += 2*op_ml.geo->n_nodes;
    return n;
}

// ------------------------------------------------------------------ packing
__global__ void k_oneloop_combine(int ncosmo, int nsub, const double* __restrict__ op1 /*[c][nsub][8]*/,
                                  const int* __restrict__ midx /*[1000][8]*/, const double* __restrict__ mw /*[1000][8]*/,
                                  int nj, const int* __restrict__ jidx /*[1000][8]*/, const double* __restrict__ jw /*[1000][8]*/,
                                  const double* __restrict__ lin1 /*[c][lin_stride]: 3 x nj first*/, int lin_stride,
                                  const double* __restrict__ plk /*[c][1000]*/,
                                  const double* __restrict__ k1, double* __restrict__ ol /*[c][4][1000]*/)
{
    const int i = blockIdx.x*blockDim.x + threadIdx.x, c = blockIdx.y;
    if (i >= ONELOOP_N) return;
    // the four 2-D integrals and J00, J01, J11 at k_i: computed there, or interpolated (8-point Lagrange in ln k) from the
    // computed points
    double a[4] = {0.0, 0.0, 0.0, 0.0}, j[3] = {0.0, 0.0, 0.0};
    const double* base = op1 + (size_t)c*nsub*8;
    const double* jb = lin1 + (size_t)c*lin_stride;
#pragma unroll
    for (int l = 0; l < 8; l++) {
        const double w = mw[(size_t)i*8 + l];
        if (w != 0.0) {
            const double* r = base + (size_t)midx[(size_t)i*8 + l]*8;
#pragma unroll
            for (int t = 0; t < 4; t++) a[t] = fma(w, r[t], a[t]);
        }
        const double wj = jw[(size_t)i*8 + l];
        if (wj != 0.0) {
            const int m = jidx[(size_t)i*8 + l];
#pragma unroll
            for (int t = 0; t < 3; t++) j[t] = fma(wj, jb[(size_t)t*nj + m], j[t]);
        }
    }
    const double k = k1[i], P = plk[(size_t)c*ONELOOP_N + i];
    double* o = ol + (size_t)c*4*ONELOOP_N;
    // OneLoopPS.cpp:56-132: 2 I_ab + 6 k^2 J_ab P_L ; P22bar = I23 + 2/3 I32 + 1/5 I33 (:164-185)
    o[i]               = 2.0*a[0] + 6.0*k*k*j[0]*P;
    o[ONELOOP_N + i]   = 2.0*a[1] + 6.0*k*k*j[1]*P;
    o[2*ONELOOP_N + i] = 2.0*a[2] + 6.0*k*k*j[2]*P;
    o[3*ONELOOP_N + i] = a[3];         // I23 + 2/3 I32 + 1/5 I33, combined in the operator (PTPlan ctor)
}

__global__ void k_pack_ik(int nk, int nsub, const double* __restrict__ s /*[c][nsub][32]*/, const int* __restrict__ midx /*[nk][8]*/,
                          const double* __restrict__ mw /*[nk][8]*/, double* __restrict__ I, double* __restrict__ K)
{
    const int i = blockIdx.x*blockDim.x + threadIdx.x, c = blockIdx.y;
    if (i >= nk) return;
    // computed at k_i, or interpolated from the computed points (PTPlan ctor)
    double v[29];
#pragma unroll
    for (int j = 0; j < 29; j++) v[j] = 0.0;
#pragma unroll 1
    for (int l = 0; l < 8; l++) {
        const double w = mw[(size_t)i*8 + l];
        if (w == 0.0) continue;
        const double* r = s + ((size_t)c*nsub + midx[(size_t)i*8 + l])*32;
#pragma unroll
        for (int j = 0; j < 29; j++) v[j] = fma(w, r[j], v[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; j++) I[((size_t)c*16 + j)*nk + i] = v[j];
#pragma unroll
    for (int j = 0; j < 13; j++) K[((size_t)c*13 + j)*nk + i] = v[16 + j];
}

// ImnOneLoop parts on all of k_interp from the computed subset: ML [c][21][nk] <- s [c][21][nsub]
__global__ void k_fill_ml(int nk, int nsub, const double* __restrict__ s, const int* __restrict__ midx, const double* __restrict__ mw,
                          double* __restrict__ ML)
{
    const int i = blockIdx.x*blockDim.x + threadIdx.x, row = blockIdx.y;       // row = c*21 + t
    if (i >= nk) return;
    const double* r = s + (size_t)row*nsub;
    double v = 0.0;
#pragma unroll
    for (int l = 0; l < 8; l++) {
        const double w = mw[(size_t)i*8 + l];
        if (w == 0.0) continue;
        const double x = r[midx[(size_t)i*8 + l]];
        v = fma(w, x, v);
    }
    ML[(size_t)row*nk + i] = v;
}

__global__ void k_pack_lin(int nk, int nsub, const int* __restrict__ midx /*[nk][8]*/, const double* __restrict__ mw /*[nk][8]*/,
                           int nout, int skip, int with_zel, const double* __restrict__ s /*[c][nout], the first `skip` belong to the one-loop stage*/,
                           const double* __restrict__ q3, const double* __restrict__ kurt,
                           double* __restrict__ J, double* __restrict__ sigk, double* __restrict__ scal,
                           double* __restrict__ X0, double* __restrict__ YY)
{
    const int i = blockIdx.x*blockDim.x + threadIdx.x, c = blockIdx.y;
    const double* r = s + (size_t)c*nout + skip;
    if (i < 6*nk) {
        // computed at k_i, or interpolated from the computed points (PTPlan ctor)
        const int j = i/nk, ik = i - j*nk;
        const double* rj = r + (size_t)j*nsub;
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < 8; l++) {
            const double w = mw[(size_t)ik*8 + l];
            if (w != 0.0) v = fma(w, rj[midx[(size_t)ik*8 + l]], v);
        }
        J[(size_t)c*6*nk + i] = v;
    }
    r += 6*nsub;
    if (i < nk) sigk[(size_t)c*nk + i] = r[i];
    r += nk;
    if (i == 0) {
        scal[c*4 + 0] = r[0];
        scal[c*4 + 1] = kurt ? kurt[c] : 0.0;
        scal[c*4 + 2] = q3[c];
        scal[c*4 + 3] = 0.0;
    }
    if (with_zel && i < ZEL_NR) {
        X0[(size_t)c*ZEL_NR + i] = r[1 + i];
        YY[(size_t)c*ZEL_NR + i] = r[1 + ZEL_NR + i];
    }
}

__global__ void k_plin_at(int nspl, const PlinParams* __restrict__ params, const double* __restrict__ spl,
                          int nq, const double* __restrict__ q, double* __restrict__ out)
{
    const int c = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= nq) return;
    out[(size_t)c*nq + j] = plin_eval(params[c], spl + (size_t)c*5*nspl, q[j]);
}

void PTPlan::run_oneloop(SpectraBatch& sp, PTResult& res, bool sync)
{
    Context& c = *ctx;
    cudaStream_t s = c.stream;
    const int nc = sp.ncosmo;
    res.ensure(nc, nk);
    const int ntile = (nc + 7)/8;

    // ---- tile maps: the table layout of a batch depends on its size only (SpectraBatch::table_index), so the index
    // arrays are built and uploaded when the batch size changes, not on every run
    if (idx_nc_a != nc) {
        std::vector<int> slot(ntile*8), rowI(ntile*8), tabL(nc);
        for (int t = 0; t < ntile; t++)
            for (int r = 0; r < 8; r++) {
                slot[t*8 + r] = sp.table_index(std::min(t*8 + r, nc - 1), SP_LIN);
                rowI[t*8 + r] = r;
            }
        for (int i = 0; i < nc; i++) tabL[i] = sp.table_index(i, SP_LIN);
        s_slot.upload(slot, s);
        s_rowA.upload(rowI, s);
        d_tabL.upload(tabL, s);
        // stage A: tiles of nodal_diag_rows() cosmologies
        const int R = nodal_diag_rows(), nt = (nc + R - 1)/R;
        std::vector<int> dslot((size_t)nt*R);
        for (int i = 0; i < nt*R; i++) dslot[i] = sp.table_index(std::min(i, nc - 1), SP_LIN);
        s_rowB.upload(dslot, s);
        idx_nc_a = nc;
    }

    // ---- stage A: one-loop tables
    s_op1.ensure((size_t)(nc + 8)*n_1loop_sub*8);
    nodal_apply_diag(c, op_1loop, sp.tables.p, s_rowB.p, nc, s_op1.p);
    s_lin.ensure((size_t)nc*lin_all.nout);
    lin_all.apply(c, sp.tables.p, d_tabL.p, nc, s_lin.p);
    s_plk.ensure((size_t)nc*ONELOOP_N);
    {
        dim3 g((ONELOOP_N + 255)/256, nc);
        k_plin_at<<<g, 256, 0, s>>>(sp.nspl, sp.params.p, sp.spl.p, ONELOOP_N, d_k1loop.p, s_plk.p);
        k_oneloop_combine<<<g, 256, 0, s>>>(nc, n_1loop_sub, s_op1.p, d_olmap_idx.p, d_olmap_w.p, (int)k_1loop_j.size(),
                                            d_jmap_idx.p, d_jmap_w.p, s_lin.p, lin_all.nout, s_plk.p, d_k1loop.p,
                                            res.oneloop.p);
        PRSD_CUDA(cudaGetLastError());
        c.launches += 2;
    }
    sp.set_oneloop_tables(res.oneloop.p);
    if (sync) c.sync();
}

void PTPlan::run(SpectraBatch& sp, PTResult& res, bool sync)
{
    PRSD_REQUIRE(!only_oneloop, "this plan only holds the one-loop stage");
    PRSD_REQUIRE(sp.ctx == ctx, "PTPlan::run: the plan and the spectra batch belong to different contexts");
    res.src_uid = 0;                      // invalid until this run has been enqueued completely
    run_oneloop(sp, res, false);          // stage B is enqueued behind stage A without a host round trip
    Context& c = *ctx;
    cudaStream_t s = c.stream;
    const int nc = sp.ncosmo;
    const int ntile = (nc + 7)/8;
    if (idx_nc_b != nc) {
        std::vector<int> tabSq(nc), tab22(nc), mslot(nc*4);
        for (int i = 0; i < nc; i++) {
            tabSq[i] = sp.table_index(i, SP_LIN_SQ);
            tab22[i] = sp.table_index(i, SP_P22BAR);
            const int kinds[4] = {SP_LIN, SP_DD1, SP_DV1, SP_VV1};
            for (int r = 0; r < 4; r++) mslot[i*4 + r] = sp.table_index(i, kinds[r]);
        }
        d_tabSq.upload(tabSq, s);
        d_tab22.upload(tab22, s);
        s_tabidx.upload(mslot, s);
        idx_nc_b = nc;
    }

    // ---- stage B: I_mn, K_mn
    s_ik.ensure((size_t)ntile*8*nk_sub*32);
    op_ik.apply(c, sp.tables.p, s_slot.p, s_rowA.p, s_rowA.p, ntile, s_ik.p);
    {
        dim3 g((nk + 127)/128, nc);
        k_pack_ik<<<g, 128, 0, s>>>(nk, nk_sub, s_ik.p, d_kmap_idx.p, d_kmap_w.p, res.I.p, res.K.p);
        PRSD_CUDA(cudaGetLastError());
        c.launches++;
    }
    // ---- stage B: ImnOneLoop (one tile = one cosmology, rows = spectrum pairs)
    if (with_ml) {
        if (nk_sub == nk) nodal_apply_ml(c, op_ml, sp.tables.p, s_tabidx.p, nc, res.ML.p);
        else {
            s_ml.ensure((size_t)nc*N_ML*3*nk_sub);
            nodal_apply_ml(c, op_ml, sp.tables.p, s_tabidx.p, nc, s_ml.p);
            dim3 g((nk + 127)/128, nc*N_ML*3);
            k_fill_ml<<<g, 128, 0, s>>>(nk, nk_sub, s_ml.p, d_kmap_idx.p, d_kmap_w.p, res.ML.p);
            PRSD_CUDA(cudaGetLastError());
            c.launches++;
        }
    }
    // ---- 1-D functionals
    // (the 1-D functionals of this batch were computed with the one-loop stage: lin_all)
    s_small.ensure((size_t)2*nc);
    lin_invq2_100.apply(c, sp.tables.p, d_tabSq.p, nc, s_small.p);
    lin_kurt.apply(c, sp.tables.p, d_tab22.p, nc, s_small.p + nc);
    {
        const int nmax = std::max(6*nk, ZEL_NR);
        dim3 g((nmax + 127)/128, nc);
        k_pack_lin<<<g, 128, 0, s>>>(nk, nk_sub, d_kmap_idx.p, d_kmap_w.p, lin_all.nout, n_lin1, with_zel ? 1 : 0, s_lin.p, s_small.p, s_small.p + nc,
                                    res.J.p, res.sigmasq_k.p, res.scalars.p, res.X0.p, res.YY.p);
        PRSD_CUDA(cudaGetLastError());
        c.launches++;
    }
    res.src_uid = sp.uid;
    res.src_gen = sp.lin_gen;
    if (sync) c.sync();
}

} // namespace prsd
