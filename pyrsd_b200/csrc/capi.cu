// capi.cu -- extern "C" entry points declared in include/pyrsd_b200.h.
#include "../../include/pyrsd_b200.h"
#include "plan.cuh"
#include "galaxy.cuh"
#include "xilm.cuh"
#include "discrete.cuh"
#include "transfer.cuh"
#include "exchange.cuh"
#include "biasrel.cuh"
#include "gp.cuh"

#include <cstring>
#include <list>
#include <mutex>

using namespace prsd;
namespace prsd { double fp64_peak_tflops(Context& ctx, int mode); }

struct prsd_ctx { Context c; explicit prsd_ctx(int d) : c(d) {} };
struct prsd_spectra { SpectraBatch s; prsd_spectra(Context& c, int nc, int ns, const double* k, const double* T, const double* p) : s(c, nc, ns, k, T, p) {} };
struct prsd_plan { PTPlan p; prsd_plan(Context& c, const double* k, int nk, const QuadConfig& q, bool ml, bool zel) : p(c, k, nk, q, ml, zel) {} };
struct prsd_result { PTResult r; Context* ctx = nullptr; };
struct prsd_galaxy {
    GalaxyModel g;
    DevBuf<double> dk, dmu, dout;
    int async_npts = 0;             // evaluation points resident in dk / dmu from the last asynchronous call that passed them
    prsd_galaxy(Context& c, const GalaxyConfig& cfg, const double* km, int nkm, const double* ki, int nki) : g(c, cfg, km, nkm, ki, nki) {}
};
struct prsd_gridop { GridOp op; DevBuf<double> din, dout; prsd_gridop(Context& c, int nb, int nk, int nmu, const double* w) : op(c, nb, nk, nmu, w) {} };
struct prsd_exchange { Exchange x; prsd_exchange(Context& c, int world, int rank, size_t n) : x(c, world, rank, n) {} };
struct prsd_window {
    WindowOp op;
    DevBuf<double> din, dout;
    prsd_window(Context& c, int nk, const double* k, int nell, const int* ells, double q, int N) : op(c, nk, k, nell, ells, q, N) {}
};
struct prsd_gp { GPModel m; prsd_gp(Context& c, int n, int nd, const double* x, const double* xm, const double* xs, const double* met, double amp, const double* al, double ym, double ys) : m(c, n, nd, x, xm, xs, met, amp, al, ym, ys) {} };
struct prsd_zelplan { ZelPlan z; Context* ctx = nullptr; };

static thread_local std::string g_err;

template <typename F>
static int guarded(F&& f)
{
    try {
        f();
        return PRSD_OK;
    } catch (const Error& e) {
        g_err = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_err = e.what();
        return PRSD_ERR_INTERNAL;
    } catch (...) {
        g_err = "unknown exception";
        return PRSD_ERR_INTERNAL;
    }
}

// Every entry point that takes a context-owning handle runs with that context's device current (the current device
// is per-thread state: another Context, torch, or a call from a different host thread may have changed it) and
// restores the caller's device on exit.
struct DeviceGuard {
    int prev = -1, want = -1;
    cudaStream_t prev_stream = nullptr;
    explicit DeviceGuard(const Context* c)
    {
        if (!c) return;
        want = c->device;
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != want) PRSD_CUDA(cudaSetDevice(want));
        prev_stream = pool_set_stream(c->stream);       // temporaries of this call are stream-ordered on the context's stream
    }
    ~DeviceGuard()
    {
        if (want < 0) return;
        pool_set_stream(prev_stream);
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
};

template <typename F>
static int guarded(const Context* c, F&& f)
{
    return guarded([&] { DeviceGuard dg(c); f(); });
}

static const Context* dev_of(const prsd_ctx* h) { return h ? &h->c : nullptr; }
static const Context* dev_of(const prsd_spectra* h) { return h ? h->s.ctx : nullptr; }
static const Context* dev_of(const prsd_plan* h) { return h ? h->p.ctx : nullptr; }
static const Context* dev_of(const prsd_result* h) { return h ? h->ctx : nullptr; }
static const Context* dev_of(const prsd_galaxy* h) { return h ? h->g.ctx : nullptr; }
static const Context* dev_of(const prsd_gridop* h) { return h ? h->op.ctx : nullptr; }
static const Context* dev_of(const prsd_gp* h) { return h ? h->m.ctx : nullptr; }
static const Context* dev_of(const prsd_window* h) { return h ? h->op.ctx : nullptr; }
static const Context* dev_of(const prsd_exchange* h) { return h ? h->x.ctx : nullptr; }

static QuadConfig cfg_from(const double* c)
{
    QuadConfig q;
    if (c) {
        q.qmin = c[0]; q.ln_far = c[1]; q.ln_mid = c[2]; q.bao_lo = c[3]; q.bao_hi = c[4]; q.bao_dq = c[5];
        q.ln_inner = c[6]; q.ln_thin = c[7];
        const int NT = QuadConfig::NTIER;
        for (int i = 0; i < NT; i++) {
            q.tier[i] = {c[8 + 4*i], (int)c[9 + 4*i], (int)c[10 + 4*i], (int)c[11 + 4*i]};
            q.ml_tier[i] = {c[8 + 4*NT + 4*i], (int)c[9 + 4*NT + 4*i], (int)c[10 + 4*NT + 4*i], (int)c[11 + 4*NT + 4*i]};
        }
        PRSD_REQUIRE(q.qmin >= 1e-8 && q.qmin < 1e-5, "quadrature config: qmin = %g outside [1e-8, 1e-5)", q.qmin);
        for (int i = 0; i < QuadConfig::NTIER; i++)
            for (int n : {q.tier[i].n_outer, q.tier[i].n_inner, q.tier[i].n_thin, q.ml_tier[i].n_outer, q.ml_tier[i].n_inner, q.ml_tier[i].n_thin})
                PRSD_REQUIRE(n >= 2 && n <= 64, "quadrature config: bad point count %d", n);
        PRSD_REQUIRE(q.ln_far > 0 && q.ln_mid > 0 && q.bao_dq > 0 && q.ln_inner > 0 && q.ln_thin >= 0, "quadrature config: widths must be positive");
    }
    return q;
}

// small cache of node geometries for the one-off calls (keyed by device, qmax, config, k)
struct GeoKey {
    int device; double qmax; QuadConfig cfg; std::vector<double> k;
    bool operator==(const GeoKey& o) const
    {
        return device == o.device && qmax == o.qmax && k == o.k && cfg == o.cfg;
    }
};
static std::mutex g_geo_mu;
static std::list<std::pair<GeoKey, std::shared_ptr<NodeGeometry>>> g_geo_cache;

static std::shared_ptr<NodeGeometry> cached_geometry(Context& c, const double* k, int nk, double qmax, const QuadConfig& cfg)
{
    GeoKey key{c.device, qmax, cfg, std::vector<double>(k, k + nk)};
    std::lock_guard<std::mutex> lock(g_geo_mu);
    for (auto it = g_geo_cache.begin(); it != g_geo_cache.end(); ++it)
        if (it->first == key) {
            g_geo_cache.splice(g_geo_cache.begin(), g_geo_cache, it);
            return g_geo_cache.front().second;
        }
    auto g = make_geometry(c, k, nk, qmax, cfg, knot_grid(), 4);      // consumed by the DMMA kernel
    g_geo_cache.emplace_front(std::move(key), g);
    while (g_geo_cache.size() > 6) g_geo_cache.pop_back();
    return g;
}

// apply a small operator with rows = (cosmology) and tables (kindA, kindB); returns [ncosmo][nk][ncol]
static void apply_rows(SpectraBatch& sp, const BilinearOp& op, int kindA, int kindB, std::vector<double>& host)
{
    Context& c = *sp.ctx;
    const int nc = sp.ncosmo, ntile = (nc + 3)/4;
    // 4 cosmologies per tile: slots 0..3 = kindA tables, 4..7 = kindB tables
    std::vector<int> slot(ntile*8), rA(ntile*8), rB(ntile*8);
    for (int t = 0; t < ntile; t++)
        for (int r = 0; r < 8; r++) {
            const int ci = std::min(t*4 + (r & 3), nc - 1);
            slot[t*8 + r] = sp.table_index(ci, r < 4 ? kindA : kindB);
            rA[t*8 + r] = r & 3;
            rB[t*8 + r] = 4 + (r & 3);
        }
    DevBuf<int> ds, dA, dB;
    ds.upload(slot, c.stream);
    dA.upload(rA, c.stream);
    dB.upload(rB, c.stream);
    const int nk = op.geo->nk;
    DevBuf<double> out((size_t)ntile*8*nk*op.ncol);
    op.apply(c, sp.tables.p, ds.p, dA.p, dB.p, ntile, out.p);
    std::vector<double> tmp((size_t)ntile*8*nk*op.ncol);
    out.download(tmp.data(), tmp.size(), c.stream);
    c.sync();
    // rows 0..3 of each tile are the cosmologies (rows 4..7 duplicate them)
    host.assign((size_t)nc*nk*op.ncol, 0.0);
    for (int i = 0; i < nc; i++) {
        const int t = i/4, r = i%4;
        std::memcpy(&host[(size_t)i*nk*op.ncol], &tmp[((size_t)(t*8 + r))*nk*op.ncol], sizeof(double)*nk*op.ncol);
    }
}

static void eval_kernel_generic(SpectraBatch& sp, int kid, int convention, int kindQ, int kindR, double qmax,
                                int nk, const double* k, const double* cfg9, double* out)
{
    // convention 0: Imn, 1: Kmn, 2: ImnOneLoop (see Imn.cpp:120-140, Kmn.cpp:118-138, ImnOneLoop.cpp:82-93)
    Context& c = *sp.ctx;
    if (nk <= 0) return;
    const QuadConfig base = cfg_from(cfg9);
    const QuadConfig cfg = convention == 2 ? base.ml_variant() : base;
    // the reference returns 0 for k <= 0 (Imn.cpp:122)
    std::vector<double> kpos;
    std::vector<int> where;
    for (int i = 0; i < nk; i++) {
        PRSD_REQUIRE(std::isfinite(k[i]), "k[%d] is not finite", i);
        if (k[i] > 0) { kpos.push_back(k[i]); where.push_back(i); }
    }
    for (size_t i = 0; i < (size_t)sp.ncosmo*nk; i++) out[i] = 0.0;
    if (kpos.empty()) return;
    auto geo = cached_geometry(c, kpos.data(), (int)kpos.size(), qmax, cfg);
    const bool sym = kernel_is_symmetric(kid);
    const double V8 = 1.0/(8.0*M_PI*M_PI), V4 = 1.0/(4.0*M_PI*M_PI);
    const int np = (int)kpos.size();
    std::vector<double> pos, neg;
    BilinearOp op;
    if (convention == 0)      op.build(c, geo, {{kid, +1, sym ? 2*V8 : V8}});
    else if (convention == 1) op.build(c, geo, {{kid, +1, sym ? V4 : 0.5*V4}});
    else                      op.build(c, geo, {{kid, +1, V8}});
    apply_rows(sp, op, kindQ, kindR, pos);
    const bool need_neg = (convention == 2) || !sym;
    if (need_neg) {
        BilinearOp on;
        on.build(c, geo, {{kid, -1, convention == 1 ? 0.5*V4 : V8}});
        apply_rows(sp, on, kindR, kindQ, neg);      // a = r, b = q on the v < 0 half
    }
    for (int ci = 0; ci < sp.ncosmo; ci++)
        for (int i = 0; i < np; i++) {
            double v = pos[((size_t)ci*np + i)*8];
            if (need_neg) v += neg[((size_t)ci*np + i)*8];
            out[(size_t)ci*nk + where[i]] = v;
        }
}

// The 1-D operators are cosmology independent and some are expensive to build (X_Zel / Y_Zel integrate j0, j1(q r)
// against every table basis function: ~1.3 s for the 1024-point Zel'dovich r grid), so the one-off entry points keep
// the most recent ones per device.
static std::shared_ptr<LinearOp> cached_linear_op(Context& c, const std::vector<LinOutSpec>& outs)
{
    struct Entry { int dev; std::vector<LinOutSpec> outs; std::shared_ptr<LinearOp> op; };
    static std::mutex mu;
    static std::list<Entry> cache;
    auto same = [](const std::vector<LinOutSpec>& a, const std::vector<LinOutSpec>& b) {
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i < a.size(); i++)
            if (a[i].type != b[i].type || a[i].sub != b[i].sub || a[i].param != b[i].param || a[i].lo != b[i].lo ||
                a[i].hi != b[i].hi || a[i].scale != b[i].scale) return false;
        return true;
    };
    {
        std::lock_guard<std::mutex> lock(mu);
        for (auto it = cache.begin(); it != cache.end(); ++it)
            if (it->dev == c.device && same(it->outs, outs)) {
                cache.splice(cache.begin(), cache, it);
                return cache.front().op;
            }
    }
    auto op = std::make_shared<LinearOp>();
    op->build(c, outs);
    std::lock_guard<std::mutex> lock(mu);
    cache.push_front({c.device, outs, op});
    while (cache.size() > 12) cache.pop_back();
    return op;
}

static void eval_linear_generic(SpectraBatch& sp, const std::vector<LinOutSpec>& outs, int kind, double* out)
{
    Context& c = *sp.ctx;
    std::shared_ptr<LinearOp> opp = cached_linear_op(c, outs);
    LinearOp& op = *opp;
    std::vector<int> idx(sp.ncosmo);
    for (int i = 0; i < sp.ncosmo; i++) idx[i] = sp.table_index(i, kind);
    DevBuf<int> di;
    di.upload(idx, c.stream);
    DevBuf<double> o((size_t)sp.ncosmo*op.nout);
    op.apply(c, sp.tables.p, di.p, sp.ncosmo, o.p);
    o.download(out, (size_t)sp.ncosmo*op.nout, c.stream);
    c.sync();
}

extern "C" {

const char* prsd_last_error(void) { return g_err.c_str(); }
int prsd_version(void) { return 100; }

int prsd_ctx_create(int device, prsd_ctx** out)
{
    return guarded([&] { PRSD_REQUIRE(out, "null output"); *out = new prsd_ctx(device); });
}
int prsd_ctx_destroy(prsd_ctx* ctx) { return guarded(dev_of(ctx), [&] { delete ctx; }); }
int prsd_ctx_sync(prsd_ctx* ctx) { return guarded(dev_of(ctx), [&] { PRSD_REQUIRE(ctx, "null context"); ctx->c.sync(); }); }
int prsd_ctx_launches(prsd_ctx* ctx, long long* n)
{
    return guarded(dev_of(ctx), [&] { PRSD_REQUIRE(ctx && n, "null argument"); *n = ctx->c.launches; });
}
int prsd_ctx_stream(prsd_ctx* ctx, void** s)
{
    return guarded(dev_of(ctx), [&] { PRSD_REQUIRE(ctx && s, "null argument"); *s = (void*)ctx->c.stream; });
}

int prsd_knot_grid(int* n, double* lnq)
{
    return guarded([&] {
        const KnotGrid& g = knot_grid();
        if (n) *n = g.size();
        if (lnq) std::memcpy(lnq, g.x.data(), sizeof(double)*g.size());
    });
}

int prsd_quad_config_default(double* c9)
{
    return guarded([&] {
        PRSD_REQUIRE(c9, "null argument");
        QuadConfig q;
        double* c = c9;
        c[0] = q.qmin; c[1] = q.ln_far; c[2] = q.ln_mid; c[3] = q.bao_lo; c[4] = q.bao_hi; c[5] = q.bao_dq;
        c[6] = q.ln_inner; c[7] = q.ln_thin;
        const int NT = QuadConfig::NTIER;
        for (int i = 0; i < NT; i++) {
            c[8 + 4*i] = q.tier[i].k_below; c[9 + 4*i] = q.tier[i].n_outer; c[10 + 4*i] = q.tier[i].n_inner; c[11 + 4*i] = q.tier[i].n_thin;
            const int o = 8 + 4*NT + 4*i;
            c[o] = q.ml_tier[i].k_below; c[o + 1] = q.ml_tier[i].n_outer; c[o + 2] = q.ml_tier[i].n_inner; c[o + 3] = q.ml_tier[i].n_thin;
        }
    });
}

int prsd_spectra_create(prsd_ctx* ctx, int ncosmo, int nspl, const double* k, const double* T,
                        const double* pars, prsd_spectra** out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && k && T && pars && out, "null argument");
        *out = new prsd_spectra(ctx->c, ncosmo, nspl, k, T, pars);
    });
}
int prsd_spectra_destroy(prsd_spectra* sp) { return guarded(dev_of(sp), [&] { delete sp; }); }
int prsd_spectra_rescale(prsd_spectra* sp, const double* fac)
{
    return guarded(dev_of(sp), [&] { PRSD_REQUIRE(sp && fac, "null argument"); sp->s.rescale_linear(fac); });
}
int prsd_spectra_plin(prsd_spectra* sp, int nq, const double* q, double* out)
{
    return guarded(dev_of(sp), [&] { PRSD_REQUIRE(sp && (nq == 0 || (q && out)), "null argument"); sp->s.eval_linear(nq, q, out); });
}
int prsd_spectra_table(prsd_spectra* sp, int kind, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && out, "null argument");
        PRSD_REQUIRE(kind >= 0 && kind < SP_NKIND, "invalid table kind %d", kind);
        SpectraBatch& s = sp->s;
        for (int i = 0; i < s.ncosmo; i++)
            PRSD_CUDA(cudaMemcpyAsync(out + (size_t)i*TAB_M, s.table_ptr(i, kind), TAB_M*sizeof(double), cudaMemcpyDeviceToHost, s.ctx->stream));
        s.ctx->sync();
    });
}
int prsd_spectra_set_oneloop(prsd_spectra* sp, const double* tables)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && tables, "null argument");
        SpectraBatch& s = sp->s;
        DevBuf<double> d;
        d.upload(tables, (size_t)s.ncosmo*4*ONELOOP_N, s.ctx->stream);
        s.set_oneloop_tables(d.p);
        s.ctx->sync();
    });
}
int prsd_spectra_oneloop_eval(prsd_spectra* sp, int which, int full, int nq, const double* q, double* out)
{
    return guarded(dev_of(sp), [&] { PRSD_REQUIRE(sp && (nq == 0 || (q && out)), "null argument"); sp->s.eval_oneloop(which, full != 0, nq, q, out); });
}

int prsd_eval_imn(prsd_spectra* sp, int m, int n, int nk, const double* k, const double* cfg9, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && (nk == 0 || (k && out)), "null argument");
        // the reference aborts on invalid indices (Imn.cpp:186); here: error code
        PRSD_REQUIRE(m >= 0 && m <= 3 && n >= 0 && n <= 3, "Imn: invalid indices, m = %d, n = %d", m, n);
        eval_kernel_generic(sp->s, 4*m + n, 0, SP_LIN, SP_LIN, 1e5, nk, k, cfg9, out);
    });
}

int prsd_eval_kmn(prsd_spectra* sp, int m, int n, int tidal, int part, int nk, const double* k, const double* cfg9, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && (nk == 0 || (k && out)), "null argument");
        int kid = -1;
        // Kmn::Evaluate dispatch on 2 (3 m + tidal) + n (Kmn.cpp:151-183)
        switch (2*(3*m + (tidal ? 1 : 0)) + n) {
            case 0: kid = KID_K00; break;   case 1: kid = KID_K01; break;
            case 2: kid = KID_K00S; break;  case 3: kid = KID_K01S; break;
            case 4: kid = KID_K02S; break;  case 6: kid = KID_K10; break;
            case 7: kid = KID_K11; break;   case 8: kid = KID_K10S; break;
            case 9: kid = KID_K11S; break;
            case 12: kid = part == 0 ? KID_K20A : KID_K20B; break;
            case 14: kid = part == 0 ? KID_K20SA : KID_K20SB; break;
            default: break;
        }
        PRSD_REQUIRE(kid >= 0 && m >= 0 && n >= 0, "Kmn: invalid indices, m = %d, n = %d", m, n);
        eval_kernel_generic(sp->s, kid, 1, SP_LIN, SP_LIN, 1e5, nk, k, cfg9, out);
    });
}

int prsd_eval_jmn(prsd_spectra* sp, int m, int n, int nk, const double* k, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && (nk == 0 || (k && out)), "null argument");
        const int id = 3*m + n;
        PRSD_REQUIRE(m >= 0 && n >= 0 && (id == 0 || id == 1 || id == 2 || id == 3 || id == 4 || id == 6) && n <= 2,
                     "Jmn: invalid indices, m = %d, n = %d", m, n);
        if (nk == 0) return;
        std::vector<LinOutSpec> o;
        std::vector<int> where;
        for (int i = 0; i < nk; i++) if (k[i] > 0) { o.push_back({LK_JMN, id, k[i], 1e-5, 1e5, 1.0}); where.push_back(i); }
        for (size_t i = 0; i < (size_t)sp->s.ncosmo*nk; i++) out[i] = 0.0;
        if (o.empty()) return;
        std::vector<double> tmp((size_t)sp->s.ncosmo*o.size());
        eval_linear_generic(sp->s, o, SP_LIN, tmp.data());
        for (int c = 0; c < sp->s.ncosmo; c++)
            for (size_t i = 0; i < o.size(); i++) out[(size_t)c*nk + where[i]] = tmp[(size_t)c*o.size() + i];
    });
}

int prsd_eval_imn_oneloop(prsd_spectra* sp, int p1, int p2, int which, int m, int n, int nk, const double* k,
                          const double* cfg9, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && (nk == 0 || (k && out)), "null argument");
        PRSD_REQUIRE(sp->s.have_oneloop, "ImnOneLoop: one-loop spectra have not been set for this batch");
        PRSD_REQUIRE(p1 >= 0 && p1 <= 2 && p2 <= 2, "ImnOneLoop: invalid spectra (%d, %d)", p1, p2);
        PRSD_REQUIRE(which >= 0 && which <= 2, "ImnOneLoop: invalid part %d", which);
        const bool equal = p2 < 0;
        if (equal) p2 = p1;
        int kid = -1, kid_cross = -1;
        switch (2*m + n) {          // ImnOneLoop.cpp:97-116
            case 1: kid = kid_cross = KID_H01; break;
            case 2: kid = kid_cross = KID_H02; break;
            case 3: kid = kid_cross = KID_K20A; break;
            case 4: kid = kid_cross = KID_K20B; break;
            case 7: kid = kid_cross = KID_F23; break;
            case 8: kid = KID_F23; kid_cross = KID_F32; break;     // reference quirk (:108-109, :160-165, :199-200)
            case 9: kid = kid_cross = KID_F33; break;
            default: break;
        }
        PRSD_REQUIRE(kid >= 0 && m >= 0 && n >= 0, "ImnOneLoop: invalid indices, m = %d, n = %d", m, n);
        const int K1 = SP_DD1 + p1, K2 = SP_DD1 + p2;
        SpectraBatch& s = sp->s;
        if (which == 0) eval_kernel_generic(s, kid, 2, SP_LIN, SP_LIN, 10.0, nk, k, cfg9, out);
        else if (which == 2) eval_kernel_generic(s, kid, 2, K1, K2, 10.0, nk, k, cfg9, out);
        else {
            eval_kernel_generic(s, kid_cross, 2, SP_LIN, K2, 10.0, nk, k, cfg9, out);
            const size_t tot = (size_t)s.ncosmo*nk;
            if (equal) { for (size_t i = 0; i < tot; i++) out[i] *= 2.0; }
            else {
                std::vector<double> part2(tot);
                eval_kernel_generic(s, kid_cross, 2, K1, SP_LIN, 10.0, nk, k, cfg9, part2.data());
                for (size_t i = 0; i < tot; i++) out[i] += part2[i];
            }
        }
    });
}

int prsd_eval_ps_integral(prsd_spectra* sp, int what, int n, const double* x, double factor, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && out, "null argument");
        std::vector<LinOutSpec> o;
        int kind = SP_LIN;
        if (what == 0) { n = 1; o.push_back({LK_SIGV, 0, 0.0, 1e-5, 1e2, 1.0}); }
        else {
            PRSD_REQUIRE(n >= 0 && (n == 0 || x), "null argument");
            if (n == 0) return;
            for (int i = 0; i < n; i++) {
                const double v = x[i];
                PRSD_REQUIRE(std::isfinite(v) && v > 0, "argument %d = %g must be positive", i, v);
                switch (what) {
                    case 1: {
                        const double hi = factor*v;
                        if (hi >= 1e-5) o.push_back({LK_SIGV, 0, 0.0, 1e-5, hi, 1.0});
                        else o.push_back({LK_SIGV, 0, 0.0, hi, 1e-5, -1.0});
                        break;
                    }
                    case 2: o.push_back({LK_XZEL, 0, v, 1e-5, 1e2, 1.0}); break;
                    case 3: o.push_back({LK_YZEL, 0, v, 1e-5, 1e2, 1.0}); break;
                    case 4: o.push_back({LK_INVQ2, 0, 0.0, 1e-5, 1e2, v*v*v*v/(10.0*M_PI*M_PI)}); kind = SP_LIN_SQ; break;
                    case 5: o.push_back({LK_SIGMA_R, 0, v, 1e-5, 1e2, 1.0}); break;
                    case 6: o.push_back({LK_SIGMA3, 0, v, 1e-5*v, 1e2*v, 1.0}); break;
                    default: PRSD_REQUIRE(false, "unknown integral %d", what);
                }
            }
        }
        eval_linear_generic(sp->s, o, kind, out);
        if (what == 5) for (size_t i = 0; i < (size_t)sp->s.ncosmo*n; i++) out[i] = std::sqrt(out[i]);
    });
}

int prsd_eval_correlation(prsd_spectra* sp, int n, const double* r, double kmin, double kmax, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && out && (n == 0 || r), "null argument");
        PRSD_REQUIRE(kmin > 0 && kmax > kmin, "CorrelationFunction: need 0 < kmin < kmax (got %g, %g)", kmin, kmax);
        if (n <= 0) return;
        std::vector<LinOutSpec> o;
        for (int i = 0; i < n; i++) {
            PRSD_REQUIRE(std::isfinite(r[i]) && r[i] > 0, "r[%d] = %g must be positive", i, r[i]);
            o.push_back({LK_XI, 0, r[i], kmin, kmax, 1.0});
        }
        eval_linear_generic(sp->s, o, SP_LIN, out);
    });
}

int prsd_eval_kurtosis(prsd_spectra* sp, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp && out, "null argument");
        PRSD_REQUIRE(sp->s.have_oneloop, "VelocityKurtosis: one-loop spectra have not been set for this batch");
        eval_linear_generic(sp->s, {{LK_INVQ2, 0, 0.0, 1e-5, 1e1, 0.25/(2.0*M_PI*M_PI)}}, SP_P22BAR, out);
    });
}

int prsd_plan_create(prsd_ctx* ctx, int nk, const double* k, const double* cfg9, int flags, prsd_plan** out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && k && out, "null argument");
        *out = new prsd_plan(ctx->c, k, nk, cfg_from(cfg9), (flags & 1) != 0, (flags & 2) != 0);
    });
}
int prsd_plan_destroy(prsd_plan* plan) { return guarded(dev_of(plan), [&] { delete plan; }); }
int prsd_plan_info(prsd_plan* plan, double* info)
{
    return guarded(dev_of(plan), [&] {
        PRSD_REQUIRE(plan && info, "null argument");
        const PTPlan& p = plan->p;
        info[0] = (double)p.nodes_total();
        info[1] = (double)p.device_bytes();
        info[2] = p.build_seconds;
        info[3] = p.nk;
        info[4] = TAB_M;
        info[5] = p.op_1loop.geo ? (double)p.op_1loop.geo->n_nodes : 0.0;
        info[6] = p.op_ik.geo ? (double)p.op_ik.geo->n_nodes : 0.0;
        info[7] = p.op_ml.geo ? (double)p.op_ml.geo->n_nodes : 0.0;
    });
}
static void plan_run(prsd_plan* plan, prsd_spectra* sp, prsd_result** res, bool sync);

int prsd_plan_build_profile(prsd_plan* plan, double* ms8)
{
    return guarded(dev_of(plan), [&] {
        PRSD_REQUIRE(plan && ms8, "null argument");
        for (int i = 0; i < 8; i++) ms8[i] = plan->p.build_stage_ms[i];
    });
}

int prsd_plan_run_async(prsd_plan* plan, prsd_spectra* sp, prsd_result** res)
{
    return guarded(dev_of(plan), [&] { plan_run(plan, sp, res, false); });
}

int prsd_plan_run(prsd_plan* plan, prsd_spectra* sp, prsd_result** res)
{
    return guarded(dev_of(plan), [&] { plan_run(plan, sp, res, true); });
}

static void plan_run(prsd_plan* plan, prsd_spectra* sp, prsd_result** res, bool sync)
{
    {
        PRSD_REQUIRE(plan && sp && res, "null argument");
        PRSD_REQUIRE(plan->p.ctx == sp->s.ctx, "prsd_plan_run: the plan and the spectra batch belong to different contexts");
        if (!*res) { *res = new prsd_result(); (*res)->ctx = plan->p.ctx; }
        PRSD_REQUIRE((*res)->ctx == plan->p.ctx, "prsd_plan_run: the result belongs to another context");
        plan->p.run(sp->s, (*res)->r, sync);
    }
}
int prsd_result_destroy(prsd_result* res) { return guarded(dev_of(res), [&] { delete res; }); }
int prsd_result_get(prsd_result* res, int what, double* out, long long n_expected)
{
    return guarded(dev_of(res), [&] {
        PRSD_REQUIRE(res && out, "null argument");
        PTResult& r = res->r;
        const size_t c = r.ncosmo, nk = r.nk;
        const DevBuf<double>* b = nullptr;
        size_t n = 0;
        switch (what) {
            case 0: b = &r.I; n = c*16*nk; break;
            case 1: b = &r.J; n = c*6*nk; break;
            case 2: b = &r.K; n = c*13*nk; break;
            case 3: b = &r.ML; n = c*N_ML*3*nk; break;
            case 4: b = &r.oneloop; n = c*4*ONELOOP_N; break;
            case 5: b = &r.sigmasq_k; n = c*nk; break;
            case 6: b = &r.scalars; n = c*4; break;
            case 7: b = &r.X0; n = c*ZEL_NR; break;
            case 8: b = &r.YY; n = c*ZEL_NR; break;
            default: PRSD_REQUIRE(false, "unknown result id %d", what);
        }
        PRSD_REQUIRE((long long)n == n_expected, "result %d has %zu elements, caller expects %lld", what, n, n_expected);
        b->download(out, n, res->ctx->stream);
        res->ctx->sync();
    });
}


/* ------------------------------------------------------------------ FFTLog */
int prsd_fftlog_dir(prsd_ctx* ctx, int batch, int N, double mu, double q, double dlnr, double kr, int kropt, int direction,
                    const double* a, double* out, double* kr_out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && a && out, "null argument");
        PRSD_REQUIRE(batch > 0, "FFTLog: empty batch");
        auto plan = fftlog_plan(ctx->c, N, mu, q, dlnr, kr, kropt, direction);
        DevBuf<double> din, dout((size_t)batch*N);
        din.upload(a, (size_t)batch*N, ctx->c.stream);
        fftlog_apply(ctx->c, *plan, batch, din.p, dout.p);
        dout.download(out, (size_t)batch*N, ctx->c.stream);
        ctx->c.sync();
        if (kr_out) *kr_out = plan->kr;
    });
}

int prsd_fftlog(prsd_ctx* ctx, int batch, int N, double mu, double q, double dlnr, double kr, int kropt,
                const double* a, double* out, double* kr_out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && a && out, "null argument");
        PRSD_REQUIRE(batch > 0, "FFTLog: empty batch");
        auto plan = fftlog_plan(ctx->c, N, mu, q, dlnr, kr, kropt);
        DevBuf<double> din, dout((size_t)batch*N);
        din.upload(a, (size_t)batch*N, ctx->c.stream);
        fftlog_apply(ctx->c, *plan, batch, din.p, dout.p);
        dout.download(out, (size_t)batch*N, ctx->c.stream);
        ctx->c.sync();
        if (kr_out) *kr_out = plan->kr;
    });
}

int prsd_fftlog_device(prsd_ctx* ctx, int batch, int N, double mu, double q, double dlnr, double kr, int kropt,
                       const double* d_in, double* d_out, double* kr_out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && d_in && d_out, "null argument");
        auto plan = fftlog_plan(ctx->c, N, mu, q, dlnr, kr, kropt);
        fftlog_apply(ctx->c, *plan, batch, d_in, d_out);
        if (kr_out) *kr_out = plan->kr;
    });
}

int prsd_compute_xilm(prsd_ctx* ctx, int batch, int l, int m, int nk, const double* k, const double* pk,
                      int nr, const double* r, double smoothing, double scale, double* out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && k && pk && r && out, "null argument");
        PRSD_REQUIRE(batch > 0 && nr > 0, "ComputeXiLM: empty request");
        DevBuf<double> dpk, dout((size_t)batch*nr);
        dpk.upload(pk, (size_t)batch*nk, ctx->c.stream);
        compute_xilm_dev(ctx->c, batch, l, m, nk, k, dpk.p, nr, r, smoothing, scale, dout.p);
        dout.download(out, (size_t)batch*nr, ctx->c.stream);
        ctx->c.sync();
    });
}

int prsd_compute_xilm_discrete(prsd_ctx* ctx, int method, int batch, int l, int m, int nk, const double* k, const double* pk,
                               int nr, const double* r, double smoothing, double scale, double* out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && k && pk && r && out, "null argument");
        PRSD_REQUIRE(batch > 0 && nr > 0, "ComputeXiLM: empty request");
        for (int i = 1; i < nk; i++) PRSD_REQUIRE(k[i] > k[i - 1], "ComputeXiLM: abscissae must increase");
        DevBuf<double> dk, dpk, dr, dout((size_t)batch*nr);
        dk.upload(k, nk, ctx->c.stream);
        dr.upload(r, nr, ctx->c.stream);
        dpk.upload(pk, (size_t)batch*nk, ctx->c.stream);
        xilm_discrete_dev(ctx->c, method, batch, l, m, nk, dk.p, dpk.p, nr, dr.p, smoothing, scale, dout.p);
        dout.download(out, (size_t)batch*nr, ctx->c.stream);
        ctx->c.sync();
    });
}

int prsd_discrete_integrate(prsd_ctx* ctx, int method, int batch, int n, const double* x, const double* y, double* out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && x && y && out, "null argument");
        PRSD_REQUIRE(batch > 0, "discrete integration: empty request");
        DevBuf<double> dx, dy, dout((size_t)batch);
        dx.upload(x, n, ctx->c.stream);
        dy.upload(y, (size_t)batch*n, ctx->c.stream);
        discrete_integrate_dev(ctx->c, method, batch, n, dx.p, dy.p, dout.p);
        dout.download(out, (size_t)batch, ctx->c.stream);
        ctx->c.sync();
    });
}

int prsd_compute_xilm_fftlog(prsd_ctx* ctx, int batch, int l, int m, int N, const double* k, const double* pk,
                             double q, double smoothing, double* r_out, double* xi_out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && k && pk && r_out && xi_out, "null argument");
        PRSD_REQUIRE(batch > 0 && N >= 2, "ComputeXiLM_fftlog: empty request");
        DevBuf<double> dk, dpk, dxi((size_t)batch*N);
        dk.upload(k, N, ctx->c.stream);
        dpk.upload(pk, (size_t)batch*N, ctx->c.stream);
        std::vector<double> r;
        compute_xilm_fftlog_dev(ctx->c, batch, l, m, N, dk.p, k, dpk.p, q, smoothing, r, dxi.p, nullptr);
        std::memcpy(r_out, r.data(), sizeof(double)*N);
        dxi.download(xi_out, (size_t)batch*N, ctx->c.stream);
        ctx->c.sync();
    });
}

/* -------------------------------------------------------------- Zel'dovich */
int prsd_zeldovich(prsd_ctx* ctx, int nrow, const double* X0, const double* XX, const double* YY, const double* sigmasq,
                   const double* ratio, int lowk, double k0_low, const double* plin, const double* q3,
                   int nk, const double* k, double* out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && X0 && YY && sigmasq && k && out, "null argument");
        PRSD_REQUIRE(nrow > 0 && nk > 0, "Zeldovich: empty request");
        Context& c = ctx->c;
        // stencil plans are cached per k list
        static std::mutex mu;
        static std::list<std::pair<std::pair<int, std::vector<double>>, std::shared_ptr<ZelPlan>>> cache;
        std::shared_ptr<ZelPlan> zp;
        {
            std::lock_guard<std::mutex> lock(mu);
            std::pair<int, std::vector<double>> key{c.device, std::vector<double>(k, k + nk)};
            for (auto& e : cache) if (e.first == key) { zp = e.second; break; }
            if (!zp) {
                zp = std::make_shared<ZelPlan>();
                zp->build(c, k, nk);
                cache.emplace_front(key, zp);
                while (cache.size() > 8) cache.pop_back();
            }
        }
        DevBuf<double> dX0, dXX, dYY, dss, drat, dpl, dq3, dout((size_t)nrow*3*nk);
        dX0.upload(X0, (size_t)nrow*ZEL_N, c.stream);
        if (XX) dXX.upload(XX, (size_t)nrow*ZEL_N, c.stream);
        dYY.upload(YY, (size_t)nrow*ZEL_N, c.stream);
        dss.upload(sigmasq, nrow, c.stream);
        if (ratio) drat.upload(ratio, nrow, c.stream);
        if (lowk) {
            PRSD_REQUIRE(plin && q3, "Zeldovich: the low-k approximation needs P_lin(k) and Q3/k^4");
            dpl.upload(plin, (size_t)nrow*nk, c.stream);
            dq3.upload(q3, nrow, c.stream);
        }
        zeldovich_eval_map(c, *zp, nrow, nullptr, dX0.p, XX ? dXX.p : nullptr, dYY.p, dss.p, 1, ratio ? drat.p : nullptr,
                           lowk != 0, k0_low, lowk ? dpl.p : nullptr, lowk ? dq3.p : nullptr, 1, dout.p);
        dout.download(out, (size_t)nrow*3*nk, c.stream);
        c.sync();
    });
}

/* ------------------------------------------------------------ galaxy model */
static GalaxyConfig gal_cfg_from(const double* cfg)
{
    GalaxyConfig g;
    galaxy_config_default(g);
    if (cfg) {
        g.max_mu = (int)cfg[0]; g.fog_model = (int)cfg[1];
        g.use_P00_model = cfg[2] != 0; g.use_P01_model = cfg[3] != 0; g.use_P11_model = cfg[4] != 0;
        g.use_Pdv_model = cfg[5] != 0; g.use_Pvv_model = cfg[6] != 0; g.use_Phm_model = cfg[7] != 0;
        g.use_tidal_bias = cfg[8] != 0; g.use_so_correction = cfg[9] != 0;
        g.k0_low = cfg[10];
        g.interpolate = cfg[11] != 0;
        for (int i = 0; i < 10; i++) g.hz00[i] = cfg[16 + i];
        for (int i = 0; i < 9; i++) g.hz11[i] = cfg[26 + i];
        for (int i = 0; i < 20; i++) g.hzhm[i] = cfg[35 + i];
    }
    return g;
}

int prsd_galaxy_config_default(double* cfg64)
{
    return guarded([&] {
        PRSD_REQUIRE(cfg64, "null argument");
        GalaxyConfig g;
        galaxy_config_default(g);
        for (int i = 0; i < 64; i++) cfg64[i] = 0.0;
        cfg64[0] = g.max_mu; cfg64[1] = g.fog_model;
        cfg64[2] = g.use_P00_model; cfg64[3] = g.use_P01_model; cfg64[4] = g.use_P11_model;
        cfg64[5] = g.use_Pdv_model; cfg64[6] = g.use_Pvv_model; cfg64[7] = g.use_Phm_model;
        cfg64[8] = g.use_tidal_bias; cfg64[9] = g.use_so_correction; cfg64[10] = g.k0_low; cfg64[11] = g.interpolate;
        for (int i = 0; i < 10; i++) cfg64[16 + i] = g.hz00[i];
        for (int i = 0; i < 9; i++) cfg64[26 + i] = g.hz11[i];
        for (int i = 0; i < 20; i++) cfg64[35 + i] = g.hzhm[i];
    });
}

int prsd_galaxy_create(prsd_ctx* ctx, const double* cfg64, int nkm, const double* kmodel, int nk_interp,
                       const double* k_interp, prsd_galaxy** out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && kmodel && k_interp && out, "null argument");
        *out = new prsd_galaxy(ctx->c, gal_cfg_from(cfg64), kmodel, nkm, k_interp, nk_interp);
    });
}
int prsd_galaxy_destroy(prsd_galaxy* g) { return guarded(dev_of(g), [&] { delete g; }); }

int prsd_galaxy_power(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap, const double* par,
                      const double* stoch, int npts, const double* k, const double* mu, double* out)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && sp && res && par && stoch && out, "null argument");
        PRSD_REQUIRE((k == nullptr) == (mu == nullptr), "k and mu must both be given or both be NULL");
        Context& c = *g->g.ctx;
        if (k) {
            g->dk.upload(k, npts, c.stream);
            g->dmu.upload(mu, npts, c.stream);
            g->async_npts = npts;
        } else PRSD_REQUIRE(g->async_npts == npts && npts > 0, "k == NULL reuses the points of the previous asynchronous call (%d), not %d", g->async_npts, npts);
        g->dout.ensure((size_t)nw*npts);
        g->g.power(sp->s, res->r, nw, cmap, par, stoch, npts, g->dk.p, g->dmu.p, g->dout.p, out);
    });
}

/* asynchronous variants: see include/pyrsd_b200.h ("streaming") */
int prsd_galaxy_power_async(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap, const double* par,
                            const double* stoch, int npts, const double* k, const double* mu, double* out, long long* ticket)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && sp && res && par && stoch && out, "null argument");
        PRSD_REQUIRE((k == nullptr) == (mu == nullptr), "k and mu must both be given or both be NULL");
        Context& c = *g->g.ctx;
        if (k) {
            g->dk.upload(k, npts, c.stream);
            g->dmu.upload(mu, npts, c.stream);
            g->async_npts = npts;
        } else PRSD_REQUIRE(g->async_npts == npts && npts > 0, "k == NULL reuses the points of the previous asynchronous call (%d), not %d", g->async_npts, npts);
        if (g->dout.n < (size_t)nw*npts) { g->g.wait(); g->dout.ensure((size_t)nw*npts); }   // a copy may still read the old block
        g->g.power(sp->s, res->r, nw, cmap, par, stoch, npts, g->dk.p, g->dmu.p, g->dout.p, out, false);
        if (ticket) *ticket = g->g.last_ticket();
    });
}

int prsd_galaxy_set_mu_corrections(prsd_galaxy* g, int nw, const double* corr)
{
    return guarded(dev_of(g), [&] { PRSD_REQUIRE(g, "null argument"); g->g.set_mu_corrections(nw, corr); });
}

int prsd_galaxy_set_pair_overrides(prsd_galaxy* g, int nw, const double* ovr)
{
    return guarded(dev_of(g), [&] { PRSD_REQUIRE(g, "null argument"); g->g.set_pair_overrides(nw, ovr); });
}

int prsd_galaxy_set_loaded(prsd_galaxy* g, unsigned mask, const double* loaded)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g, "null argument");
        PRSD_REQUIRE(mask < 16u, "mask: bits 0..3 = Pdv, P00_mu0, P01_mu2, P11_mu4");
        g->g.set_loaded(mask, loaded);
    });
}

int prsd_galaxy_wait(prsd_galaxy* g)
{
    return guarded(dev_of(g), [&] { PRSD_REQUIRE(g, "null argument"); g->g.wait(); });
}

int prsd_galaxy_wait_copy(prsd_galaxy* g, long long ticket)
{
    return guarded(dev_of(g), [&] { PRSD_REQUIRE(g, "null argument"); g->g.wait_copy(ticket); });
}

int prsd_galaxy_power_device(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap, const double* par,
                             const double* stoch, int npts, const double* d_k, const double* d_mu, double* d_out)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && sp && res && par && stoch && d_k && d_mu && d_out, "null argument");
        g->g.power(sp->s, res->r, nw, cmap, par, stoch, npts, d_k, d_mu, d_out);
    });
}

int prsd_galaxy_poles(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap, const double* par,
                      const double* stoch, int nk, const double* k, int nell, const int* ells, int nmu, double* out)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && sp && res && par && stoch && k && ells && out, "null argument");
        PRSD_REQUIRE(nk > 0 && nell > 0, "prsd_galaxy_poles: empty request");
        Context& c = *g->g.ctx;
        g->dk.upload(k, nk, c.stream);
        g->dout.ensure((size_t)nw*nell*nk);
        g->g.poles(sp->s, res->r, nw, cmap, par, stoch, nk, g->dk.p, nell, ells, nmu, g->dout.p);
        g->dout.download(out, (size_t)nw*nell*nk, c.stream);
        c.sync();
    });
}

int prsd_galaxy_poles_device(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap, const double* par,
                             const double* stoch, int nk, const double* d_k, int nell, const int* ells, int nmu, double* d_out)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && sp && res && par && stoch && d_k && ells && d_out, "null argument");
        g->g.poles(sp->s, res->r, nw, cmap, par, stoch, nk, d_k, nell, ells, nmu, d_out);
    });
}

int prsd_galaxy_chi2(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap, const double* par,
                     const double* stoch, int nk, const double* k, int nell, const int* ells, int nmu,
                     const double* data, const double* cinv, double* chi2, double* poles)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && sp && res && par && stoch && k && ells && data && cinv && chi2, "null argument");
        PRSD_REQUIRE(nk > 0 && nell > 0 && nw > 0, "prsd_galaxy_chi2: empty request");
        Context& c = *g->g.ctx;
        const size_t n = (size_t)nell*nk;
        g->dk.upload(k, nk, c.stream);
        g->dout.ensure((size_t)nw*n);
        DevBuf<double> ddata, dcinv, dchi((size_t)nw);
        ddata.upload(data, n, c.stream);
        dcinv.upload(cinv, n*n, c.stream);
        g->g.chi2(sp->s, res->r, nw, cmap, par, stoch, nk, g->dk.p, nell, ells, nmu, ddata.p, dcinv.p, g->dout.p, dchi.p);
        dchi.download(chi2, (size_t)nw, c.stream);
        if (poles) g->dout.download(poles, (size_t)nw*n, c.stream);
        c.sync();
    });
}

int prsd_gridop_create(prsd_ctx* ctx, int nbin, int nk, int nmu, const double* weights, prsd_gridop** out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && weights && out, "null argument");
        *out = new prsd_gridop(ctx->c, nbin, nk, nmu, weights);
    });
}

int prsd_gridop_destroy(prsd_gridop* op)
{
    return guarded(dev_of(op), [&] { delete op; });
}

int prsd_gridop_apply(prsd_gridop* g, int nw, const double* power, double* out, int device_ptrs)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && power && out, "null argument");
        PRSD_REQUIRE(nw > 0, "prsd_gridop_apply: empty request");
        GridOp& op = g->op;
        Context& c = *op.ctx;
        if (device_ptrs) { op.apply(nw, power, out); return; }
        const size_t nin = (size_t)nw*op.nk*op.nmu, nout = (size_t)nw*op.nbin*op.nk;
        g->din.ensure(nin);
        g->dout.ensure(nout);
        g->din.upload(power, nin, c.stream);
        op.apply(nw, g->din.p, g->dout.p);
        g->dout.download(out, nout, c.stream);
        c.sync();
    });
}

int prsd_exchange_create(prsd_ctx* ctx, int world, int rank, long long n_per_rank, prsd_exchange** out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && out, "null argument");
        PRSD_REQUIRE(n_per_rank > 0, "prsd_exchange_create: empty contribution");
        *out = new prsd_exchange(ctx->c, world, rank, (size_t)n_per_rank);
    });
}

int prsd_exchange_destroy(prsd_exchange* x) { return guarded(dev_of(x), [&] { delete x; }); }

int prsd_exchange_handles(prsd_exchange* x, unsigned char* out128)
{
    return guarded(dev_of(x), [&] { PRSD_REQUIRE(x && out128, "null argument"); x->x.handles(out128); });
}

int prsd_exchange_open(prsd_exchange* x, const unsigned char* all_handles)
{
    return guarded(dev_of(x), [&] { PRSD_REQUIRE(x && all_handles, "null argument"); x->x.open(all_handles); });
}

int prsd_gridop_apply_exchange(prsd_gridop* g, int nw, const double* power, prsd_exchange* x)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && power && x, "null argument");
        PRSD_REQUIRE(g->op.ctx == x->x.ctx, "prsd_gridop_apply_exchange: the operator and the exchange belong to different contexts");
        PRSD_REQUIRE((size_t)nw*g->op.nbin*g->op.nk == x->x.n_per_rank,
                     "prsd_gridop_apply_exchange: %d walkers x %d bins x %d k do not fill the exchange's %zu doubles per rank",
                     nw, g->op.nbin, g->op.nk, x->x.n_per_rank);
        double* slot = x->x.begin_step();           // this rank's slot of its own receive buffer
        g->op.apply(nw, power, slot);
        x->x.finish_step();                         // push to the peers' slots + flag exchange
    });
}

int prsd_exchange_received(prsd_exchange* x, void** dev_ptr, int* status)
{
    return guarded(dev_of(x), [&] {
        PRSD_REQUIRE(x && dev_ptr, "null argument");
        *dev_ptr = (void*)x->x.received();
        if (status) *status = x->x.status();
    });
}

int prsd_window_create(prsd_ctx* ctx, int nk, const double* k, int nell, const int* ells, double q, int N, prsd_window** out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && k && ells && out, "null argument");
        *out = new prsd_window(ctx->c, nk, k, nell, ells, q, N);
    });
}

int prsd_window_destroy(prsd_window* w)
{
    return guarded(dev_of(w), [&] { delete w; });
}

int prsd_window_info(prsd_window* w, int* info, double* rr)
{
    return guarded(dev_of(w), [&] {
        PRSD_REQUIRE(w, "null argument");
        if (info) { info[0] = w->op.N; info[1] = w->op.off_in; info[2] = w->op.off_out; info[3] = w->op.nk; }
        if (rr) std::memcpy(rr, w->op.rr.data(), sizeof(double)*w->op.nk);
    });
}

int prsd_window_set_kernel(prsd_window* w, const double* kern)
{
    return guarded(dev_of(w), [&] {
        PRSD_REQUIRE(w && kern, "null argument");
        w->op.set_kernel(kern);
    });
}

int prsd_window_apply(prsd_window* w, int nw, const double* poles, double* out, int mode, int device_ptrs, double* xi,
                      double* xi_conv)
{
    return guarded(dev_of(w), [&] {
        PRSD_REQUIRE(w && poles && out, "null argument");
        PRSD_REQUIRE(nw > 0, "prsd_window_apply: empty request");
        WindowOp& op = w->op;
        Context& c = *op.ctx;
        if (device_ptrs) { op.apply(nw, poles, out, mode); return; }
        const size_t n = (size_t)nw*op.nell*op.nk;
        w->din.ensure(n);
        w->dout.ensure(n);
        w->din.upload(poles, n, c.stream);
        op.apply(nw, w->din.p, w->dout.p, mode);
        w->dout.download(out, n, c.stream);
        if (xi) op.xi.download(xi, n, c.stream);
        if (xi_conv) (mode == 0 ? op.xic : op.xi).download(xi_conv, n, c.stream);
        c.sync();
    });
}

int prsd_spline_resample_device(prsd_ctx* ctx, int nrows, int n, const double* x, const double* d_y, int nq, const double* xq,
                                double* d_out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && x && d_y && xq && d_out, "null argument");
        PRSD_REQUIRE(nrows > 0 && nq > 0, "prsd_spline_resample_device: empty request");
        for (int i = 1; i < n; i++) PRSD_REQUIRE(x[i] > x[i - 1], "prsd_spline_resample_device: x must be increasing");
        // The second derivatives of an interpolating cubic spline feel a far-away knot through a factor
        // (2 - sqrt 3)^distance = 0.268^distance: 64 knots beyond the queried range the end condition is below
        // 1e-36, so only that window of every row is solved (the k_out of a fit covers a fraction of the padded grid)
        double qlo = xq[0], qhi = xq[0];
        for (int i = 1; i < nq; i++) { qlo = std::min(qlo, xq[i]); qhi = std::max(qhi, xq[i]); }
        int lo = 0, hi = n - 1;
        while (lo + 1 < n && x[lo + 1] <= qlo) lo++;
        while (hi - 1 >= 0 && x[hi - 1] >= qhi) hi--;
        const int margin = 64;
        lo = std::max(0, lo - margin);
        hi = std::min(n - 1, hi + margin);
        if (qlo < x[0] || qhi > x[n - 1] || hi - lo + 1 < 4 || hi - lo + 1 > (3*n)/4) {
            spline_eval_host_x(ctx->c, nrows, n, x, d_y, nq, xq, d_out);
            return;
        }
        const int nwin = hi - lo + 1;
        DevBuf<double> win((size_t)nrows*nwin);
        PRSD_CUDA(cudaMemcpy2DAsync(win.p, sizeof(double)*nwin, d_y + lo, sizeof(double)*n, sizeof(double)*nwin, nrows,
                                    cudaMemcpyDeviceToDevice, ctx->c.stream));
        spline_eval_host_x(ctx->c, nrows, nwin, x + lo, win.p, nq, xq, d_out);
    });
}

int prsd_chi2(prsd_ctx* ctx, int nw, int n, const double* model, const double* data, const double* cinv, double* chi2,
              int model_on_device)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && model && data && cinv && chi2, "null argument");
        PRSD_REQUIRE(nw > 0 && n > 0, "prsd_chi2: empty request");
        Context& c = ctx->c;
        DevBuf<double> dmodel, ddata, dcinv, dchi((size_t)nw);
        ddata.upload(data, (size_t)n, c.stream);
        dcinv.upload(cinv, (size_t)n*n, c.stream);
        if (!model_on_device) dmodel.upload(model, (size_t)nw*n, c.stream);
        chi2_rows(c, nw, n, model_on_device ? model : dmodel.p, ddata.p, dcinv.p, dchi.p);
        dchi.download(chi2, (size_t)nw, c.stream);
        c.sync();
    });
}

int prsd_galaxy_base(prsd_galaxy* g, int nw, int* nrows, double* out)
{
    return guarded(dev_of(g), [&] {
        PRSD_REQUIRE(g && nrows, "null argument");
        *nrows = GalaxyModel::n_base();
        if (out) g->g.base_to_host(nw, out);
    });
}

int prsd_galaxy_moments(prsd_galaxy* g, int nw, double* out)
{
    return guarded(dev_of(g), [&] { PRSD_REQUIRE(g && out, "null argument"); g->g.moments_to_host(nw, out); });
}


/* -------------------------------------------- one-loop tables for the pygcl-style classes */
int prsd_oneloop_tables(prsd_spectra* sp, const double* cfg9, double* out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp, "null argument");
        Context& c = *sp->s.ctx;
        const QuadConfig cfg = cfg_from(cfg9);
        static std::mutex mu;
        static std::list<std::pair<std::pair<int, QuadConfig>, std::shared_ptr<PTPlan>>> cache;
        std::shared_ptr<PTPlan> plan;
        {
            std::lock_guard<std::mutex> lock(mu);
            for (auto& e : cache)
                if (e.first.first == c.device && e.first.second == cfg) { plan = e.second; break; }
            if (!plan) {
                plan = std::make_shared<PTPlan>(c, nullptr, 0, cfg, false, false, true);
                cache.emplace_front(std::make_pair(c.device, cfg), plan);
                while (cache.size() > 2) cache.pop_back();
            }
        }
        PTResult res;
        plan->run_oneloop(sp->s, res);
        if (out) sp->s.get_oneloop_tables(out);
    });
}

int prsd_spectra_set_oneloop_one(prsd_spectra* sp, int which, const double* table)
{
    return guarded(dev_of(sp), [&] { PRSD_REQUIRE(sp && table, "null argument"); sp->s.set_oneloop_table(which, table); });
}

int prsd_spectra_get_oneloop(prsd_spectra* sp, double* out)
{
    return guarded(dev_of(sp), [&] { PRSD_REQUIRE(sp && out, "null argument"); sp->s.get_oneloop_tables(out); });
}

/* CubicSpline(x, y)(xq)  [Spline.i, Spline.cpp:211-273, 362-367]: netlib cubic spline, 0 outside the domain */
int prsd_cubic_spline_eval(prsd_ctx* ctx, int batch, int n, const double* x, const double* y, int nq, const double* xq, double* out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && x && y && xq && out, "null argument");
        PRSD_REQUIRE(batch > 0 && nq > 0, "CubicSpline: empty request");
        for (int i = 1; i < n; i++) PRSD_REQUIRE(x[i] > x[i - 1], "CubicSpline: abscissae must be increasing");
        DevBuf<double> dy;
        dy.upload(y, (size_t)batch*n, ctx->c.stream);
        DevBuf<double> dout((size_t)batch*nq);
        spline_eval_host_x(ctx->c, batch, n, x, dy.p, nq, xq, dout.p);
        dout.download(out, (size_t)batch*nq, ctx->c.stream);
        ctx->c.sync();
    });
}


/* LinearSpline(x, y)(xq) / EvaluateDerivative  [Spline.i:24, Spline.cpp:58-130] */
int prsd_linear_spline_eval(prsd_ctx* ctx, int batch, int n, const double* x, const double* y, int nq, const double* xq, int deriv,
                            double* out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && x && y && xq && out, "null argument");
        PRSD_REQUIRE(batch > 0 && nq > 0 && n >= 2, "LinearSpline: empty request");
        for (int i = 1; i < n; i++) PRSD_REQUIRE(x[i] > x[i - 1], "LinearSpline: abscissae must be increasing");
        Context& c = ctx->c;
        DevBuf<double> dx, dy, dq, dout((size_t)batch*nq);
        dx.upload(x, n, c.stream);
        dy.upload(y, (size_t)batch*n, c.stream);
        dq.upload(xq, nq, c.stream);
        linear_spline_dev(c, batch, n, dx.p, dy.p, nq, dq.p, deriv, dout.p);
        dout.download(out, (size_t)batch*nq, c.stream);
        c.sync();
    });
}

/* ----------------------------------------------------- measurement support */
int prsd_ctx_profile(prsd_ctx* ctx, int enable)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx, "null argument");
        ctx->c.prof_reset();
        ctx->c.profiling = enable != 0;
    });
}

int prsd_ctx_profile_read(prsd_ctx* ctx, double* ms8, long long* count8)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && ms8 && count8, "null argument");
        ctx->c.prof_collect();
        static_assert(PROF_NTAGS == 8, "header documents 8 families");
        for (int t = 0; t < PROF_NTAGS; t++) { ms8[t] = ctx->c.prof_ms[t]; count8[t] = ctx->c.prof_count[t]; }
    });
}

int prsd_fp64_peak(prsd_ctx* ctx, int mode, double* tflops)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && tflops && (mode == 0 || mode == 1), "bad argument");
        *tflops = fp64_peak_tflops(ctx->c, mode);
    });
}

int prsd_gp_create(prsd_ctx* ctx, int n, int ndim, const double* x, const double* x_mean, const double* x_scale, const double* metric,
                   double amp, const double* alpha, double y_mean, double y_scale, prsd_gp** out)
{
    return guarded(dev_of(ctx), [&] {
        PRSD_REQUIRE(ctx && x && x_mean && x_scale && metric && alpha && out, "null argument");
        *out = new prsd_gp(ctx->c, n, ndim, x, x_mean, x_scale, metric, amp, alpha, y_mean, y_scale);
    });
}
int prsd_gp_destroy(prsd_gp* gp) { return guarded(dev_of(gp), [&] { delete gp; }); }
int prsd_gp_predict(prsd_gp* gp, int npred, const double* x, double* out, int device_ptrs)
{
    return guarded(dev_of(gp), [&] {
        PRSD_REQUIRE(gp && (npred == 0 || (x && out)), "null argument");
        if (device_ptrs) gp->m.predict_device(npred, x, out);
        else gp->m.predict(npred, x, out);
    });
}

int prsd_bias_to_mass(prsd_spectra* sp, int nw, const int* cmap, const double* rescale, const double* b1, double rho_bar,
                      double delta_halo, double mass_lo, double mass_hi, double* mass_out, double* sigma_out)
{
    return guarded(dev_of(sp), [&] {
        PRSD_REQUIRE(sp, "null argument");
        bias_to_mass(sp->s, nw, cmap, rescale, b1, rho_bar, delta_halo, mass_lo, mass_hi, mass_out, sigma_out);
    });
}

int prsd_spectra_reload(prsd_spectra* sp, const double* k, const double* T, const double* pars)
{
    return guarded(dev_of(sp), [&] { PRSD_REQUIRE(sp && k && T && pars, "null argument"); sp->s.reload_host(k, T, pars); });
}

/* page-locked host memory for the streaming entry points (callers without their own pinned allocator) */
int prsd_host_alloc(size_t bytes, void** out)
{
    return guarded([&] {
        PRSD_REQUIRE(out && bytes > 0, "bad argument");
        PRSD_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    });
}

int prsd_host_free(void* p)
{
    return guarded([&] { if (p) PRSD_CUDA(cudaFreeHost(p)); });
}

int prsd_spectra_reload_device(prsd_spectra* sp, const double* d_k, const double* d_T, const double* d_pars)
{
    return guarded(dev_of(sp), [&] { PRSD_REQUIRE(sp && d_k && d_T && d_pars, "null argument"); sp->s.reload_device(d_k, d_T, d_pars); });
}

} // extern "C"
