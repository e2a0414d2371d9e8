// spectra.cu -- device-resident spectra tables (see spectra.cuh).
#include "spectra.cuh"
#include <atomic>
#include "pcr.cuh"

#include <cmath>
#include <mutex>

namespace prsd {

const KnotGrid& knot_grid()
{
    static KnotGrid g = make_knot_grid();
    if (g.size() != TAB_M) throw Error(ERR_INTERNAL, strprintf("knot grid has %d knots, TAB_M = %d", g.size(), TAB_M));
    return g;
}

const double* knot_grid_dev(Context& ctx)
{
    static std::mutex mu;
    static std::vector<std::pair<int, double*>> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& e : cache) if (e.first == ctx.device) return e.second;
    double* d = nullptr;
    PRSD_CUDA(cudaMalloc((void**)&d, TAB_M*sizeof(double)));
    PRSD_CUDA(cudaMemcpy(d, knot_grid().x.data(), TAB_M*sizeof(double), cudaMemcpyHostToDevice));
    cache.push_back({ctx.device, d});
    return d;
}

// ------------------------------------------------------------ spline build
// Same arithmetic as cubic_spline_build() in pt_math.cuh (netlib fmm/spline.f), split so that the serial
// sweeps are as short as possible: the eliminated diagonal depends only on the abscissae (k_spline_factor:
// one division per knot, once per distinct abscissa set), after which a spline costs two sweeps of fused
// multiply-adds by one thread (operands in shared memory) between CTA-parallel set-up and coefficient passes.
__global__ void k_spline_factor(int n, const double* __restrict__ X, int x_stride, double* __restrict__ invb)
{
    const double* x = X + (size_t)blockIdx.x*x_stride;
    double* o = invb + (size_t)blockIdx.x*n;
    if (threadIdx.x != 0) return;
    double dm = x[1] - x[0];
    double ib = 1.0/(-dm);                       // b[0] = -d[0]
    o[0] = ib;
    for (int i = 1; i < n; i++) {
        const double di = (i < n - 1) ? x[i + 1] - x[i] : 0.0;
        const double bi = (i < n - 1) ? 2.0*(dm + di) : -dm;     // b[n-1] = -d[n-2]
        ib = 1.0/fma(-dm*dm, ib, bi);
        o[i] = ib;
        dm = di;
    }
}

__global__ void k_spline_build(int n, const double* __restrict__ X, int x_stride, const double* __restrict__ invb_all,
                               int invb_stride, const double* __restrict__ Y, double* __restrict__ out)
{
    extern __shared__ double sm[];
    double* x = sm;
    double* y = sm + n;
    double* ib = sm + 2*n;
    double* c = sm + 3*n;
    double* d = sm + 4*n;
    const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* Xs = X + (size_t)s*x_stride;
    const double* Ys = Y + (size_t)s*n;
    const double* IB = invb_all + (size_t)s*invb_stride;
    for (int i = tid; i < n; i += nt) { x[i] = Xs[i]; y[i] = Ys[i]; ib[i] = IB[i]; }
    __syncthreads();
    for (int i = tid; i < n - 1; i += nt) d[i] = x[i + 1] - x[i];
    __syncthreads();
    // right-hand side: second differences of the slopes
    for (int i = tid; i < n; i += nt)
        if (i >= 1 && i <= n - 2) c[i] = (y[i + 1] - y[i])/d[i] - (y[i] - y[i - 1])/d[i - 1];
    __syncthreads();
    if (tid == 0) {
        double c0 = 0.0, cn = 0.0;
        if (n > 3) {
            c0 = c[2]/(x[3] - x[1]) - c[1]/(x[2] - x[0]);
            cn = c[n - 2]/(x[n - 1] - x[n - 3]) - c[n - 3]/(x[n - 2] - x[n - 4]);
            c0 = c0*d[0]*d[0]/(x[3] - x[0]);
            cn = -cn*d[n - 2]*d[n - 2]/(x[n - 1] - x[n - 4]);
        }
        c[0] = c0;
        c[n - 1] = cn;
        // forward elimination of the right-hand side: c_i -= (d_{i-1}/b_{i-1}) c_{i-1}
        double cp = c0;
        for (int i = 1; i < n; i++) {
            cp = fma(-(d[i - 1]*ib[i - 1]), cp, c[i]);
            c[i] = cp;
        }
        // back substitution
        double cn1 = c[n - 1]*ib[n - 1];
        c[n - 1] = cn1;
        for (int i = n - 2; i >= 0; i--) {
            cn1 = fma(-d[i], cn1, c[i])*ib[i];
            c[i] = cn1;
        }
    }
    __syncthreads();
    double* o = out + (size_t)s*5*n;
    for (int i = tid; i < n; i += nt) {
        double bi, ci, di;
        if (i < n - 1) {
            bi = (y[i + 1] - y[i])/d[i] - d[i]*(c[i + 1] + 2.0*c[i]);
            di = (c[i + 1] - c[i])/d[i];
        } else {
            bi = (y[n - 1] - y[n - 2])/d[n - 2] + d[n - 2]*(c[n - 2] + 2.0*c[n - 1]);
            di = (c[n - 1] - c[n - 2])/d[n - 2];
        }
        ci = 3.0*c[i];
        o[i] = x[i];
        o[n + i] = y[i];
        o[2*n + i] = bi;
        o[3*n + i] = ci;
        o[4*n + i] = di;
    }
}

// The same spline with the tridiagonal system solved by parallel cyclic reduction (pcr.cuh): no serial sweep at all.
// Shared memory: x, y, h, sigma and the eight reduction arrays, 12 n doubles (n <= 2048).
__global__ void __launch_bounds__(256)
k_spline_build_pcr(int n, const double* __restrict__ X, int x_stride, const double* __restrict__ Y, double* __restrict__ out)
{
    extern __shared__ double sm[];
    double* x = sm;
    double* y = sm + n;
    double* h = sm + 2*n;
    double* sig = sm + 3*n;
    double* w = sm + 4*n;                 // a b c r | a2 b2 c2 r2
    const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* Xs = X + (size_t)s*x_stride;
    const double* Ys = Y + (size_t)s*n;
    for (int i = tid; i < n; i += nt) { x[i] = Xs[i]; y[i] = Ys[i]; }
    __syncthreads();
    for (int i = tid; i < n - 1; i += nt) h[i] = x[i + 1] - x[i];
    __syncthreads();
    double* a = w; double* b = w + n; double* c = w + 2*n; double* r = w + 3*n;
    for (int i = tid; i < n; i += nt) {
        if (i >= 1 && i <= n - 2) {
            a[i] = h[i - 1];
            b[i] = 2.0*(h[i - 1] + h[i]);
            c[i] = h[i];
            r[i] = (y[i + 1] - y[i])/h[i] - (y[i] - y[i - 1])/h[i - 1];
        }
    }
    __syncthreads();
    if (tid == 0) {
        // end conditions of netlib spline.f (Spline.cpp:229-242): third derivatives from the cubics through
        // the first / last four points
        double r0 = 0.0, rn = 0.0;
        if (n > 3) {
            r0 = r[2]/(x[3] - x[1]) - r[1]/(x[2] - x[0]);
            rn = r[n - 2]/(x[n - 1] - x[n - 3]) - r[n - 3]/(x[n - 2] - x[n - 4]);
            r0 = r0*h[0]*h[0]/(x[3] - x[0]);
            rn = -rn*h[n - 2]*h[n - 2]/(x[n - 1] - x[n - 4]);
        }
        a[0] = 0.0;       b[0] = -h[0];           c[0] = h[0];      r[0] = r0;
        a[n - 1] = h[n - 2]; b[n - 1] = -h[n - 2]; c[n - 1] = 0.0;  r[n - 1] = rn;
    }
    __syncthreads();
    pcr_solve(1, n, a, b, c, r, w + 4*n, w + 5*n, w + 6*n, w + 7*n, sig);
    double* o = out + (size_t)s*5*n;
    for (int i = tid; i < n; i += nt) {
        double bi, di;
        if (i < n - 1) {
            bi = (y[i + 1] - y[i])/h[i] - h[i]*(sig[i + 1] + 2.0*sig[i]);
            di = (sig[i + 1] - sig[i])/h[i];
        } else {
            bi = (y[n - 1] - y[n - 2])/h[n - 2] + h[n - 2]*(sig[n - 2] + 2.0*sig[n - 1]);
            di = (sig[n - 1] - sig[n - 2])/h[n - 2];
        }
        o[i] = x[i];
        o[n + i] = y[i];
        o[2*n + i] = bi;
        o[3*n + i] = 3.0*sig[i];
        o[4*n + i] = di;
    }
}

// The same spline when the abscissae are FIXED (the 1000-point one-loop grid): the second-derivative vector sigma is a
// fixed linear map of the ordinates (the inverse of the tridiagonal system, built once on the host with
// cubic_spline_build itself); a far knot enters through 0.268^distance, so a band of `band` columns is exact to
// < 1e-17 and the solve becomes a banded product -- no reduction sweeps, no barriers between them.
__global__ void __launch_bounds__(256)
k_spline_build_band(int n, int band, const double* __restrict__ X, const double* __restrict__ Sb /*[band][n]*/,
                    const int* __restrict__ lo, const double* __restrict__ Y, double* __restrict__ out)
{
    extern __shared__ double sm_b[];
    double* x = sm_b;
    double* y = sm_b + n;
    double* sig = sm_b + 2*n;
    const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* Ys = Y + (size_t)s*n;
    for (int i = tid; i < n; i += nt) { x[i] = X[i]; y[i] = Ys[i]; }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const double* yy = y + lo[i];
        // four independent partial sums, loads of a group in flight together: the loop waits on L2 latency (ncu:
        // long-scoreboard stalls 30 per issue with one chain)
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int jj = 0;
        for (; jj + 3 < band; jj += 4) {
            const double e0 = __ldg(Sb + (size_t)jj*n + i), e1 = __ldg(Sb + (size_t)(jj + 1)*n + i);
            const double e2 = __ldg(Sb + (size_t)(jj + 2)*n + i), e3 = __ldg(Sb + (size_t)(jj + 3)*n + i);
            a0 = fma(e0, yy[jj], a0);
            a1 = fma(e1, yy[jj + 1], a1);
            a2 = fma(e2, yy[jj + 2], a2);
            a3 = fma(e3, yy[jj + 3], a3);
        }
        for (; jj < band; jj++) a0 = fma(__ldg(Sb + (size_t)jj*n + i), yy[jj], a0);
        sig[i] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    double* o = out + (size_t)s*5*n;
    for (int i = tid; i < n; i += nt) {
        double bi, di;
        if (i < n - 1) {
            const double h = x[i + 1] - x[i];
            bi = (y[i + 1] - y[i])/h - h*(sig[i + 1] + 2.0*sig[i]);
            di = (sig[i + 1] - sig[i])/h;
        } else {
            const double h = x[n - 1] - x[n - 2];
            bi = (y[n - 1] - y[n - 2])/h + h*(sig[n - 2] + 2.0*sig[n - 1]);
            di = (sig[n - 1] - sig[n - 2])/h;
        }
        o[i] = x[i];
        o[n + i] = y[i];
        o[2*n + i] = bi;
        o[3*n + i] = 3.0*sig[i];
        o[4*n + i] = di;
    }
}

// host: banded operator y -> sigma of the netlib spline on the abscissae X (band columns around every row)
void spline_band_operator(int n, const double* X, int band, std::vector<double>& Sb, std::vector<int>& lo)
{
    std::vector<double> op((size_t)n*n), y(n), b(n), c(n), d(n);
    for (int j = 0; j < n; j++) {
        std::fill(y.begin(), y.end(), 0.0);
        y[j] = 1.0;
        cubic_spline_build(n, X, y.data(), b.data(), c.data(), d.data());
        for (int i = 0; i < n; i++) op[(size_t)i*n + j] = c[i]*(1.0/3.0);
    }
    Sb.assign((size_t)band*n, 0.0);
    lo.assign(n, 0);
    for (int i = 0; i < n; i++) {
        lo[i] = std::min(std::max(i - band/2, 0), n - band);
        for (int jj = 0; jj < band; jj++) Sb[(size_t)jj*n + i] = op[(size_t)i*n + lo[i] + jj];
    }
}

void spline_factor(Context& ctx, int nfac, int n, const double* dX, int x_stride, double* d_invb)
{
    k_spline_factor<<<nfac, 32, 0, ctx.stream>>>(n, dX, x_stride, d_invb);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

void spline_build_batched(Context& ctx, int nspline, int n, const double* dX, int x_stride,
                          const double* dY, double* dspl, const double* d_invb)
{
    PRSD_REQUIRE(n >= 4 && n <= 4096, "spline_build_batched: n = %d out of range [4, 4096]", n);
    if (n <= 2048) {
        const size_t smem_pcr = (size_t)12*n*sizeof(double);
        if (smem_pcr > 48*1024)
            PRSD_CUDA(cudaFuncSetAttribute(k_spline_build_pcr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pcr));
        ProfScope prof(ctx, PROF_SPLINE);
        k_spline_build_pcr<<<nspline, 256, smem_pcr, ctx.stream>>>(n, dX, x_stride, dY, dspl);
        PRSD_CUDA(cudaGetLastError());
        ctx.launches++;
        return;
    }
    // longer splines: serial sweeps with factored abscissae
    const size_t smem = (size_t)5*n*sizeof(double);
    if (smem > 48*1024)
        PRSD_CUDA(cudaFuncSetAttribute(k_spline_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof(ctx, PROF_SPLINE);
    static thread_local DevBuf<double> scratch;       // factors when the caller holds none
    if (!d_invb) {
        const int nfac = x_stride == 0 ? 1 : nspline;
        scratch.ensure((size_t)nfac*n);
        spline_factor(ctx, nfac, n, dX, x_stride, scratch.p);
        d_invb = scratch.p;
    }
    k_spline_build<<<nspline, 128, smem, ctx.stream>>>(n, dX, x_stride, d_invb, x_stride == 0 ? 0 : n, dY, dspl);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

// ------------------------------------------------------------ table builds
__global__ void k_plin_table(int nspl, const PlinParams* __restrict__ params, const double* __restrict__ spl,
                             const double* __restrict__ lnq, double* __restrict__ tables)
{
    const int c = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= TAB_MPAD) return;
    double* lin = tables + (size_t)(c*SP_NKIND + SP_LIN)*TAB_MPAD;
    double* sq = tables + (size_t)(c*SP_NKIND + SP_LIN_SQ)*TAB_MPAD;
    double v = 0.0;
    if (j < TAB_M) v = plin_eval(params[c], spl + (size_t)c*5*nspl, exp(lnq[j]));
    lin[j] = v;
    sq[j] = v*v;
}

__global__ void k_plin_eval(int nspl, const PlinParams* __restrict__ params, const double* __restrict__ spl,
                            int nq, const double* __restrict__ q, double* __restrict__ out)
{
    const int c = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= nq) return;
    out[(size_t)c*nq + j] = plin_eval(params[c], spl + (size_t)c*5*nspl, q[j]);
}

__global__ void k_scale_linear(const double* __restrict__ fac, PlinParams* __restrict__ params, double* __restrict__ tables)
{
    const int c = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    const double f = fac[c];
    if (j == 0) params[c].amp *= f;
    if (j >= TAB_MPAD) return;
    tables[(size_t)(c*SP_NKIND + SP_LIN)*TAB_MPAD + j] *= f;
    tables[(size_t)(c*SP_NKIND + SP_LIN_SQ)*TAB_MPAD + j] *= f*f;
}

SpectraBatch::SpectraBatch(Context& c, int ncosmo_, int nspl_, const double* k, const double* T, const double* pars)
    : ctx(&c), ncosmo(ncosmo_), nspl(nspl_)
{
    PRSD_REQUIRE(ncosmo > 0, "SpectraBatch: ncosmo must be positive");
    static std::atomic<uint64_t> next_uid{1};
    uid = next_uid++;
    PRSD_REQUIRE(nspl >= 4 && nspl <= 4096, "SpectraBatch: %d transfer-function samples unsupported", nspl);
    h_params.resize(ncosmo);
    for (int i = 0; i < ncosmo; i++) {
        const double* kk = k + (size_t)i*nspl;
        const double* TT = T + (size_t)i*nspl;
        const double* p = pars + (size_t)i*8;
        for (int j = 0; j < nspl; j++) {
            PRSD_REQUIRE(kk[j] > 0 && (j == 0 || kk[j] > kk[j - 1]), "SpectraBatch: k must be positive and increasing (cosmology %d, sample %d)", i, j);
            PRSD_REQUIRE(std::isfinite(TT[j]), "SpectraBatch: non-finite T(k) (cosmology %d, sample %d)", i, j);
        }
        // pars = {n_s, amplitude, h, Omega0_m, Omega0_b, T_cmb, -, -}
        plin_params_init(h_params[i], nspl, kk, TT, p[0], p[1], p[2], p[3], p[4], p[5]);
    }
    cudaStream_t s = ctx->stream;
    params.upload(h_params, s);
    DevBuf<double> dk, dT;
    dk.upload(k, (size_t)ncosmo*nspl, s);
    dT.upload(T, (size_t)ncosmo*nspl, s);
    spl.alloc((size_t)ncosmo*5*nspl);
    spline_build_batched(*ctx, ncosmo, nspl, dk.p, nspl, dT.p, spl.p);
    tables.alloc((size_t)ncosmo*SP_NKIND*TAB_MPAD);
    tables.zero(s);
    dim3 grid((TAB_MPAD + 255)/256, ncosmo);
    k_plin_table<<<grid, 256, 0, s>>>(nspl, params.p, spl.p, knot_grid_dev(*ctx), tables.p);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
    ctx->sync();
}

__global__ void k_params_init(int nspl, const double* __restrict__ k, const double* __restrict__ T,
                              const double* __restrict__ pars, PlinParams* __restrict__ out)
{
    const int c = blockIdx.x;
    const double* p = pars + (size_t)c*8;
    PlinParams P;
    plin_params_init(P, nspl, k + (size_t)c*nspl, T + (size_t)c*nspl, p[0], p[1], p[2], p[3], p[4], p[5]);
    out[c] = P;
}

void SpectraBatch::reload_device(const double* d_k, const double* d_T, const double* d_pars)
{
    cudaStream_t s = ctx->stream;
    k_params_init<<<ncosmo, 1, 0, s>>>(nspl, d_k, d_T, d_pars, params.p);
    ctx->launches++;
    spline_build_batched(*ctx, ncosmo, nspl, d_k, nspl, d_T, spl.p);
    dim3 grid((TAB_MPAD + 255)/256, ncosmo);
    k_plin_table<<<grid, 256, 0, s>>>(nspl, params.p, spl.p, knot_grid_dev(*ctx), tables.p);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
    have_oneloop = false;
    lin_gen++;
}

void SpectraBatch::reload_host(const double* k, const double* T, const double* pars)
{
    // The copies run on the upload stream, one batch ahead of the kernels: they only wait until the previous batch's
    // table build has read the staging (early in that batch), the main stream only waits for their end.
    cudaStream_t s = ctx->stream, up = ctx->uploads();
    if (!ev_in_ready) {
        PRSD_CUDA(cudaEventCreateWithFlags(&ev_in_ready, cudaEventDisableTiming));
        PRSD_CUDA(cudaEventCreateWithFlags(&ev_in_free, cudaEventDisableTiming));
        in_k.ensure((size_t)ncosmo*nspl);
        in_T.ensure((size_t)ncosmo*nspl);
        in_pars.ensure((size_t)ncosmo*8);
        PRSD_CUDA(cudaEventRecord(ev_in_free, s));       // whatever the main stream still does with fresh blocks comes first
    }
    PRSD_CUDA(cudaStreamWaitEvent(up, ev_in_free, 0));
    in_k.upload(k, (size_t)ncosmo*nspl, up);
    in_T.upload(T, (size_t)ncosmo*nspl, up);
    in_pars.upload(pars, (size_t)ncosmo*8, up);
    PRSD_CUDA(cudaEventRecord(ev_in_ready, up));
    PRSD_CUDA(cudaStreamWaitEvent(s, ev_in_ready, 0));
    reload_device(in_k.p, in_T.p, in_pars.p);
    PRSD_CUDA(cudaEventRecord(ev_in_free, s));
}

SpectraBatch::~SpectraBatch()
{
    if (ev_in_ready) {
        if (ctx->upload_stream) cudaStreamSynchronize(ctx->upload_stream);
        cudaEventDestroy(ev_in_ready);
        cudaEventDestroy(ev_in_free);
    }
}

void SpectraBatch::rescale_linear(const double* fac_host)
{
    DevBuf<double> f;
    f.upload(fac_host, ncosmo, ctx->stream);
    dim3 grid((TAB_MPAD + 255)/256, ncosmo);
    k_scale_linear<<<grid, 256, 0, ctx->stream>>>(f.p, params.p, tables.p);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
    for (int i = 0; i < ncosmo; i++) h_params[i].amp *= fac_host[i];
    lin_gen++;
    ctx->sync();
}

void SpectraBatch::eval_linear(int nq, const double* q_host, double* out_host)
{
    if (nq <= 0) return;
    DevBuf<double> q, o((size_t)ncosmo*nq);
    q.upload(q_host, nq, ctx->stream);
    dim3 grid((nq + 255)/256, ncosmo);
    k_plin_eval<<<grid, 256, 0, ctx->stream>>>(nspl, params.p, spl.p, nq, q.p, o.p);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
    o.download(out_host, (size_t)ncosmo*nq, ctx->stream);
    ctx->sync();
}

// --------------------------------------------------------- one-loop spectra
// OneLoopPS::Evaluate returns the spline inside [KMIN, KMAX] and 0 outside
// (OneLoopPS.cpp:28-33); the tables keep exactly that (zone edges sit on 1e-5 and 10).
__global__ void k_oneloop_tabulate(const double* __restrict__ ospl, const double* __restrict__ lnq,
                                   double lnk0, double inv_dlnk, int jlo, int jhi, double* __restrict__ tables)
{
    const int c = blockIdx.y, which = blockIdx.z;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= TAB_MPAD) return;
    double v = 0.0;
    if (j >= jlo && j <= jhi) {     // knots of the zones inside [KMIN, KMAX]
        const double q = exp(lnq[j]);
        const double* s = ospl + (size_t)(c*4 + which)*5*ONELOOP_N;
        // knots that coincide with the hard limits may differ from them by an ulp
        double qq = q;
        if (fabs(q - ONELOOP_KMIN) < 1e-12*ONELOOP_KMIN) qq = ONELOOP_KMIN;
        if (fabs(q - ONELOOP_KMAX) < 1e-12*ONELOOP_KMAX) qq = ONELOOP_KMAX;
        v = cubic_spline_eval_log(ONELOOP_N, s, lnk0, inv_dlnk, qq);
    }
    tables[(size_t)(c*SP_NKIND + SP_DD1 + which)*TAB_MPAD + j] = v;
}

__global__ void k_oneloop_eval(const double* __restrict__ ospl, int nspl, const PlinParams* __restrict__ params,
                               const double* __restrict__ spl, int which, int full, double lnk0, double inv_dlnk,
                               int nq, const double* __restrict__ q, double* __restrict__ out)
{
    const int c = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= nq) return;
    const double k = q[j];
    const double* s = ospl + (size_t)(c*4 + which)*5*ONELOOP_N;
    double v = (k < ONELOOP_KMIN || k > ONELOOP_KMAX) ? 0.0 : cubic_spline_eval_log(ONELOOP_N, s, lnk0, inv_dlnk, k);
    if (full) {
        // OneLoopPS::EvaluateFull (OneLoopPS.cpp:21-26); P22bar has no linear part (:150-155)
        if (k > ONELOOP_KMAX) v = 0.0;
        else if (which != 3) v += plin_eval(params[c], spl + (size_t)c*5*nspl, k);
    }
    out[(size_t)c*nq + j] = v;
}

std::vector<double> oneloop_kgrid();   // defined in plan.cu

void SpectraBatch::set_oneloop_tables(const double* d_tables)
{
    static DevBuf<double> dk, dk_band;    // shared abscissae and the banded spline operator on them
    static DevBuf<int> dk_lo;
    static int dk_dev = -1;
    const int band = 64;
    if (dk_dev != ctx->device) {
        const std::vector<double> kg1 = oneloop_kgrid();
        dk.upload(kg1, ctx->stream);
        std::vector<double> Sb;
        std::vector<int> lo;
        spline_band_operator(ONELOOP_N, kg1.data(), band, Sb, lo);
        dk_band.upload(Sb, ctx->stream);
        dk_lo.upload(lo, ctx->stream);
        dk_dev = ctx->device;
    }
    oneloop_spl.ensure((size_t)ncosmo*4*5*ONELOOP_N);
    oneloop_raw.ensure((size_t)ncosmo*4*ONELOOP_N);
    if (d_tables != oneloop_raw.p)
        PRSD_CUDA(cudaMemcpyAsync(oneloop_raw.p, d_tables, (size_t)ncosmo*4*ONELOOP_N*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    {
        ProfScope prof(*ctx, PROF_SPLINE);
        k_spline_build_band<<<ncosmo*4, 256, 3*ONELOOP_N*sizeof(double), ctx->stream>>>(ONELOOP_N, band, dk.p, dk_band.p, dk_lo.p,
                                                                                    oneloop_raw.p, oneloop_spl.p);
        PRSD_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    const double lnk0 = std::log(ONELOOP_KMIN);
    const double inv = (ONELOOP_N - 1)/std::log(ONELOOP_KMAX/ONELOOP_KMIN);
    // zones 1..3 of the knot grid span [1e-5, 10] exactly (quad.cpp make_knot_grid)
    const KnotGrid& kg = knot_grid();
    const int jlo = kg.zones[1].start, jhi = kg.zones[3].start + kg.zones[3].n;
    dim3 grid((TAB_MPAD + 255)/256, ncosmo, 4);
    k_oneloop_tabulate<<<grid, 256, 0, ctx->stream>>>(oneloop_spl.p, knot_grid_dev(*ctx), lnk0, inv, jlo, jhi, tables.p);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
    have_oneloop = true;
}

void SpectraBatch::set_oneloop_table(int which, const double* h_table)
{
    PRSD_REQUIRE(which >= 0 && which < 4, "set_oneloop_table: invalid spectrum %d", which);
    if (!oneloop_raw.n) {
        oneloop_raw.alloc((size_t)ncosmo*4*ONELOOP_N);
        oneloop_raw.zero(ctx->stream);
    }
    for (int c = 0; c < ncosmo; c++)
        PRSD_CUDA(cudaMemcpyAsync(oneloop_raw.p + ((size_t)c*4 + which)*ONELOOP_N, h_table + (size_t)c*ONELOOP_N,
                                  ONELOOP_N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    set_oneloop_tables(oneloop_raw.p);
    ctx->sync();
}

void SpectraBatch::get_oneloop_tables(double* h_out)
{
    PRSD_REQUIRE(have_oneloop, "one-loop tables have not been computed for this batch");
    oneloop_raw.download(h_out, (size_t)ncosmo*4*ONELOOP_N, ctx->stream);
    ctx->sync();
}

void SpectraBatch::eval_oneloop(int which, bool full, int nq, const double* q_host, double* out_host)
{
    PRSD_REQUIRE(have_oneloop, "one-loop tables have not been computed for this batch");
    PRSD_REQUIRE(which >= 0 && which < 4, "eval_oneloop: invalid spectrum %d", which);
    if (nq <= 0) return;
    DevBuf<double> q, o((size_t)ncosmo*nq);
    q.upload(q_host, nq, ctx->stream);
    const double lnk0 = std::log(ONELOOP_KMIN);
    const double inv = (ONELOOP_N - 1)/std::log(ONELOOP_KMAX/ONELOOP_KMIN);
    dim3 grid((nq + 255)/256, ncosmo);
    k_oneloop_eval<<<grid, 256, 0, ctx->stream>>>(oneloop_spl.p, nspl, params.p, spl.p, which, full ? 1 : 0, lnk0, inv, nq, q.p, o.p);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
    o.download(out_host, (size_t)ncosmo*nq, ctx->stream);
    ctx->sync();
}

} // namespace prsd
