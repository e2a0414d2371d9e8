// fftlog.cu -- batched FFTLog (see fftlog.cuh).
#include "fftlog.cuh"
#include "fftlog_host.h"

#include <cmath>
#include <list>
#include <mutex>

namespace prsd {

// --------------------------------------------------------------- plan cache
struct FKey { int dev, N, kropt, dir; double mu, q, dlnr, kr; };
static bool operator==(const FKey& a, const FKey& b)
{
    return a.dev == b.dev && a.N == b.N && a.kropt == b.kropt && a.dir == b.dir && a.mu == b.mu && a.q == b.q && a.dlnr == b.dlnr && a.kr == b.kr;
}
static std::mutex g_mu;
static std::list<std::pair<FKey, std::shared_ptr<FFTLogPlan>>> g_cache;

std::shared_ptr<FFTLogPlan> fftlog_plan(Context& ctx, int N, double mu, double q, double dlnr, double kr, int kropt, int dir)
{
    PRSD_REQUIRE(dir == 1 || dir == -1, "FFTLog: direction must be +1 or -1");
    PRSD_REQUIRE(N >= 2 && N <= 8192, "FFTLog: N = %d outside [2, 8192]", N);
    PRSD_REQUIRE(dlnr != 0.0 && std::isfinite(dlnr), "FFTLog: dlnr must be non-zero");
    PRSD_REQUIRE(kr > 0, "FFTLog: kr must be positive");
    FKey key{ctx.device, N, kropt, dir, mu, q, dlnr, kr};
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& e : g_cache) if (e.first == key) return e.second;
    auto p = std::make_shared<FFTLogPlan>();
    p->N = N; p->mu = mu; p->q = q; p->dlnr = dlnr;
    p->kr = kropt ? fftlog_krgood(mu, q, dlnr, kr) : kr;
    p->pow2 = (N & (N - 1)) == 0 && N >= 8;
    std::vector<double> re, im;
    fftlog_multipliers(N, mu, q, dlnr, p->kr, re, im);
    if (dir == -1) {
        // backward transform (fftlog.f:523-527, fhtq :644-784): the rotation by the phase of u_m is the SAME, its
        // amplitude divides instead of multiplying (u_m -> u_m / |u_m|^2, m = 0 and the Nyquist term -> 1 / u, with
        // 0 -> 0 for the singular m = 0 case), and both power-law biases change sign (q -> -q in the bias factors only)
        const int nh = N/2;
        re[0] = re[0] == 0.0 ? 0.0 : 1.0/re[0];
        for (int m = 1; m <= nh; m++) {
            if (m == nh && N % 2 == 0) { re[m] = 1.0/re[m]; im[m] = 0.0; continue; }
            const double a2 = re[m]*re[m] + im[m]*im[m];
            re[m] /= a2;
            im[m] /= a2;
        }
        p->q = -q;
    }
    p->ure.upload(re, ctx.stream);
    p->uim.upload(im, ctx.stream);
    ctx.sync();
    g_cache.emplace_front(key, p);
    while (g_cache.size() > 64) g_cache.pop_back();
    return p;
}

std::shared_ptr<FFTLogPlan> fftlog_plan_custom(Context& ctx, int N, double dlnr, const std::vector<double>& re,
                                               const std::vector<double>& im)
{
    PRSD_REQUIRE(N >= 2 && N <= 8192, "FFTLog: N = %d outside [2, 8192]", N);
    PRSD_REQUIRE((int)re.size() == N/2 + 1 && (int)im.size() == N/2 + 1, "FFTLog: multiplier table must have N/2 + 1 entries");
    auto p = std::make_shared<FFTLogPlan>();
    p->N = N; p->mu = 0; p->q = 0; p->dlnr = dlnr; p->kr = 1;
    p->pow2 = (N & (N - 1)) == 0 && N >= 8;
    p->ure.upload(re, ctx.stream);
    p->uim.upload(im, ctx.stream);
    ctx.sync();
    return p;
}

// ------------------------------------------------------------ real-space g
// g[d] = (1/N) sum_{m=0}^{N-1} U_m exp(2 pi i d m / N), U_{N-m} = conj(U_m)
__global__ void k_fftlog_g(int N, const double* __restrict__ ure, const double* __restrict__ uim, double* __restrict__ g)
{
    const int d = blockIdx.x*blockDim.x + threadIdx.x;
    if (d >= N) return;
    const int nh = N/2;
    double s = ure[0];
    const int mmax = (N % 2 == 0) ? nh - 1 : nh;
    for (int m = 1; m <= mmax; m++) {
        const long long dm = ((long long)d*m) % N;
        double sn, cs;
        sincospi(2.0*(double)dm/(double)N, &sn, &cs);
        s += 2.0*(ure[m]*cs - uim[m]*sn);
    }
    if (N % 2 == 0) s += ure[nh]*((d & 1) ? -1.0 : 1.0);
    g[d] = s/N;
}

void FFTLogPlan::build_g(Context& ctx)
{
    if (g.n) return;
    g.alloc(N);
    k_fftlog_g<<<(N + 127)/128, 128, 0, ctx.stream>>>(N, ure.p, uim.p, g.p);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

// ----------------------------------------------------- power-of-two kernel
// smem: z[N] (double2) | tw[N/2] (double2)
__global__ void __launch_bounds__(256)
k_fftlog_pow2(int N, int logN, int batch, double q, double dlnr, double lnkr,
              const double* __restrict__ ure, const double* __restrict__ uim,
              const double* __restrict__ in, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    double2* z = reinterpret_cast<double2*>(smem_raw2);
    double2* tw = z + N;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b0 = 2*blockIdx.x, b1 = b0 + 1;
    const bool two = b1 < batch;
    const double jc = 0.5*(N + 1);

    for (int j = tid; j < N/2; j += nt) {
        double s, c;
        sincospi(-2.0*(double)j/(double)N, &s, &c);
        tw[j] = make_double2(c, s);
    }
    for (int i = tid; i < N; i += nt) {
        double a0 = in[(size_t)b0*N + i];
        double a1 = two ? in[(size_t)b1*N + i] : 0.0;
        if (q != 0.0) {
            const double f = exp(-q*((i + 1) - jc)*dlnr);      // fht: a(r) = A(r) (r/rc)^(-q)
            a0 *= f; a1 *= f;
        }
        z[i] = make_double2(a0, a1);
    }
    __syncthreads();
    // forward DIF: natural order in, bit-reversed order out
    for (int half = N >> 1, shift = 0; half >= 1; half >>= 1, shift++) {
        for (int t = tid; t < N/2; t += nt) {
            const int j = t & (half - 1);
            const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
            const double2 u = z[i0], v = z[i1];
            const double2 w = tw[j << shift];
            const double dr = u.x - v.x, di = u.y - v.y;
            z[i0] = make_double2(u.x + v.x, u.y + v.y);
            z[i1] = make_double2(dr*w.x - di*w.y, dr*w.y + di*w.x);
        }
        __syncthreads();
    }
    // multiply by U_m; slot p holds mode m = bitrev(p)
    for (int p = tid; p < N; p += nt) {
        const int m = (int)(__brev((unsigned)p) >> (32 - logN));
        double ur, ui;
        if (m <= N/2) { ur = ure[m]; ui = uim[m]; }
        else { ur = ure[N - m]; ui = -uim[N - m]; }
        const double2 v = z[p];
        z[p] = make_double2(v.x*ur - v.y*ui, v.x*ui + v.y*ur);
    }
    __syncthreads();
    // inverse DIT: bit-reversed in, natural out (conjugate twiddles)
    for (int half = 1, shift = logN - 1; half < N; half <<= 1, shift--) {
        for (int t = tid; t < N/2; t += nt) {
            const int j = t & (half - 1);
            const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
            const double2 w = tw[j << shift];
            const double2 u = z[i0], v0 = z[i1];
            const double vr = v0.x*w.x + v0.y*w.y, vi = v0.y*w.x - v0.x*w.y;   // v * conj(w)
            z[i0] = make_double2(u.x + vr, u.y + vi);
            z[i1] = make_double2(u.x - vr, u.y - vi);
        }
        __syncthreads();
    }
    // reversal, 1/N, output bias (fhtq :772-782, fht :632-638)
    const double invN = 1.0/N;
    for (int i = tid; i < N; i += nt) {
        const double2 v = z[N - 1 - i];
        double f = invN;
        if (q != 0.0) f *= exp(-q*(((i + 1) - jc)*dlnr + lnkr));
        out[(size_t)b0*N + i] = v.x*f;
        if (two) out[(size_t)b1*N + i] = v.y*f;
    }
}

// ------------------------------------------------- register-tiled kernel (N = 1024 .. 8192)
// Stockham autosort FFT with 16 complex values per thread: every pass is a radix-16 (8, 4, 2 for the remainder)
// butterfly in registers, so N = 4096 needs 3 + 3 shared-memory exchanges instead of the 12 + 12 passes of the
// radix-2 kernel above.  One CTA of N/16 threads transforms two real sequences packed as one complex one;
// twiddles exp(-2 pi i j / N), j < N/2, come from a table built once per plan.
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }

// (tr + i ti) * exp(SIGN 2 pi i e / 16), e = 0 .. 7 known at compile time after unrolling
template <int SIGN>
__device__ __forceinline__ double2 mul_w16(double tr, double ti, int e)
{
    const double C[8] = {1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977,
                         0.0, -0.38268343236508977, -0.70710678118654752, -0.92387953251128674};
    const double S[8] = {0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674,
                         1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977};
    if (e == 0) return make_double2(tr, ti);
    if (e == 4) return SIGN < 0 ? make_double2(ti, -tr) : make_double2(-ti, tr);
    const double c = C[e], sn = SIGN*S[e];
    return make_double2(tr*c - ti*sn, tr*sn + ti*c);
}

// in-register DFT of R = 2, 4, 8, 16 values, natural order in and out; SIGN = -1 forward, +1 inverse
template <int R, int SIGN>
__device__ __forceinline__ void dft_regs(double2* v)
{
#pragma unroll
    for (int len = R; len >= 2; len >>= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int b = 0; b < R; b += len) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const double2 u = v[b + j], w = v[b + j + half];
                v[b + j] = make_double2(u.x + w.x, u.y + w.y);
                v[b + j + half] = mul_w16<SIGN>(u.x - w.x, u.y - w.y, j*(16/len));
            }
        }
    }
    // decimation in frequency leaves bit-reversed order
#pragma unroll
    for (int i = 0; i < R; i++) {
        int rev = 0;
#pragma unroll
        for (int bit = 1, t = i; bit < R; bit <<= 1, t >>= 1) rev = (rev << 1) | (t & 1);
        if (i < rev) { const double2 tmp = v[i]; v[i] = v[rev]; v[rev] = tmp; }
    }
}

// shared-memory index with one pad element per 16: the radix-16 scatter of the first pass (positions 16 t + r)
// and the stride-N/16 gathers then fall on distinct banks
__device__ __forceinline__ int zpad(int i) { return i + (i >> 4); }

// one Stockham pass of radix R over z[N] (in place through registers); Ns = product of the previous radices
template <int R, int SIGN>
__device__ __forceinline__ void fft_pass(double2* __restrict__ z, const double2* __restrict__ tw, int N, int Ns)
{
    constexpr int ITEMS = 16/R;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int stride = N/R;
    const int tunit = N/(Ns*R);
    double2 v[16];
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const int j = tid + it*nthr;
        const int k = j & (Ns - 1);
#pragma unroll
        for (int r = 0; r < R; r++) v[it*R + r] = z[zpad(j + r*stride)];
        if (Ns > 1) {
            // twiddles exp(SIGN 2 pi i r k / (Ns R)), r = 1 .. R-1: one table read, the powers by multiplication
            // (depth <= 4 products: a few ulp)
            double2 w[R];
            const int idx = k*tunit;                 // < N/R <= N/2
            w[1] = tw[idx];
            if (SIGN > 0) w[1].y = -w[1].y;
#pragma unroll
            for (int r = 2; r < R; r++) {
                // r = 2^p: square of r/2; otherwise (highest power of two below r) x (remainder)
                int hp = 1;
                while (hp*2 <= r) hp *= 2;
                w[r] = (hp == r) ? cmul(w[r/2], w[r/2]) : cmul(w[hp], w[r - hp]);
            }
#pragma unroll
            for (int r = 1; r < R; r++) v[it*R + r] = cmul(v[it*R + r], w[r]);
        }
        dft_regs<R, SIGN>(v + it*R);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const int j = tid + it*nthr;
        const int k = j & (Ns - 1);
        const int j0 = (j - k)*R + k;
#pragma unroll
        for (int r = 0; r < R; r++) z[zpad(j0 + r*Ns)] = v[it*R + r];
    }
    __syncthreads();
}

template <int SIGN>
__device__ __forceinline__ void fft_all(double2* z, const double2* tw, int N, int r0, int r1, int r2, int r3)
{
    int Ns = 1;
    const int rad[4] = {r0, r1, r2, r3};
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int R = rad[s];
        if (R == 16) fft_pass<16, SIGN>(z, tw, N, Ns);
        else if (R == 8) fft_pass<8, SIGN>(z, tw, N, Ns);
        else if (R == 4) fft_pass<4, SIGN>(z, tw, N, Ns);
        else if (R == 2) fft_pass<2, SIGN>(z, tw, N, Ns);
        Ns *= (R > 0 ? R : 1);
    }
}

// WIN: the sequence is a window of a longer zero-padded one (mcfit-style transforms, transfer.cu): element i of the
// n_io inputs of row b sits at in[b*in_stride + i], is multiplied by pre[i] and placed at slot off_in + i; outputs
// are slots off_out .. off_out + n_io - 1, multiplied by post[i]; no power-law bias.
struct FFTWindow {
    const double* pre;
    const double* post;
    int n_io, off_in, off_out;
    long long in_stride, out_stride;
};

template <bool WIN>
__global__ void __launch_bounds__(512)
k_fftlog_r16(int N, int r0, int r1, int r2, int r3, int ntw, int batch, double q, double dlnr, double lnkr,
             const double* __restrict__ ure, const double* __restrict__ uim, const double2* __restrict__ twg,
             const double* __restrict__ in, double* __restrict__ out, FFTWindow win)
{
    extern __shared__ __align__(16) unsigned char smem_raw3[];
    double2* z = reinterpret_cast<double2*>(smem_raw3);
    double2* tw = z + N + N/16;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b0 = 2*blockIdx.x, b1 = b0 + 1;
    const bool two = b1 < batch;
    const double jc = 0.5*(N + 1);
    for (int j = tid; j < ntw; j += nt) tw[j] = __ldg(twg + j);      // a pass of radix R reads entries < N/R only
    for (int i = tid; i < N; i += nt) {
        double a0, a1;
        if (WIN) {
            const int j = i - win.off_in;
            const bool inside = j >= 0 && j < win.n_io;
            const double f = inside ? win.pre[j] : 0.0;
            a0 = inside ? f*in[(size_t)b0*win.in_stride + j] : 0.0;
            a1 = (inside && two) ? f*in[(size_t)b1*win.in_stride + j] : 0.0;
        } else {
            a0 = in[(size_t)b0*N + i];
            a1 = two ? in[(size_t)b1*N + i] : 0.0;
            if (q != 0.0) {
                const double f = exp(-q*((i + 1) - jc)*dlnr);      // fht: a(r) = A(r) (r/rc)^(-q)
                a0 *= f; a1 *= f;
            }
        }
        z[zpad(i)] = make_double2(a0, a1);
    }
    __syncthreads();
    fft_all<-1>(z, tw, N, r0, r1, r2, r3);
    // multiply by U_m (natural order)
    for (int m = tid; m < N; m += nt) {
        double ur, ui;
        if (m <= N/2) { ur = ure[m]; ui = uim[m]; }
        else { ur = ure[N - m]; ui = -uim[N - m]; }
        const double2 v = z[zpad(m)];
        z[zpad(m)] = make_double2(v.x*ur - v.y*ui, v.x*ui + v.y*ur);
    }
    __syncthreads();
    fft_all<+1>(z, tw, N, r0, r1, r2, r3);
    // reversal, 1/N, output bias (fhtq :772-782, fht :632-638)
    const double invN = 1.0/N;
    if (WIN) {
        for (int j = tid; j < win.n_io; j += nt) {
            const double2 v = z[zpad(N - 1 - (win.off_out + j))];
            const double f = invN*win.post[j];
            out[(size_t)b0*win.out_stride + j] = v.x*f;
            if (two) out[(size_t)b1*win.out_stride + j] = v.y*f;
        }
        return;
    }
    for (int i = tid; i < N; i += nt) {
        const double2 v = z[zpad(N - 1 - i)];
        double f = invN;
        if (q != 0.0) f *= exp(-q*(((i + 1) - jc)*dlnr + lnkr));
        out[(size_t)b0*N + i] = v.x*f;
        if (two) out[(size_t)b1*N + i] = v.y*f;
    }
}

// ------------------------------------------------------------ generic kernel
// out[j] = sum_l a_l g[(N-1-j-l) mod N]   (one CTA per sequence, a and g in shared memory)
__global__ void __launch_bounds__(256)
k_fftlog_direct(int N, double q, double dlnr, double lnkr, const double* __restrict__ g,
                const double* __restrict__ in, double* __restrict__ out)
{
    extern __shared__ double sm_d[];
    double* a = sm_d;
    double* gs = sm_d + N;
    const int tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x;
    const double jc = 0.5*(N + 1);
    for (int i = tid; i < N; i += nt) {
        double v = in[(size_t)b*N + i];
        if (q != 0.0) v *= exp(-q*((i + 1) - jc)*dlnr);
        a[i] = v;
        gs[i] = g[i];
    }
    __syncthreads();
    for (int j = tid; j < N; j += nt) {
        double s = 0.0;
        int d = N - 1 - j;                   // index (N-1-j-l) mod N, decreasing with l
        for (int l = 0; l < N; l++) {
            s = fma(a[l], gs[d], s);
            d = (d == 0) ? N - 1 : d - 1;
        }
        if (q != 0.0) s *= exp(-q*(((j + 1) - jc)*dlnr + lnkr));
        out[(size_t)b*N + j] = s;
    }
}

void fftlog_apply(Context& ctx, FFTLogPlan& p, int batch, const double* d_in, double* d_out, const FFTLogWindow* win)
{
    if (batch <= 0) return;
    const int N = p.N;
    ProfScope prof(ctx, PROF_FFTLOG);
    if (p.pow2 && N >= 1024) {
        // radices: as many 16s as possible, then one of 8 / 4 / 2
        int rad[4] = {0, 0, 0, 0}, ns = 0, rest = N;
        while (rest >= 16) { rad[ns++] = 16; rest /= 16; }
        if (rest > 1) rad[ns++] = rest;
        if (!p.tw.n) {
            std::vector<double> t((size_t)N);
            for (int j = 0; j < N/2; j++) { t[2*j] = std::cos(-2.0*M_PI*j/N); t[2*j + 1] = std::sin(-2.0*M_PI*j/N); }
            p.tw.upload(t, ctx.stream);
            ctx.sync();
        }
        const int ntw = N/rad[ns - 1];                 // the smallest radix comes last
        const size_t smem = (size_t)(N + N/16)*sizeof(double2) + (size_t)ntw*sizeof(double2);
        PRSD_REQUIRE(smem <= ctx.smem_optin, "FFTLog: N = %d needs %zu B of shared memory", N, smem);
        if (win) {
            FFTWindow fw{win->pre, win->post, win->n_io, win->off_in, win->off_out, win->in_stride, win->out_stride};
            PRSD_CUDA(cudaFuncSetAttribute(k_fftlog_r16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fftlog_r16<true><<<(batch + 1)/2, N/16, smem, ctx.stream>>>(N, rad[0], rad[1], rad[2], rad[3], ntw, batch, 0.0, p.dlnr, 0.0,
                p.ure.p, p.uim.p, reinterpret_cast<const double2*>(p.tw.p), d_in, d_out, fw);
        } else {
            PRSD_CUDA(cudaFuncSetAttribute(k_fftlog_r16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fftlog_r16<false><<<(batch + 1)/2, N/16, smem, ctx.stream>>>(N, rad[0], rad[1], rad[2], rad[3], ntw, batch, p.q, p.dlnr,
                std::log(p.kr), p.ure.p, p.uim.p, reinterpret_cast<const double2*>(p.tw.p), d_in, d_out, FFTWindow{});
        }
    } else if (p.pow2) {
        PRSD_REQUIRE(!win, "FFTLog: windowed transforms need a power-of-two N in [1024, 8192]");
        int logN = 0;
        while ((1 << logN) < N) logN++;
        const size_t smem = (size_t)N*sizeof(double2) + (size_t)(N/2)*sizeof(double2);
        PRSD_REQUIRE(smem <= ctx.smem_optin, "FFTLog: N = %d needs %zu B of shared memory", N, smem);
        PRSD_CUDA(cudaFuncSetAttribute(k_fftlog_pow2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fftlog_pow2<<<(batch + 1)/2, 256, smem, ctx.stream>>>(N, logN, batch, p.q, p.dlnr, std::log(p.kr),
                                                              p.ure.p, p.uim.p, d_in, d_out);
    } else {
        PRSD_REQUIRE(!win, "FFTLog: windowed transforms need a power-of-two N in [1024, 8192]");
        p.build_g(ctx);
        PRSD_REQUIRE(d_in != d_out, "FFTLog: in-place transform needs a power-of-two length");
        const size_t smem = (size_t)2*N*sizeof(double);
        if (smem > 48*1024) PRSD_CUDA(cudaFuncSetAttribute(k_fftlog_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fftlog_direct<<<batch, 256, smem, ctx.stream>>>(N, p.q, p.dlnr, std::log(p.kr), p.g.p, d_in, d_out);
    }
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
