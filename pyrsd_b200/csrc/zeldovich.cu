// zeldovich.cu -- Zel'dovich spectra (see zeldovich.cuh).
#include "zeldovich.cuh"
#include "fftlog_host.h"
#include "plan.cuh"

#include <cmath>

namespace prsd {

__global__ void k_zel_fold(int nk, const double* g, const int* ja, const int* jb, const double* wa, const double* wb, double* G);

void ZelPlan::build(Context& ctx, const double* kin, int nk_)
{
    PRSD_REQUIRE(nk_ > 0, "ZelPlan: empty k list");
    nk = nk_;
    k.assign(kin, kin + nk);
    for (int i = 0; i < nk; i++) PRSD_REQUIRE(k[i] > 0 && std::isfinite(k[i]), "ZelPlan: k[%d] = %g must be > 0", i, k[i]);
    const int N = ZEL_N;
    // grid constants of ZeldovichPS::InitializeR (ZeldovichPS.cpp:41-53)
    const double nc = 0.5*(N + 1), logrmin = -2.0, logrmax = 5.0;
    const double logrc = 0.5*(logrmin + logrmax), dlogr = (logrmax - logrmin)/N;
    const double dlnr = dlogr*std::log(10.0);

    g.alloc((size_t)(ZEL_NMAX + 1)*N);
    std::vector<int> hja((size_t)nk*16), hjb((size_t)nk*16);
    std::vector<double> hwa((size_t)nk*16), hwb((size_t)nk*16);
    std::vector<double> kmesh(N);
    for (int n = 0; n <= ZEL_NMAX; n++) {
        // FortranFFTLog(NUM_PTS, dlogr*log(10.), mu = n + 1/2, q = 0, kr = 1, kropt = 1)
        auto fp = fftlog_plan(ctx, N, 0.5 + n, 0.0, dlnr, 1.0, 1);
        fp->build_g(ctx);
        PRSD_CUDA(cudaMemcpyAsync(g.p + (size_t)n*N, fp->g.p, N*sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
        const double logkc = std::log10(fp->kr) - logrc;
        for (int j = 1; j <= N; j++) kmesh[j - 1] = std::pow(10.0, logkc + (j - nc)*dlogr);
        for (int i = 0; i < nk; i++) {
            // nearest_interp_1d (ZeldovichPS.cpp:58-82)
            int kn = 0;
            double d = std::fabs(k[i] - kmesh[0]);
            for (int j = 1; j < N; j++) {
                const double d2 = std::fabs(k[i] - kmesh[j]);
                if (d2 < d) { kn = j; d = d2; }
            }
            int l = kn;
            if (k[i] > kmesh[kn]) l = kn + 1;
            else if (k[i] < kmesh[kn]) l = kn - 1;
            double w_kn = 1.0, w_l = 0.0;
            if (l != kn && l >= 0 && l < N) {
                w_kn = (k[i] - kmesh[l])/(kmesh[kn] - kmesh[l]);
                w_l = 1.0 - w_kn;
            } else l = kn;
            hja[(size_t)i*16 + n] = kn; hwa[(size_t)i*16 + n] = w_kn;
            hjb[(size_t)i*16 + n] = l;  hwb[(size_t)i*16 + n] = w_l;
        }
    }
    ja.upload(hja, ctx.stream); jb.upload(hjb, ctx.stream);
    wa.upload(hwa, ctx.stream); wb.upload(hwb, ctx.stream);
    dk.upload(k, ctx.stream);
    r.upload(zeldovich_rgrid(), ctx.stream);
    G.alloc((size_t)nk*16*N);
    k_zel_fold<<<dim3((N + 255)/256, nk, 16), 256, 0, ctx.stream>>>(nk, g.p, ja.p, jb.p, wa.p, wb.p, G.p);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    ctx.sync();
}

// G[ik][n][l] = w_a g_n(N-1-j_a-l) + w_b g_n(N-1-j_b-l): the circular FFTLog kernel of order n + 1/2 folded with
// the reference's two-point output interpolation for this k (cosmology independent, built once per k list)
__global__ void k_zel_fold(int nk, const double* __restrict__ g, const int* __restrict__ ja, const int* __restrict__ jb,
                           const double* __restrict__ wa, const double* __restrict__ wb, double* __restrict__ G)
{
    const int ik = blockIdx.y, n = blockIdx.z, N = ZEL_N;
    const int l = blockIdx.x*blockDim.x + threadIdx.x;
    if (l >= N) return;
    int d1 = N - 1 - ja[(size_t)ik*16 + n] - l; if (d1 < 0) d1 += N;
    int d2 = N - 1 - jb[(size_t)ik*16 + n] - l; if (d2 < 0) d2 += N;
    const double* gn = g + (size_t)n*N;
    G[((size_t)ik*16 + n)*N + l] = wa[(size_t)ik*16 + n]*gn[d1] + wb[(size_t)ik*16 + n]*gn[d2];
}

// One CTA per (tile of ZW output rows, k): every thread owns r_l for l = tid (mod 256), keeps the 16 folded kernel
// values of that l in registers and reuses them for the ZW rows.  For n >= 1 the three spectra need
//   S_p = sum_n n^p rho^n G_n,  p = 0, 1, 2,   rho = k Y / r
// (Fsec of ZeldovichPS.cpp:198-203, 241-246, 285-296 expanded in powers of n), evaluated by Horner.
static const int ZW = 8;      // 4 rows per CTA at two CTAs per SM (128 registers) measured the same: not occupancy-bound

// ONLY00: P00 alone (the Jennings fits need P00 at a second normalisation, power_dm.py:676-718); P01, P11 are
// written as 0
template <bool ONLY00>
__global__ void __launch_bounds__(256)
k_zeldovich(int nk, int nrow, const double* __restrict__ kk, const double* __restrict__ rr, const double* __restrict__ G,
            const int* __restrict__ cmap,
            const double* __restrict__ X0, const double* __restrict__ XXin, const double* __restrict__ YY,
            const double* __restrict__ sigmasq, int ss_stride, const double* __restrict__ ratio, int lowk, double k0_low,
            const double* __restrict__ plin, const double* __restrict__ q3coef, int q3_stride, double* __restrict__ out)
{
    // w: output row ("walker"); c: row of the X0 / YY / sigma^2 / Q3 inputs it reads
    const int w0 = blockIdx.x*ZW, ik = blockIdx.y, tid = threadIdx.x;
    const int N = ZEL_N;
    const double k = kk[ik];
    const double k2 = k*k, k4 = k2*k2;

    if (lowk && k < k0_low) {
        if (tid < ZW && w0 + tid < nrow) {
            const int w = w0 + tid, c = cmap ? cmap[w] : w;
            const double rat = ratio ? ratio[w] : 1.0;
            const double s2 = rat*sigmasq[(size_t)c*ss_stride];
            double* o = out + (size_t)w*3*nk;
            const double P = plin[(size_t)w*nk + ik];
            const double Q3 = k4*q3coef[(size_t)c*q3_stride]*rat*rat;
            o[ik]        = (1.0 - k2*s2 + 0.5*k4*s2*s2)*P + 0.5*Q3;      // ZeldovichPS.cpp:205-208
            o[nk + ik]   = (1.0 - 2.0*k2*s2)*P + Q3;                     // :229-232
            o[2*nk + ik] = (1.0 - 3.0*k2*s2)*P + 1.5*Q3;                 // :262-265
        }
        return;
    }

    __shared__ int s_c[ZW];
    __shared__ double s_rat[ZW], s_s2[ZW], s_E0[ZW];
    __shared__ double s_red[8][3*ZW];
    if (tid < ZW) {
        const int w = min(w0 + tid, nrow - 1);
        const int c = cmap ? cmap[w] : w;
        const double rat = ratio ? ratio[w] : 1.0;
        s_c[tid] = c; s_rat[tid] = rat; s_s2[tid] = rat*sigmasq[(size_t)c*ss_stride];
        s_E0[tid] = exp(-k2*s_s2[tid]);
    }
    __syncthreads();

    double acc[3*ZW];
#pragma unroll
    for (int i = 0; i < 3*ZW; i++) acc[i] = 0.0;
    const double* Gk = G + (size_t)ik*16*N;
    for (int l = tid; l < N; l += 256) {
        double gl[16];
#pragma unroll
        for (int n = 0; n < 16; n++) gl[n] = __ldg(Gk + (size_t)n*N + l);
        const double r = rr[l];
        const double r15 = r*sqrt(r);
        const double kr = k/r;
#pragma unroll
        for (int j = 0; j < ZW; j++) {
            const int c = s_c[j];
            const double rat = s_rat[j], s2 = s_s2[j];
            const double x0 = rat*X0[(size_t)c*N + l];
            const double yy = rat*YY[(size_t)c*N + l];
            const double xx = XXin ? rat*XXin[(size_t)c*N + l] : x0 + 2.0*s2;
            const double E0 = s_E0[j];
            const double E = exp(-0.5*k2*(xx + yy));
            const double kxx = k2*xx, kyy = k2*yy;
            const double rho = kr*yy;
            const double g0 = r15*gl[0];
            const double pre = r15*E*rho;
            if (ONLY00) {
                double S0 = 0.0;
#pragma unroll
                for (int n = ZEL_NMAX; n >= 1; n--) S0 = fma(S0, rho, gl[n]);
                acc[3*j] += g0*(E - E0) + pre*S0;
                continue;
            }
            const double common = k4*xx*xx - 2.0*k2*x0;      // k^4 XX^2 - 2 k^2 X0
            const double c11 = common + 2.0*(kxx - 1.0)*kyy + k4*yy*yy;
            // n = 0 (Fprim of the three spectra, ZeldovichPS.cpp:190-196, 234-239, 273-283)
            double a00 = g0*(E - E0);
            double a01 = g0*k2*((xx + yy)*E - 2.0*s2*E0);
            double a11 = g0*(c11*E - 4.0*k4*s2*s2*E0);
            // n >= 1: v_n = r^1.5 E rho^n G_n;  P00 += v_n, P01 += v_n (kxx + kyy - 2n),
            //         P11 += v_n (c11 - 4 n (kxx - 1) - 4 n kyy + 4 n (n - 1))
            // S0 = sum_n G_n rho^(n-1), S1 = sum_n n G_n rho^(n-1), S2 = sum_n n^2 G_n rho^(n-1) from ONE Horner
            // recurrence carrying the polynomial P = S0 and its derivatives (P1 = P', P2 = P''/2):
            // S1 = (rho P)' = P + rho P',  S2 = S1 + rho (rho P)'' = P + 3 rho P' + rho^2 P''
            double S0 = 0.0, P1 = 0.0, P2 = 0.0;
#pragma unroll
            for (int n = ZEL_NMAX; n >= 1; n--) {
                P2 = fma(P2, rho, P1);
                P1 = fma(P1, rho, S0);
                S0 = fma(S0, rho, gl[n]);
            }
            const double rP1 = rho*P1;
            double S1 = S0 + rP1;
            double S2 = fma(2.0*rho, fma(rho, P2, P1), S1);
            S0 *= pre; S1 *= pre; S2 *= pre;
            a00 += S0;
            a01 += (kxx + kyy)*S0 - 2.0*S1;
            a11 += c11*S0 - 4.0*(kxx + kyy)*S1 + 4.0*S2;
            acc[3*j] += a00; acc[3*j + 1] += a01; acc[3*j + 2] += a11;
        }
    }
    // block reduction
#pragma unroll
    for (int i = 0; i < 3*ZW; i++) {
        double v = acc[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if ((tid & 31) == 0) s_red[tid >> 5][i] = v;
    }
    __syncthreads();
    if (tid < 3*ZW) {
        const int j = tid/3, which = tid - 3*j;
        if (w0 + j < nrow) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += s_red[w][tid];
            const double pre = sqrt(0.5*M_PI)/(k*sqrt(k));
            // ZeldovichP00 / P01 / P11::Evaluate (:182-188, :222-227, :255-261)
            const double fac = which == 0 ? 4.0*M_PI : (which == 1 ? -2.0*M_PI : M_PI);
            out[(size_t)(w0 + j)*3*nk + which*nk + ik] = fac*pre*t;
        }
    }
}

void zeldovich_eval(Context& ctx, const ZelPlan& plan, int ncosmo, const double* X0, const double* YY,
                    const double* sigmasq, int sigmasq_stride, const double* ratio, bool lowk, double k0_low,
                    const double* plin, const double* q3coef, int q3_stride, double* out)
{
    zeldovich_eval_map(ctx, plan, ncosmo, nullptr, X0, nullptr, YY, sigmasq, sigmasq_stride, ratio, lowk, k0_low, plin, q3coef, q3_stride, out);
}

void zeldovich_eval_xx(Context& ctx, const ZelPlan& plan, int ncosmo, const double* X0, const double* XX, const double* YY,
                       const double* sigmasq, int sigmasq_stride, const double* ratio, bool lowk, double k0_low,
                       const double* plin, const double* q3coef, int q3_stride, double* out)
{
    zeldovich_eval_map(ctx, plan, ncosmo, nullptr, X0, XX, YY, sigmasq, sigmasq_stride, ratio, lowk, k0_low, plin, q3coef, q3_stride, out);
}

void zeldovich_eval_map(Context& ctx, const ZelPlan& plan, int nrow, const int* cmap, const double* X0, const double* XX,
                        const double* YY, const double* sigmasq, int sigmasq_stride, const double* ratio, bool lowk,
                        double k0_low, const double* plin, const double* q3coef, int q3_stride, double* out, bool only_p00)
{
    if (nrow <= 0) return;
    PRSD_REQUIRE(!lowk || (plin && q3coef), "zeldovich_eval: low-k approximation needs P_lin(k) and Q3");
    dim3 grid((nrow + ZW - 1)/ZW, plan.nk);
    ProfScope prof(ctx, PROF_ZELDOVICH);
    if (only_p00)
        k_zeldovich<true><<<grid, 256, 0, ctx.stream>>>(plan.nk, nrow, plan.dk.p, plan.r.p, plan.G.p, cmap, X0, XX, YY, sigmasq, sigmasq_stride, ratio, lowk ? 1 : 0, k0_low, plin, q3coef, q3_stride, out);
    else
        k_zeldovich<false><<<grid, 256, 0, ctx.stream>>>(plan.nk, nrow, plan.dk.p, plan.r.p, plan.G.p, cmap, X0, XX, YY, sigmasq, sigmasq_stride, ratio, lowk ? 1 : 0, k0_low, plin, q3coef, q3_stride, out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
