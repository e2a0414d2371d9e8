// zeldovich.cu -- Zel'dovich spectra (see zeldovich.cuh).
#include "zeldovich.cuh"
#include "fftlog_host.h"
#include "plan.cuh"

#include <cmath>

namespace prsd {

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
    ctx.sync();
}

// one CTA per (k, cosmology)
__global__ void __launch_bounds__(256)
k_zeldovich(int nk, const double* __restrict__ kk, const double* __restrict__ rr, const double* __restrict__ g,
            const int* __restrict__ ja, const int* __restrict__ jb, const double* __restrict__ wa, const double* __restrict__ wb,
            const int* __restrict__ cmap,
            const double* __restrict__ X0, const double* __restrict__ XXin, const double* __restrict__ YY,
            const double* __restrict__ sigmasq, int ss_stride, const double* __restrict__ ratio, int lowk, double k0_low,
            const double* __restrict__ plin, const double* __restrict__ q3coef, int q3_stride, double* __restrict__ out)
{
    // w: output row ("walker"); c: row of the X0 / YY / sigma^2 / Q3 inputs it reads
    const int ik = blockIdx.x, w = blockIdx.y, tid = threadIdx.x;
    const int c = cmap ? cmap[w] : w;
    const int N = ZEL_N;
    const double k = kk[ik];
    const double rat = ratio ? ratio[w] : 1.0;
    const double s2 = rat*sigmasq[(size_t)c*ss_stride];
    const double k2 = k*k, k4 = k2*k2;
    double* o = out + (size_t)w*3*nk;

    if (lowk && k < k0_low) {
        if (tid == 0) {
            const double P = plin[(size_t)w*nk + ik];
            const double Q3 = k4*q3coef[(size_t)c*q3_stride]*rat*rat;
            o[ik]        = (1.0 - k2*s2 + 0.5*k4*s2*s2)*P + 0.5*Q3;      // ZeldovichPS.cpp:205-208
            o[nk + ik]   = (1.0 - 2.0*k2*s2)*P + Q3;                     // :229-232
            o[2*nk + ik] = (1.0 - 3.0*k2*s2)*P + 1.5*Q3;                 // :262-265
        }
        return;
    }

    __shared__ int s_ja[16], s_jb[16];
    __shared__ double s_wa[16], s_wb[16];
    __shared__ double s_red[3][8];
    if (tid < 16) {
        s_ja[tid] = ja[(size_t)ik*16 + tid]; s_jb[tid] = jb[(size_t)ik*16 + tid];
        s_wa[tid] = wa[(size_t)ik*16 + tid]; s_wb[tid] = wb[(size_t)ik*16 + tid];
    }
    __syncthreads();

    double a00 = 0.0, a01 = 0.0, a11 = 0.0;
    const double E0 = exp(-k2*s2);
    for (int l = tid; l < N; l += 256) {
        const double r = rr[l];
        const double x0 = rat*X0[(size_t)c*N + l];
        const double yy = rat*YY[(size_t)c*N + l];
        const double xx = XXin ? rat*XXin[(size_t)c*N + l] : x0 + 2.0*s2;
        const double E = exp(-0.5*k2*(xx + yy));
        const double r15 = r*sqrt(r);
        const double kxx = k2*xx, kyy = k2*yy;
        const double common = k4*xx*xx - 2.0*k2*x0;      // k^4 XX^2 - 2 k^2 X0
        // n = 0 (Fprim of the three spectra, ZeldovichPS.cpp:190-196, 234-239, 273-283)
        {
            int d1 = N - 1 - s_ja[0] - l; if (d1 < 0) d1 += N;
            int d2 = N - 1 - s_jb[0] - l; if (d2 < 0) d2 += N;
            const double G = s_wa[0]*g[d1] + s_wb[0]*g[d2];
            a00 += r15*(E - E0)*G;
            a01 += r15*k2*((xx + yy)*E - 2.0*s2*E0)*G;
            const double t1 = (common + 2.0*(kxx - 1.0)*kyy + k4*yy*yy)*E;
            const double t2 = -4.0*k4*s2*s2*E0;
            a11 += r15*(t1 + t2)*G;
        }
        // n >= 1 (Fsec, :198-203, 241-246, 285-296)
        const double rho = k*yy/r;
        double pw = r15*E;
#pragma unroll 5
        for (int n = 1; n <= ZEL_NMAX; n++) {
            pw *= rho;
            int d1 = N - 1 - s_ja[n] - l; if (d1 < 0) d1 += N;
            int d2 = N - 1 - s_jb[n] - l; if (d2 < 0) d2 += N;
            const double* gn = g + (size_t)n*N;
            const double v = pw*(s_wa[n]*gn[d1] + s_wb[n]*gn[d2]);
            const double dn = (double)n;
            a00 += v;
            a01 += v*(kxx + kyy - 2.0*dn);
            a11 += v*(common + 2.0*(kxx - 1.0)*(kyy - 2.0*dn) + k4*yy*yy - 4.0*k2*dn*yy + 4.0*dn*(dn - 1.0));
        }
    }
    // block reduction
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        a00 += __shfl_xor_sync(0xffffffffu, a00, s);
        a01 += __shfl_xor_sync(0xffffffffu, a01, s);
        a11 += __shfl_xor_sync(0xffffffffu, a11, s);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = a00; s_red[1][tid >> 5] = a01; s_red[2][tid >> 5] = a11; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0, t1 = 0, t2 = 0;
        for (int w = 0; w < 8; w++) { t0 += s_red[0][w]; t1 += s_red[1][w]; t2 += s_red[2][w]; }
        const double pre = sqrt(0.5*M_PI)/(k*sqrt(k));
        o[ik]        = 4.0*M_PI*pre*t0;      // ZeldovichP00::Evaluate (:182-188)
        o[nk + ik]   = -2.0*M_PI*pre*t1;     // ZeldovichP01::Evaluate (:222-227)
        o[2*nk + ik] = M_PI*pre*t2;          // ZeldovichP11::Evaluate (:255-261)
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
                        double k0_low, const double* plin, const double* q3coef, int q3_stride, double* out)
{
    if (nrow <= 0) return;
    PRSD_REQUIRE(!lowk || (plin && q3coef), "zeldovich_eval: low-k approximation needs P_lin(k) and Q3");
    dim3 grid(plan.nk, nrow);
    ProfScope prof(ctx, PROF_ZELDOVICH);
    k_zeldovich<<<grid, 256, 0, ctx.stream>>>(plan.nk, plan.dk.p, plan.r.p, plan.g.p, plan.ja.p, plan.jb.p, plan.wa.p, plan.wb.p,
                                            cmap, X0, XX, YY, sigmasq, sigmasq_stride, ratio, lowk ? 1 : 0, k0_low, plin, q3coef, q3_stride, out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
