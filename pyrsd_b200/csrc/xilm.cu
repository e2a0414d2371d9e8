// xilm.cu -- xi_l^m(r) = \int dk/(2 pi^2) k^m j_l(kr) P(k) by FFTLog, batched over spectra.
//
// Device version of ComputeXiLM (FFTLOG branch) / ComputeXiLM_fftlog / pk_to_xi / xi_to_pk
// (pyRSD/_gcl/cpp/CorrelationFunction.cpp:23-141): resample P(k) onto an exactly log-spaced grid
// with the reference's cubic spline, a = k^{m-1/2} P exp(-(k R)^2), FFTLog with mu = l + 1/2,
// xi = (2 pi r)^{-3/2} a, spline back onto the requested r.
#include "xilm.cuh"
#include "pt_math.cuh"
#include "spectra.cuh"

#include <cmath>

namespace prsd {

__global__ void k_xilm_pre(int N, int nspl, const double* __restrict__ spl /*[batch][5 nspl]*/, const double* __restrict__ kgrid,
                           double mexp, double smoothing, double* __restrict__ a)
{
    const int b = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double k = kgrid[i];
    const double pk = cubic_spline_eval_bsearch(nspl, spl + (size_t)b*5*nspl, k);
    const double ks = k*smoothing;
    a[(size_t)b*N + i] = pow(k, mexp)*pk*exp(-ks*ks);
}

__global__ void k_xilm_pre_direct(int N, const double* __restrict__ pk, const double* __restrict__ kgrid,
                                  double mexp, double smoothing, double* __restrict__ a)
{
    const int b = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double k = kgrid[i];
    const double ks = k*smoothing;
    a[(size_t)b*N + i] = pow(k, mexp)*pk[(size_t)b*N + i]*exp(-ks*ks);
}

__global__ void k_xilm_post(int N, const double* __restrict__ rgrid, double* __restrict__ a)
{
    const int b = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double x = 2.0*M_PI*rgrid[i];
    a[(size_t)b*N + i] *= 1.0/(x*sqrt(x));
}

__global__ void k_spline_eval_rows(int nspl, const double* __restrict__ spl, int nq, const double* __restrict__ xq,
                                   double scale, double* __restrict__ out)
{
    const int b = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= nq) return;
    out[(size_t)b*nq + i] = scale*cubic_spline_eval_bsearch(nspl, spl + (size_t)b*5*nspl, xq[i]);
}

void spline_eval_host_x(Context& ctx, int batch, int n, const double* h_x, const double* d_y, int nq, const double* h_xq, double* d_out)
{
    PRSD_REQUIRE(n >= 4, "CubicSpline: need at least 4 points");
    DevBuf<double> dx, dq, spl((size_t)batch*5*n);
    dx.upload(h_x, n, ctx.stream);
    dq.upload(h_xq, nq, ctx.stream);
    spline_build_batched(ctx, batch, n, dx.p, 0, d_y, spl.p);
    k_spline_eval_rows<<<dim3((nq + 127)/128, batch), 128, 0, ctx.stream>>>(n, spl.p, nq, dq.p, 1.0, d_out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    ctx.sync();
}

void xilm_log_grid(int N, double xmin, double xmax, std::vector<double>& grid, double& dlogr, double& logrc)
{
    // CorrelationFunction.cpp:65-73
    const double nc = 0.5*(N + 1);
    const double lmin = std::log10(xmin), lmax = std::log10(xmax);
    logrc = 0.5*(lmin + lmax);
    dlogr = (lmax - lmin)/(N - 1);
    grid.resize(N);
    for (int i = 1; i <= N; i++) grid[i - 1] = std::pow(10.0, logrc + (i - nc)*dlogr);
}

void compute_xilm_fftlog_dev(Context& ctx, int batch, int l, int m, int N, const double* d_k, const double* h_k,
                             const double* d_pk, double q, double smoothing, std::vector<double>& r_host,
                             double* d_xi, double* kr_out)
{
    // ComputeXiLM_fftlog (CorrelationFunction.cpp:23-54)
    const double nc = 0.5*(N + 1);
    const double lmin = std::log10(h_k[0]), lmax = std::log10(h_k[N - 1]);
    const double dlogr = (lmax - lmin)/(N - 1), logrc = 0.5*(lmin + lmax);
    auto plan = fftlog_plan(ctx, N, l + 0.5, q, dlogr*std::log(10.0), 1.0, 1);
    dim3 grid((N + 127)/128, batch);
    k_xilm_pre_direct<<<grid, 128, 0, ctx.stream>>>(N, d_pk, d_k, m - 0.5, smoothing, d_xi);
    ctx.launches++;
    DevBuf<double> tmp;
    double* dst = d_xi;
    if (!plan->pow2) { tmp.alloc((size_t)batch*N); dst = tmp.p; }
    fftlog_apply(ctx, *plan, batch, d_xi, dst);
    if (dst != d_xi) PRSD_CUDA(cudaMemcpyAsync(d_xi, dst, (size_t)batch*N*sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    const double logkc = std::log10(plan->kr) - logrc;
    r_host.resize(N);
    for (int j = 1; j <= N; j++) r_host[j - 1] = std::pow(10.0, logkc + (j - nc)*dlogr);
    DevBuf<double> dr;
    dr.upload(r_host, ctx.stream);
    k_xilm_post<<<grid, 128, 0, ctx.stream>>>(N, dr.p, d_xi);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    if (kr_out) *kr_out = plan->kr;
    ctx.sync();
}

void compute_xilm_dev(Context& ctx, int batch, int l, int m, int nk, const double* h_k, const double* d_pk,
                      int nr, const double* h_r, double smoothing, double scale, double* d_out)
{
    // ComputeXiLM, FFTLOG branch (CorrelationFunction.cpp:56-90)
    PRSD_REQUIRE(nk >= 4 && nk <= 4096, "ComputeXiLM: %d input samples unsupported (4..4096: the batched spline build)", nk);
    for (int i = 1; i < nk; i++) PRSD_REQUIRE(h_k[i] > h_k[i - 1] && h_k[0] > 0, "ComputeXiLM: abscissae must be positive and increasing");
    std::vector<double> kg, rg;
    double dlogr, logrc;
    xilm_log_grid(nk, h_k[0], h_k[nk - 1], kg, dlogr, logrc);
    DevBuf<double> dkin, dkg, spl((size_t)batch*5*nk), a((size_t)batch*nk);
    dkin.upload(h_k, nk, ctx.stream);
    dkg.upload(kg, ctx.stream);
    spline_build_batched(ctx, batch, nk, dkin.p, 0, d_pk, spl.p);
    dim3 grid((nk + 127)/128, batch);
    k_xilm_pre<<<grid, 128, 0, ctx.stream>>>(nk, nk, spl.p, dkg.p, m - 0.5, smoothing, a.p);
    ctx.launches++;
    const double nc = 0.5*(nk + 1);
    auto plan = fftlog_plan(ctx, nk, l + 0.5, 0.0, dlogr*std::log(10.0), 1.0, 1);
    DevBuf<double> a2;
    double* dst = a.p;
    if (!plan->pow2) { a2.alloc((size_t)batch*nk); dst = a2.p; }
    fftlog_apply(ctx, *plan, batch, a.p, dst);
    const double logkc = std::log10(plan->kr) - logrc;
    rg.resize(nk);
    for (int j = 1; j <= nk; j++) rg[j - 1] = std::pow(10.0, logkc + (j - nc)*dlogr);
    DevBuf<double> drg, drq;
    drg.upload(rg, ctx.stream);
    drq.upload(h_r, nr, ctx.stream);
    k_xilm_post<<<grid, 128, 0, ctx.stream>>>(nk, drg.p, dst);
    ctx.launches++;
    // spline of (r', xi) evaluated at the requested r
    spline_build_batched(ctx, batch, nk, drg.p, 0, dst, spl.p);
    k_spline_eval_rows<<<dim3((nr + 127)/128, batch), 128, 0, ctx.stream>>>(nk, spl.p, nr, drq.p, scale, d_out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    ctx.sync();
}

} // namespace prsd
