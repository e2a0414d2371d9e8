// gp.cu -- mean prediction of the reference's Gaussian-process simulation fits (pyRSD/rsd/simulation.py:16-256:
// GeorgeSimulationData on george.GP with an ExpSquaredKernel) on the device.
//
//   y(x*) = y_mean + y_scale * sum_i alpha_i amp exp(-1/2 sum_d (z*_d - Z_id)^2),   z = (x - x_mean) / x_scale / sqrt(metric)
//
// alpha = (K + diag(yerr^2))^-1 y_scaled depends on the training set only and is shipped with the package
// (pyrsd_b200/data/make_simulation_fits.py); a prediction is one kernel-vector product.  One CTA per prediction point,
// its threads stride over the training points (Z stored [dim][n]: coalesced), tree reduction in shared memory.  The
// stochasticity of a batch of walkers (10 bias pairs x 20 k each against 11 550 training points) is a single launch.
#include "gp.cuh"

namespace prsd {

static const int GP_THREADS = 256;
static const int GP_MAXDIM = 8;

struct GPDims { double mean[GP_MAXDIM], inv[GP_MAXDIM]; };     // inv = 1 / (x_scale sqrt(metric))

__global__ void __launch_bounds__(GP_THREADS)
k_gp_predict(int n, int nd, const double* __restrict__ Z /*[nd][n]*/, const double* __restrict__ alpha, GPDims D, double amp,
             double y_mean, double y_scale, const double* __restrict__ xpred /*[npred][nd]*/, double* __restrict__ out)
{
    __shared__ double red[GP_THREADS];
    const int p = blockIdx.x;
    double z[GP_MAXDIM];
    for (int d = 0; d < nd; d++) z[d] = (xpred[(size_t)p*nd + d] - D.mean[d])*D.inv[d];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += GP_THREADS) {
        double r2 = 0.0;
        for (int d = 0; d < nd; d++) { const double t = z[d] - Z[(size_t)d*n + i]; r2 = fma(t, t, r2); }
        acc = fma(alpha[i], exp(-0.5*r2), acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = GP_THREADS/2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[p] = y_mean + y_scale*amp*red[0];
}

GPModel::GPModel(Context& c, int n_, int nd_, const double* x, const double* x_mean, const double* x_scale, const double* metric,
                 double amp_, const double* alpha_, double y_mean_, double y_scale_)
    : ctx(&c), n(n_), nd(nd_), amp(amp_), y_mean(y_mean_), y_scale(y_scale_)
{
    PRSD_REQUIRE(n > 0 && nd > 0 && nd <= GP_MAXDIM, "GaussianProcess: %d training points with %d inputs unsupported", n, nd);
    for (int d = 0; d < nd; d++) {
        PRSD_REQUIRE(x_scale[d] > 0 && metric[d] > 0, "GaussianProcess: input %d has a non-positive scale or metric", d);
        mean[d] = x_mean[d];
        inv[d] = 1.0/(x_scale[d]*std::sqrt(metric[d]));
    }
    std::vector<double> Zt((size_t)n*nd);
    for (int i = 0; i < n; i++)
        for (int d = 0; d < nd; d++) Zt[(size_t)d*n + i] = (x[(size_t)i*nd + d] - mean[d])*inv[d];
    Z.upload(Zt, c.stream);
    alpha.upload(alpha_, n, c.stream);
    c.sync();
}

void GPModel::predict_device(int npred, const double* d_x, double* d_out)
{
    if (npred <= 0) return;
    GPDims D;
    for (int d = 0; d < GP_MAXDIM; d++) { D.mean[d] = d < nd ? mean[d] : 0.0; D.inv[d] = d < nd ? inv[d] : 0.0; }
    k_gp_predict<<<npred, GP_THREADS, 0, ctx->stream>>>(n, nd, Z.p, alpha.p, D, amp, y_mean, y_scale, d_x, d_out);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
}

void GPModel::predict(int npred, const double* x, double* out)
{
    if (npred <= 0) return;
    DevBuf<double> dx, dout((size_t)npred);
    dx.upload(x, (size_t)npred*nd, ctx->stream);
    predict_device(npred, dx.p, dout.p);
    dout.download(out, (size_t)npred, ctx->stream);
    ctx->sync();
}

} // namespace prsd
