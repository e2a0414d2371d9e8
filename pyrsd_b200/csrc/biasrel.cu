// biasrel.cu -- halo bias -> halo mass on the device: BiasToMassRelation / BiasToSigmaRelation of the reference
// (pyRSD/rsd/tools.py:393-649), the constraint behind `sigma_sB` of the default galaxy fit
// (rsdfit/theory/gal_defaults.py:174).
//
// The reference solves  bias_Tinker(rescaling * Sigma(R(M))) = b1  for M with scipy's brentq, every objective call
// integrating Sigma(R) adaptively (compute_fresh, tools.py:519-531), or reads a 100 x 70 table of such roots whose
// Sigma comes from a spline over R = logspace(-3, 4, 1000) (tools.py:463-507).  Here Sigma(R) of every cosmology of a
// batch is ONE application of a product-integration operator (linear.cu, 1000 outputs, cosmology independent), its
// cubic spline is built on the device, and every (sigma8_z, b1) pair is one thread running a bracketing root solve
// (bisection safeguarded regula falsi in ln M) to 1e-13: thousands of walkers per launch.
#include "biasrel.cuh"
#include "linear.cuh"
#include "pt_math.cuh"

#include <cmath>
#include <list>
#include <mutex>

namespace prsd {

static const int BR_NR = 1000;                         // R_interp of tools.py:470
static const double BR_LOG_RMIN = -3.0, BR_LOG_RMAX = 4.0;

__host__ __device__ inline double bias_tinker(double sigma, double delta_halo)
{
    // Tinker et al. 2010 (tools.py:693-711)
    const double delta_c = 1.686;
    const double y = log10(delta_halo);
    const double e4 = exp(-pow(4.0/y, 4.0));
    const double A = 1.0 + 0.24*y*e4, a = 0.44*y - 0.88, B = 0.183, b = 1.5, C = 0.019 + 0.107*y + 0.19*e4, c = 2.4;
    const double nu = delta_c/sigma;
    const double nua = pow(nu, a);
    return 1.0 - A*nua/(nua + pow(delta_c, a)) + B*pow(nu, b) + C*pow(nu, c);
}

__global__ void k_sqrt_rows(size_t n, double* __restrict__ v)
{
    const size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) v[i] = sqrt(v[i]);
}

__global__ void k_bias_to_mass(int nw, const int* __restrict__ cmap, const double* __restrict__ rescale, const double* __restrict__ b1,
                               const double* __restrict__ spl /*[nc][5 BR_NR]*/, double lnR0, double inv_dlnR, double rho_bar,
                               double delta_halo, double lnM_lo, double lnM_hi, double* __restrict__ mass, int* __restrict__ status)
{
    const int w = blockIdx.x*blockDim.x + threadIdx.x;
    if (w >= nw) return;
    const double* s = spl + (size_t)cmap[w]*5*BR_NR;
    const double resc = rescale[w], target = b1[w];
    auto objective = [&](double lnM) {
        const double R = cbrt(3.0*exp(lnM)/(4.0*M_PI*rho_bar));
        return bias_tinker(resc*cubic_spline_eval_log(BR_NR, s, lnR0, inv_dlnR, R), delta_halo) - target;
    };
    double a = lnM_lo, b = lnM_hi, fa = objective(a), fb = objective(b);
    if (!(fa*fb <= 0.0)) { mass[w] = nan(""); status[w] = 1; return; }      // brentq raises: f(a) and f(b) must have different signs
    // Illinois-safeguarded regula falsi with a bisection fallback: the bracket shrinks every step
    int side = 0;
    for (int it = 0; it < 200 && (b - a) > 1e-14*fmax(1.0, fabs(a)); it++) {
        double x = (fa*b - fb*a)/(fa - fb);
        if (!(x > a && x < b) || (it & 7) == 7) x = 0.5*(a + b);
        const double fx = objective(x);
        if (fx == 0.0) { a = b = x; break; }
        if (fx*fb > 0.0) { b = x; fb = fx; if (side == -1) fa *= 0.5; side = -1; }
        else             { a = x; fa = fx; if (side == +1) fb *= 0.5; side = +1; }
    }
    mass[w] = exp(0.5*(a + b));
    status[w] = 0;
}

namespace {
std::shared_ptr<LinearOp> sigma_operator(Context& c)
{
    // cosmology independent: kept per device
    static std::mutex mu;
    static std::list<std::pair<int, std::shared_ptr<LinearOp>>> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& e : cache) if (e.first == c.device) return e.second;
    std::vector<LinOutSpec> o(BR_NR);
    for (int i = 0; i < BR_NR; i++) {
        const double R = std::pow(10.0, BR_LOG_RMIN + (BR_LOG_RMAX - BR_LOG_RMIN)*i/(BR_NR - 1));
        o[i] = {LK_SIGMA_R, 0, R, 1e-5, 1e2, 1.0};                // PowerSpectrum::Sigma integrates [1e-5, 100] (PowerSpectrum.cpp:28-45)
    }
    auto op = std::make_shared<LinearOp>();
    op->build(c, o);
    cache.emplace_back(c.device, op);
    return op;
}
}

void sigma_R_table(SpectraBatch& sp, double* d_sigma /*[ncosmo][BR_NR]*/)
{
    Context& c = *sp.ctx;
    auto op = sigma_operator(c);
    std::vector<int> idx(sp.ncosmo);
    for (int i = 0; i < sp.ncosmo; i++) idx[i] = sp.table_index(i, SP_LIN);
    DevBuf<int> di;
    di.upload(idx, c.stream);
    op->apply(c, sp.tables.p, di.p, sp.ncosmo, d_sigma);
    const size_t n = (size_t)sp.ncosmo*BR_NR;
    k_sqrt_rows<<<(unsigned)((n + 255)/256), 256, 0, c.stream>>>(n, d_sigma);
    PRSD_CUDA(cudaGetLastError());
    c.launches++;
    PRSD_CUDA(cudaStreamSynchronize(c.stream));      // idx goes out of scope
}

std::vector<double> bias_relation_rgrid()
{
    std::vector<double> R(BR_NR);
    for (int i = 0; i < BR_NR; i++) R[i] = std::pow(10.0, BR_LOG_RMIN + (BR_LOG_RMAX - BR_LOG_RMIN)*i/(BR_NR - 1));
    return R;
}

void bias_to_mass(SpectraBatch& sp, int nw, const int* cmap, const double* rescale, const double* b1, double rho_bar,
                  double delta_halo, double mass_lo, double mass_hi, double* mass_out, double* sigma_out)
{
    Context& c = *sp.ctx;
    cudaStream_t s = c.stream;
    PRSD_REQUIRE(nw > 0 && rescale && b1 && mass_out, "bias_to_mass: empty request");
    PRSD_REQUIRE(rho_bar > 0 && delta_halo > 1 && mass_lo > 0 && mass_hi > mass_lo, "bias_to_mass: bad constants");
    const int nc = sp.ncosmo;
    std::vector<int> hmap(nw);
    for (int w = 0; w < nw; w++) {
        hmap[w] = cmap ? cmap[w] : (nc == 1 ? 0 : w);
        PRSD_REQUIRE(hmap[w] >= 0 && hmap[w] < nc, "bias_to_mass: walker %d maps to cosmology %d (batch has %d)", w, hmap[w], nc);
        PRSD_REQUIRE(rescale[w] > 0 && std::isfinite(b1[w]), "bias_to_mass: walker %d has a non-positive sigma8 ratio or a bad bias", w);
    }
    DevBuf<double> sig((size_t)nc*BR_NR), spl((size_t)nc*5*BR_NR), dR, dres, db1, dmass((size_t)nw);
    DevBuf<int> dmap, dstat((size_t)nw);
    sigma_R_table(sp, sig.p);
    if (sigma_out) sig.download(sigma_out, (size_t)nc*BR_NR, s);
    dR.upload(bias_relation_rgrid(), s);
    spline_build_batched(c, nc, BR_NR, dR.p, 0, sig.p, spl.p);
    dmap.upload(hmap, s);
    dres.upload(rescale, nw, s);
    db1.upload(b1, nw, s);
    const double lnR0 = BR_LOG_RMIN*std::log(10.0), inv = (BR_NR - 1)/((BR_LOG_RMAX - BR_LOG_RMIN)*std::log(10.0));
    k_bias_to_mass<<<(nw + 127)/128, 128, 0, s>>>(nw, dmap.p, dres.p, db1.p, spl.p, lnR0, inv, rho_bar, delta_halo,
                                                  std::log(mass_lo), std::log(mass_hi), dmass.p, dstat.p);
    PRSD_CUDA(cudaGetLastError());
    c.launches++;
    std::vector<int> stat(nw);
    dmass.download(mass_out, nw, s);
    dstat.download(stat.data(), nw, s);
    c.sync();
    for (int w = 0; w < nw; w++)
        PRSD_REQUIRE(stat[w] == 0, "bias_to_mass: no root for walker %d (b1 = %g, sigma8 ratio = %g) in [%g, %g] M_sun/h: "
                     "f(a) and f(b) must have different signs", w, b1[w], rescale[w], mass_lo, mass_hi);
}

} // namespace prsd
