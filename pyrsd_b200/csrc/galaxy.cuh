// galaxy.cuh -- GalaxySpectrum P(k, mu) assembly on the device.
//
// Device restatement of the reference's Python model layer for the default galaxy model:
//   dark-matter terms    pyRSD/rsd/power/dm/P00.py .. P22.py, power_dm.py:676-718, 848-884
//   halo (biased) terms  pyRSD/rsd/power/biased/P00.py .. P22.py, power_biased.py:325-525
//   HZPT broadband       pyRSD/rsd/hzpt/__init__.py:185-287, P01.py:57-104, P11.py:86-129, Phm.py:195-235
//   galaxy sum, FOG, AP  pyRSD/rsd/power/gal/__init__.py:24-40, core.py:159-233, fog_kernels.py:40-84,
//                        Pcc/__init__.py:39-60, rsd/tools.py:54-75, 166-222
// Three kernels per batch of walkers:
//   k_gal_base   : every cosmology-level quantity on the model k grid (one thread per k)
//   k_gal_moments: P[mu^0], P[mu^2], P[mu^4] (, P[mu^6]) of the 10 (b1, b1_bar) pairs and their
//                  not-a-knot splines (the reference's RSDSpline on self.k)
//   k_gal_pkmu   : AP remap, spline evaluation, FOG damping, weighted sum of the 10 sub-spectra
#pragma once
#include "plan.cuh"
#include "zeldovich.cuh"

namespace prsd {

static const int GAL_NPAR = 56;        // doubles per walker, see include/pyrsd_b200.h
static const int GAL_NPAIR = 10;
static const int GAL_NSTOCH = 20;      // GP_NK of power_biased.py:24
static const int GAL_NPAIROVR = 8;     // per-pair overrides of use_mean_bias: mean b1, 6 nonlinear biases, halo sigma_v
static const int GAL_NLOAD = 4;        // loadable dark-matter terms: Pdv, P00_mu0, P01_mu2, P11_mu4 (rsd/sim_loader.py:9)

struct GalaxyConfig {
    int max_mu;                 // 0, 2, 4 or 6
    int fog_model;              // 0 modified_lorentzian, 1 lorentzian, 2 gaussian
    int use_P00_model, use_P01_model, use_P11_model, use_Pdv_model, use_Pvv_model, use_Phm_model;
    int use_tidal_bias, use_so_correction;
    double k0_low;              // low-k transition of the Zel'dovich drivers (5e-3)
    int interpolate;            // 1: the Zel'dovich terms are read from sigma8 x k tables (InterpolatedHZPTModels, the
                                // reference default), 0: evaluated exactly at every call
    double hz00[10];            // A0(amp, alpha) R R1 R1h R2h of HaloZeldovichBase.default_parameters
    double hz11[9];             // A0(amp, alpha, beta) R R1h of HaloZeldovichP11.default_parameters
    double hzhm[20];            // A0 R1 R1h R2h R (amp, alpha, beta, run) of HaloMatterBase.default_parameters
};
void galaxy_config_default(GalaxyConfig& c);

// launch-time constants of the assembly kernels
struct GalDev {
    GalaxyConfig cfg;
    int nkm, nk, nspl;
    double lnkm0, inv_dlnkm;        // model grid (log-uniform)
    double lnki0, inv_dlnki;        // k_interp grid
    double lnk1l0, inv_dlnk1l;      // one-loop grid
};

// walkers per chunk when P(k, mu) is copied to the host under the kernels of the next chunk
static const int GAL_PIPE_CHUNK = 64;     // measured at 256 walkers x 40000 points: 64 -> 7.48 ms per step, 32 -> 7.65, unchunked 8.03

struct GalaxyModel {
    Context* ctx;
    GalaxyConfig cfg;
    int nkm = 0;
    std::vector<double> km;             // model k grid (numpy.logspace(log10 kmin, log10 kmax, Nk))
    DevBuf<double> d_km, d_kstoch;
    ZelPlan zel;                        // Zel'dovich stencils for the model grid
    // interpolate = 1 (rsd/hzpt/interpolated.py:8-72, 208-263): Zel'dovich P00 / P01 / P11 of every cosmology on
    // 100 sigma8_z in [0.3, 1] x 300 k in [INTERP_KMIN, INTERP_KMAX] -- one batched k_zeldovich launch -- read
    // bilinearly in (sigma8_z, k) as the reference's RegularGridInterpolator does; a sigma8 outside the table falls
    // back to the direct evaluation (its InterpolationDomainError branch).  Rebuilt when the spectra change.
    static const int ZT_NS = 100, ZT_NK = 300;
    ZelPlan zel_tab;
    std::vector<double> h_s8grid, h_ktab;
    DevBuf<double> d_s8grid, d_ktab, d_ztab;   // d_ztab [ncosmo][ZT_NS][3][ZT_NK]
    uint64_t ztab_uid = 0, ztab_gen = 0;
    int ztab_nc = 0;
    std::vector<double> ztab_s80;
    void ensure_zel_tables(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par);
    int nk_interp = 0;
    DevBuf<double> d_kinterp;
    // scratch
    DevBuf<double> ptspl, zel1, zel2, plin_m, ratio1, ratio2, base, pairspl, grpspl, grpconst, d_par, d_stoch, ol_m;
    DevBuf<int> d_cmap;
    std::vector<double> h_kinterp;
    // banded resampling operator k_interp -> model grid (see k_pt_resample)
    int rs_band = 0;
    DevBuf<double> d_rsE;
    DevBuf<int> d_rsJ;
    // stochasticity spline evaluated on the model grid, and the second derivatives of the moment splines, as fixed
    // (banded) linear maps of the ordinates (see k_gal_moments)
    DevBuf<double> d_stE, d_momE;
    DevBuf<int> d_momJ;
    int mom_band = 0;
    // optional additive corrections of the P[mu^2], P[mu^4] knots of every (walker, pair): [nw][10][2][nkm], used by the
    // power / poles calls with the same number of walkers until cleared (correct_mu2 / correct_mu4 of the reference)
    DevBuf<double> d_corr;
    int corr_nw = 0;
    void set_mu_corrections(int nw, const double* h_corr);      // h_corr == NULL clears
    // use_mean_bias (power_biased.py:219-237): [nw][10][GAL_NPAIROVR] = sqrt(b1 b1_bar), b2_00_{a,b,c,d}, b2_01_{a,b} and the
    // halo sigma_v (<= 0: sigma_v) at that bias; same life cycle as the corrections
    DevBuf<double> d_ovr;
    int ovr_nw = 0;
    void set_pair_overrides(int nw, const double* h_ovr);       // h_ovr == NULL clears
    // user-loaded dark-matter terms on the model grid (SimLoaderMixin._load, rsd/sim_loader.py:56-66): [GAL_NLOAD][nkm],
    // bit j of the mask = row j (Pdv, P00_mu0, P01_mu2, P11_mu4) replaces the model's term for every walker
    DevBuf<double> d_loaded;
    unsigned loaded_mask = 0;
    void set_loaded(unsigned mask, const double* h_loaded);     // mask == 0 clears
    // asynchronous mode: walker parameters (and evaluation points) go up on the context's upload stream behind ev_par_free
    // (= the previous call's kernels have read them), the main stream waits for ev_par_ready
    cudaEvent_t ev_par_ready = nullptr, ev_par_free = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_event = nullptr;
    cudaEvent_t copy_done[2] = {nullptr, nullptr};      // end of the device-to-host copies of the last two power() calls
    long long copy_ticket = 0;                          // number of power() calls with a host destination so far
    DevBuf<int> d_ident;                                // identity walker -> cosmology map (cmap == NULL)
    int ident_n = 0;
    GalDev last_G;                      // constants of the last power() call (reused by poles())
    int last_nw = 0;

    GalaxyModel(Context& c, const GalaxyConfig& cfg_, const double* kmodel, int nkm_, const double* k_interp, int nk_interp_);
    ~GalaxyModel();
    GalaxyModel(const GalaxyModel&) = delete;
    GalaxyModel& operator=(const GalaxyModel&) = delete;

    // P(k, mu) of `nw` walkers: walker w uses the tables of cosmology cmap[w] in (sp, pt).
    //   par    host [nw][GAL_NPAR]      stoch  host [nw][10][20]
    //   k, mu  host [npts] (paired)     out    DEVICE [nw][npts]
    // h_out (optional, host, ideally pinned): the result is also copied there, chunk by chunk under the kernels of
    // the following walkers
    void power(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par, const double* stoch,
               int npts, const double* d_k, const double* d_mu, double* d_out, double* h_out = nullptr, bool sync = true);
    // sync = false: the call returns once everything is enqueued -- no host-side wait, par / stoch / cmap and h_out must
    // be page-locked and stay valid until wait() (or wait_copy of the returned ticket).  The copy to h_out of call
    // i runs on the copy stream under whatever the context's stream executes next (the PT stage of batch i + 1).
    void wait();                        // everything enqueued so far has finished, results are in host memory
    void wait_copy(long long ticket);   // the host copy of the power() call that returned `ticket` has landed
    long long last_ticket() const { return copy_ticket; }
    // multipoles P_ell(k) of `nw` walkers by Simpson's rule over nmu (odd) mu in [0, 1] -- MultipoleTransfer of
    // rsd/transfers/poles.py:8-69 fused behind the assembly: d_k DEVICE [nk], ells host [nell], d_out DEVICE [nw][nell][nk]
    void poles(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par, const double* stoch,
               int nk, const double* d_k, int nell, const int* ells, int nmu, double* d_out);
    // chi^2 of every walker against a data vector [nell][nk] with precision matrix [n][n], n = nell nk
    // (rsdfit/driver.py:780-802); d_poles [nw][nell][nk] receives the model vectors, d_out [nw] the chi^2
    void chi2(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par, const double* stoch,
              int nk, const double* d_k, int nell, const int* ells, int nmu, const double* d_data, const double* d_cinv,
              double* d_poles, double* d_out);
    // diagnostic: the spline knots of P[mu^0,2,4,6] per (walker, pair): host [nw][10][4][nkm]
    void moments_to_host(int nw, double* out);
    // the dark-matter / bias building blocks on the model grid of the last call: out [nw][n_base()][nkm]
    void base_to_host(int nw, double* out);
    static int n_base();
};

// chi^2 = (M - D)^T C^-1 (M - D) of nw model vectors d_model [nw][n] (FittingDriver.chi2, rsdfit/driver.py:780-802);
// all pointers on the device, no synchronisation
void chi2_rows(Context& c, int nw, int n, const double* d_model, const double* d_data, const double* d_cinv, double* d_out);

} // namespace prsd
