/* pyrsd_b200.h -- C-ABI of the B200-native PT hot path of pyRSD.
 *
 * This is the drop-in boundary: it replaces the SWIG module `pyRSD._gcl`
 * (pyRSD/gcl.i:36-56, pyRSD/_gcl/python/*.i) for the perturbation-theory path.
 * Plain pointers and sizes only; every array is double precision, C-contiguous,
 * HOST memory unless stated otherwise.  Every function returns 0 on success or a
 * negative code (PRSD_ERR_*); prsd_last_error() gives the message.  Nothing here
 * aborts the process (the reference's Common::error does, Common.cpp:67-74) and
 * nothing falls back to the CPU: without a usable sm_100 device
 * prsd_ctx_create fails with PRSD_ERR_NO_DEVICE.
 *
 * Batch convention: a `prsd_spectra` holds `ncosmo` cosmologies; outputs are
 * [ncosmo][...] row-major.
 */
#ifndef PYRSD_B200_H
#define PYRSD_B200_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PRSD_OK              0
#define PRSD_ERR_INVALID    -1
#define PRSD_ERR_CUDA       -2
#define PRSD_ERR_NO_DEVICE  -3
#define PRSD_ERR_INTERNAL   -4

typedef struct prsd_ctx prsd_ctx;
typedef struct prsd_spectra prsd_spectra;
typedef struct prsd_plan prsd_plan;
typedef struct prsd_result prsd_result;
typedef struct prsd_galaxy prsd_galaxy;

const char* prsd_last_error(void);
int prsd_version(void);

/* ---- device context (one per GPU / process rank) */
int prsd_ctx_create(int device, prsd_ctx** out);
int prsd_ctx_destroy(prsd_ctx* ctx);
int prsd_ctx_sync(prsd_ctx* ctx);
int prsd_ctx_launches(prsd_ctx* ctx, long long* n_kernels);
int prsd_ctx_stream(prsd_ctx* ctx, void** cuda_stream);

/* ---- tabulation grid shared by every operator (ln q of the knots) */
int prsd_knot_grid(int* n, double* lnq /* may be NULL */);

/* ---- quadrature configuration: cfg[56] (csrc/quad.h QuadConfig) =
 *      {qmin, ln_far, ln_mid, bao_lo, bao_hi, bao_dq, ln_inner, ln_thin,
 *       6 x {k_below, n_outer, n_inner, n_thin}   Gauss-Legendre orders by k for I_mn, K_mn, one-loop tables,
 *       6 x {k_below, n_outer, n_inner, n_thin}   the same for ImnOneLoop}.
 *      A band applies to k < k_below (first match); at or above the last finite k_below the LAST band applies
 *      (unused bands repeat the last one with k_below = 1e30).
 *      Replaces the reference's `epsrel` constructor argument (Imn.i:8, Jmn.i:8, Kmn.i:8): accuracy is set
 *      by the fixed rule.  Every `cfg56` argument below may be NULL (defaults). */
int prsd_quad_config_default(double* cfg56);

/* ---- linear spectra: replaces Cosmology + LinearPS (LinearPS.i:5-14, LinearPS.cpp:11-15).
 *   k, T   [ncosmo][nspl]  discrete transfer function (Cosmology.GetDiscreteK/Tk)
 *   pars   [ncosmo][8]     {n_s, amplitude = delta_H^2 (sigma8_z/sigma8)^2, h, Omega0_m,
 *                           Omega0_b, T_cmb, 0, 0}                                        */
int prsd_spectra_create(prsd_ctx* ctx, int ncosmo, int nspl, const double* k, const double* T,
                        const double* pars, prsd_spectra** out);
int prsd_spectra_destroy(prsd_spectra* sp);
int prsd_spectra_rescale(prsd_spectra* sp, const double* fac /*[ncosmo]*/);   /* LinearPS::SetSigma8AtZ */
int prsd_spectra_plin(prsd_spectra* sp, int nq, const double* q, double* out /*[ncosmo][nq]*/);
int prsd_spectra_table(prsd_spectra* sp, int kind, double* out /*[ncosmo][n_knots]*/);
/* one-loop spectra: OneLoopPS(P_L, epsrel, oneloop_power) ctor / Evaluate / EvaluateFull
 * (OneLoopPS.i:5-63, OneLoopPS.cpp:21-33); which: 0 Pdd, 1 Pdv, 2 Pvv, 3 P22bar */
int prsd_spectra_set_oneloop(prsd_spectra* sp, const double* tables /*[ncosmo][4][1000]*/);
int prsd_spectra_oneloop_eval(prsd_spectra* sp, int which, int full, int nq, const double* q,
                              double* out /*[ncosmo][nq]*/);

/* OneLoopPdd/Pdv/Pvv/P22Bar(P_L) constructors: compute the four 1000-point one-loop tables of every
 * cosmology of the batch (OneLoopPS.cpp:56-185), install them in `sp`; out (may be NULL) [ncosmo][4][1000] */
int prsd_oneloop_tables(prsd_spectra* sp, const double* cfg56, double* out);
/* OneLoopPS(P_L, epsrel, oneloop_power): install ONE table, host [ncosmo][1000] */
int prsd_spectra_set_oneloop_one(prsd_spectra* sp, int which, const double* table);
int prsd_spectra_get_oneloop(prsd_spectra* sp, double* out /*[ncosmo][4][1000]*/);
/* CubicSpline(x, y)(xq) -- netlib cubic spline of the reference (Spline.cpp:211-273), 0 outside the
 * domain (:362-367); y [batch][n] on shared x [n] */
int prsd_cubic_spline_eval(prsd_ctx* ctx, int batch, int n, const double* x, const double* y, int nq,
                           const double* xq, double* out /*[batch][nq]*/);
/* LinearSpline(x, y)(xq), and EvaluateDerivative when deriv != 0  [Spline.i:24, Spline.cpp:58-130]: y [batch][n] on shared
 * abscissae, 0 outside the domain */
int prsd_linear_spline_eval(prsd_ctx* ctx, int batch, int n, const double* x, const double* y, int nq, const double* xq,
                            int deriv, double* out);

/* ---- one-off evaluations with the call shapes of the pygcl drivers ------------------- */
/* Imn(P_L)(k, m, n)  [Imn.i:8-14, Imn.cpp:142-189] */
int prsd_eval_imn(prsd_spectra* sp, int m, int n, int nk, const double* k, const double* cfg56,
                  double* out /*[ncosmo][nk]*/);
/* Kmn(P_L)(k, m, n, tidal, part)  [Kmn.i:8-14, Kmn.cpp:109-184] */
int prsd_eval_kmn(prsd_spectra* sp, int m, int n, int tidal, int part, int nk, const double* k,
                  const double* cfg56, double* out);
/* Jmn(P_L)(k, m, n)  [Jmn.i:8-14, Jmn.cpp:96-139] */
int prsd_eval_jmn(prsd_spectra* sp, int m, int n, int nk, const double* k, double* out);
/* ImnOneLoop(P1[,P2]).EvaluateLinear/Cross/OneLoop(k, m, n)  [ImnOneLoop.i:8-23,
 * ImnOneLoop.cpp:96-216]; p1, p2: 0 Pdd, 1 Pdv, 2 Pvv (p2 < 0: single-spectrum ctor);
 * which: 0 linear, 1 cross, 2 one-loop */
int prsd_eval_imn_oneloop(prsd_spectra* sp, int p1, int p2, int which, int m, int n, int nk,
                          const double* k, const double* cfg56, double* out);
/* PowerSpectrum::VelocityDispersion(), VelocityDispersion(k, factor), X_Zel, Y_Zel, Q3_Zel,
 * Sigma(R), sigma3_squared  [PowerSpectrum.i:5-45, PowerSpectrum.cpp:28-161]
 * what: 0 sigma_v^2 (n = 1, x ignored), 1 sigma_v^2(k; factor), 2 X_Zel(r), 3 Y_Zel(r), 4 Q3_Zel(k),
 *       5 Sigma(R), 6 sigma3_squared(k) */
int prsd_eval_ps_integral(prsd_spectra* sp, int what, int n, const double* x, double factor,
                          double* out /*[ncosmo][n]*/);
/* CorrelationFunction(P, kmin = 1e-4, kmax = 10)(r) = 1/(2 pi^2 r) \int_kmin^kmax k sin(k r) P(k) dk
 * [CorrelationFunction.i:39-55, CorrelationFunction.cpp:145-168] */
int prsd_eval_correlation(prsd_spectra* sp, int n, const double* r, double kmin, double kmax,
                          double* out /*[ncosmo][n]*/);
/* OneLoopP22Bar::VelocityKurtosis  [OneLoopPS.cpp:187-193] */
int prsd_eval_kurtosis(prsd_spectra* sp, double* out /*[ncosmo]*/);

/* ---- the batched PT stage: everything rsd/pt_integrals.py:42-450 tabulates ------------ */
/* flags: bit 0 = ImnOneLoop integrals, bit 1 = Zel'dovich X(r), Y(r) */
int prsd_plan_create(prsd_ctx* ctx, int nk, const double* k_interp, const double* cfg56, int flags,
                     prsd_plan** out);
int prsd_plan_destroy(prsd_plan* plan);
/* info[8] = {total 2-D nodes, device bytes, build seconds, nk, n_knots, nodes(one-loop op),
 *            nodes(I/K op), nodes(ImnOneLoop op, one half)} */
int prsd_plan_info(prsd_plan* plan, double* info8);
/* wall-clock milliseconds of the build stages: 0 node geometry on the host + upload, 1 W operators (k_build_w),
 * 2 J_mn operator on the one-loop grid, 3 main 1-D operators (J, sigma^2, X, Y), 4 rest, 5-7 unused */
int prsd_plan_build_profile(prsd_plan* plan, double* ms8);
/* *res may be NULL on first use (allocated) and is reused afterwards */
int prsd_plan_run(prsd_plan* plan, prsd_spectra* sp, prsd_result** res);
int prsd_result_destroy(prsd_result* res);
/* what: 0 I [c][16][nk], 1 J [c][6][nk], 2 K [c][13][nk], 3 ML [c][7][3][nk],
 *       4 one-loop tables [c][4][1000], 5 sigmasq_k [c][nk], 6 scalars [c][4] = {sigma_v^2,
 *       velocity kurtosis, Q3/k^4, 0}, 7 X_Zel [c][1024], 8 Y_Zel [c][1024] */
int prsd_result_get(prsd_result* res, int what, double* out, long long n_expected);

/* ---- FFTLog: FortranFFTLog(N, dlnr, mu, q, kr, kropt).Transform(a, dir = 1)
 *      [FortranFFTLog.cpp:10-33, fftlog.f:233-784]; batch of sequences, kr_out = KR() */
int prsd_fftlog(prsd_ctx* ctx, int batch, int N, double mu, double q, double dlnr, double kr, int kropt,
                const double* a /*[batch][N]*/, double* out /*[batch][N]*/, double* kr_out);
/* same with DEVICE pointers (asynchronous on the context's stream) */
/* the same with FFTLog.Transform's direction argument: +1 forward, -1 backward (fftlog.f:523-527) */
int prsd_fftlog_dir(prsd_ctx* ctx, int batch, int N, double mu, double q, double dlnr, double kr, int kropt, int direction,
                    const double* a, double* out, double* kr_out);
int prsd_fftlog_device(prsd_ctx* ctx, int batch, int N, double mu, double q, double dlnr, double kr, int kropt,
                       const double* d_in, double* d_out, double* kr_out);
/* ComputeXiLM(l, m, k, pk, r, smoothing, FFTLOG) x scale  [CorrelationFunction.cpp:56-90];
 * pk_to_xi = (-1)^(l/2) ComputeXiLM(l, 2, ...), xi_to_pk = (-1)^(l/2) 8 pi^3 ComputeXiLM(l, 2, r, xi, k) (:127-141) */
int prsd_compute_xilm(prsd_ctx* ctx, int batch, int l, int m, int nk, const double* k /*[nk]*/,
                      const double* pk /*[batch][nk]*/, int nr, const double* r, double smoothing, double scale,
                      double* out /*[batch][nr]*/);
/* ComputeXiLM(l, m, k, pk, r, smoothing, SIMPS | TRAPZ) x scale  [CorrelationFunction.cpp:92-125]: the discrete
 * quadratures of pk_to_xi / xi_to_pk that rsd/window.py:289-300 uses; method: 1 Simpson, 2 trapezoid;
 * l in {0, 1, 2, 3, 4, 6, 8} */
int prsd_compute_xilm_discrete(prsd_ctx* ctx, int method, int batch, int l, int m, int nk, const double* k /*[nk]*/,
                               const double* pk /*[batch][nk]*/, int nr, const double* r, double smoothing, double scale,
                               double* out /*[batch][nr]*/);
/* SimpsIntegrate(x, y) / TrapzIntegrate(x, y)  [DiscreteQuad.i, DiscreteQuad.cpp:3-66]; method: 1 Simpson, 2 trapezoid;
 * y [batch][n] on shared x [n] (rsd/transfers/poles.py:66 integrates every k row over mu this way) */
int prsd_discrete_integrate(prsd_ctx* ctx, int method, int batch, int n, const double* x, const double* y,
                            double* out /*[batch]*/);
/* ComputeXiLM_fftlog(l, m, k, pk, r, xi, q, smoothing)  [CorrelationFunction.cpp:23-54] */
int prsd_compute_xilm_fftlog(prsd_ctx* ctx, int batch, int l, int m, int N, const double* k, const double* pk,
                             double q, double smoothing, double* r_out /*[N]*/, double* xi_out /*[batch][N]*/);

/* ---- Zel'dovich spectra: ZeldovichP00/P01/P11(ZeldovichPS state)(k)  [ZeldovichPS.i:5-71,
 *      ZeldovichPS.cpp:112-300].  Rows = (cosmology) states: X0, XX, YY [nrow][1024] (XX may be NULL:
 *      X0 + 2 sigma^2), sigmasq [nrow]; ratio [nrow] (may be NULL) = (sigma8_new/sigma8_old)^2 of
 *      SetSigma8AtZ; lowk != 0 enables the low-k SPT forms below k0_low and then needs plin [nrow][nk]
 *      (P_L at the new sigma8) and q3 [nrow] (Q3_Zel(k)/k^4 at the OLD sigma8).
 *      out [nrow][3][nk] = P00, P01, P11. */
int prsd_zeldovich(prsd_ctx* ctx, int nrow, const double* X0, const double* XX, const double* YY,
                   const double* sigmasq, const double* ratio, int lowk, double k0_low, const double* plin,
                   const double* q3, int nk, const double* k, double* out);

/* ---- GalaxySpectrum.power(k, mu)  [rsd/power/gal/power_gal.py:404-425 and everything below it]
 * cfg64 (NULL = defaults, see prsd_galaxy_config_default):
 *   [0] max_mu  [1] fog_model (0 modified_lorentzian, 1 lorentzian, 2 gaussian)
 *   [2..7] use_P00/P01/P11/Pdv/Pvv/Phm_model  [8] use_tidal_bias  [9] use_so_correction  [10] k0_low
 *   [11] interpolate: the Zel'dovich terms of P00 / P01 / P11 / Phm come from 100 sigma8_z x 300 k tables read
 *        bilinearly (InterpolatedHZPTModels, rsd/hzpt/interpolated.py:8-72, 208-263 -- the reference's default);
 *        0 (the library default): evaluated exactly at every call
 *   [16..25] HZPT P00/P01 parameters, [26..34] HZPT P11, [35..54] HZPT Phm
 * kmodel: the model's spline grid self.k (numpy.logspace(log10 kmin, log10 kmax, Nk));
 * k_interp: the PT interpolation grid of the plan.
 * Walker parameters par [nw][56]:
 *   0 sigma8_z, 1 f, 2 alpha_par, 3 alpha_perp, 4 alpha_drag, 5..8 b1_cA b1_cB b1_sA b1_sB, 9 fs, 10 fcB,
 *   11 fsB, 12 sigma_c, 13 sigma_sA, 14 sigma_sB, 15 NcBs, 16 NsBsB, 17 f_so, 18 sigma_so,
 *   19 sigma_v (<= 0: sigma_lin), 20 D(z), 21 sigma8 (cosmo.sigma8()),
 *   22 + 4 kind + bias: nonlinear biases, kind = b2_00_a, b2_00_b, b2_00_c, b2_00_d, b2_01_a, b2_01_b
 *   evaluated at the four linear biases (host inputs: Gaussian-process fits in the reference).
 *   46..49 sigma_v of the four biases (`vel_disp_from_sims`, power_biased.py:268-286; <= 0: sigma_v), 50..55 unused.
 * stoch [nw][10][20]: stochasticity Lambda(k) of the 10 (b1, b1_bar) pairs on 20 log-spaced k between
 *   the ends of kmodel (power_biased.py:342-366), pair order cAcA cAcB cBcB cAsA cAsB cBsA cBsB sAsA sAsB sBsB.
 * cmap [nw] (NULL = identity): cosmology (row of sp / res) each walker uses.
 * k, mu [npts] paired observed coordinates; out [nw][npts]. */
int prsd_galaxy_config_default(double* cfg64);
int prsd_galaxy_create(prsd_ctx* ctx, const double* cfg64, int nkm, const double* kmodel, int nk_interp,
                       const double* k_interp, prsd_galaxy** out);
int prsd_galaxy_destroy(prsd_galaxy* g);
int prsd_galaxy_power(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap,
                      const double* par, const double* stoch, int npts, const double* k, const double* mu,
                      double* out);
/* k, mu, out are DEVICE pointers; returns after the kernels completed */
/* correct_mu2 / correct_mu4 (power_biased.py:462-520): additive corrections of the P[mu^2], P[mu^4] spline knots,
 * corr [nw][10 pairs][2][nkm] (host); used by the following calls with the same nw; corr == NULL clears */
int prsd_galaxy_set_mu_corrections(prsd_galaxy* g, int nw, const double* corr);
/* use_mean_bias (power_biased.py:219-237): both tracers of pair p of walker w carry ovr[w][p][0] = sqrt(b1 b1_bar);
 * ovr[w][p][1..6] = b2_00_a, b2_00_b, b2_00_c, b2_00_d, b2_01_a, b2_01_b at that bias, [7] = halo sigma_v there (<= 0:
 * sigma_v).  ovr [nw][10][8] (host); used by the following calls with the same nw; ovr == NULL clears */
int prsd_galaxy_set_pair_overrides(prsd_galaxy* g, int nw, const double* ovr);
/* SimLoaderMixin._load / _unload (rsd/sim_loader.py:44-75): loaded [4][nkm] (host) on the model grid kmodel, rows
 * Pdv, P00_mu0, P01_mu2, P11_mu4; bit j of mask set = row j replaces the model's term for every walker; mask == 0 clears */
int prsd_galaxy_set_loaded(prsd_galaxy* g, unsigned mask, const double* loaded);
int prsd_galaxy_power_device(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap,
                             const double* par, const double* stoch, int npts, const double* d_k,
                             const double* d_mu, double* d_out);
/* GalaxySpectrum.poles(k, ells, Nmu)  [rsd/power/dm/power_dm.py:1016-1050 -> MultipoleTransfer,
 * rsd/transfers/poles.py:8-69]: P_ell(k) = (2 ell + 1) int_0^1 dmu L_ell(mu) P(k, mu) by Simpson's rule
 * (pygcl.SimpsIntegrate) on nmu equally spaced mu in [0, 1] (nmu odd; the reference turns 40 into 41), fused
 * behind the assembly so that only [nw][nell][nk] numbers leave the device.  k [nk], ells [nell <= 8]. */
int prsd_galaxy_poles(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap,
                      const double* par, const double* stoch, int nk, const double* k, int nell, const int* ells,
                      int nmu, double* out /*[nw][nell][nk]*/);
/* d_k and d_out are DEVICE pointers */
int prsd_galaxy_poles_device(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap,
                             const double* par, const double* stoch, int nk, const double* d_k, int nell,
                             const int* ells, int nmu, double* d_out);
/* FittingDriver.chi2  [rsdfit/driver.py:780-802]: chi^2 = (M - D)^T C^-1 (M - D) of every walker, M = the walker's
 * multipoles in the order [ell][k] (prsd_galaxy_poles), data [nell nk], cinv [nell nk][nell nk] the precision
 * matrix; chi2 [nw]; poles (may be NULL) [nw][nell][nk] receives the model vectors */
int prsd_galaxy_chi2(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap,
                     const double* par, const double* stoch, int nk, const double* k, int nell, const int* ells,
                     int nmu, const double* data, const double* cinv, double* chi2, double* poles);
/* dark-matter and bias building blocks of the last call on the model grid (rsd/power/dm/power_dm.py:797-884,
 * dm/P00.py ... P22.py): out [nw][*nrows][nkm] (out may be NULL to query *nrows), rows in the order
 * 0 P_lin, 1 P00, 2 P01, 3 P11[mu4], 4 Pdv, 5 Pvv, 6 Pdd, 7 P02[mu2] no-velocity, 8 P02[mu4] no-velocity,
 * 9 P12[mu4] no-velocity, 10 P22[mu4] no-velocity, 11-20 K00 K00s K10 K10s K11 K11s K20_a K20_b K20s_a K20s_b,
 * 21 f^2 (h01 + h03 ImnOneLoop), 22 f^2 (h02 + h04), 23 f^3 I03, 24 sigma^2(k)^2, 25 P00 Zel'dovich, 26, 27 mu^6 terms */
int prsd_galaxy_base(prsd_galaxy* g, int nw, int* nrows, double* out);
/* spline knots of P[mu^0], P[mu^2], P[mu^4], P[mu^6] of the last call: out [nw][10][4][nkm] */
int prsd_galaxy_moments(prsd_galaxy* g, int nw, double* out);

/* ---- transfer functions: theory -> what a survey measures --------------------------- */
typedef struct prsd_gridop prsd_gridop;
typedef struct prsd_window prsd_window;
/* GriddedWedgeTransfer.average / GriddedMultipoleTransfer.__call__  [rsd/transfers/grid.py:165-187, 300-319]:
 * out[w][b][k] = sum_mu weights[b][k][mu] * power[w][k][mu].  The caller folds the mode counts N(k, mu), the
 * mu-bin membership (np.digitize on mu_edges) and, for multipoles, (2 ell + 1) L_ell(mu) into `weights`
 * (host, [nbin][nk][nmu], 0 at null grid points): weights = N * leg * [mu in bin] / sum_mu(N * [mu in bin]). */
int prsd_gridop_create(prsd_ctx* ctx, int nbin, int nk, int nmu, const double* weights, prsd_gridop** out);
int prsd_gridop_destroy(prsd_gridop* op);
/* power [nw][nk][nmu] -> out [nw][nbin][nk]; device_ptrs != 0: both are DEVICE pointers and the call does not
 * synchronise */
int prsd_gridop_apply(prsd_gridop* op, int nw, const double* power, double* out, int device_ptrs);

/* ---- the per-step exchange of an ensemble sharded over the GPUs of one box ------------------------------------------
 * The reference gathers one likelihood per walker per ensemble step through emcee's MPI pool
 * [rsdfit/util/mpi_manager.py, rsdfit/engines/mcmc.py].  Here every rank (one process per GPU) owns a receive buffer
 * [2][world][n_per_rank] doubles and a flag word per peer; the ranks map each other's buffers (CUDA IPC: exchange the 128
 * handle bytes by any means, e.g. torch.distributed.all_gather).  The producing kernel of a step writes this rank's
 * n_per_rank results straight into its own slot of its receive buffer; a push kernel stores the slot into the same slot
 * of EVERY peer's buffer over NVLink / NVSwitch peer memory; a one-CTA kernel then raises this rank's flag at every peer
 * and waits (bounded: 2 s, then the status word is set) for theirs.  No staging copy, no library collective; consecutive
 * steps alternate between the two halves of the buffer. */
typedef struct prsd_exchange prsd_exchange;
int prsd_exchange_create(prsd_ctx* ctx, int world, int rank, long long n_per_rank, prsd_exchange** out);
int prsd_exchange_destroy(prsd_exchange* x);
int prsd_exchange_handles(prsd_exchange* x, unsigned char* out128);
/* all_handles [world][128] in rank order (this rank's own entry is ignored); call on every rank, then synchronise the
 * ranks once before the first step */
int prsd_exchange_open(prsd_exchange* x, const unsigned char* all_handles);
/* gridded transfer of nw walkers (prsd_gridop_apply with device pointers) whose [nw][nbin][nk] results are this rank's
 * contribution to the step: enqueued on the context's stream, returns at once */
int prsd_gridop_apply_exchange(prsd_gridop* op, int nw, const double* power, prsd_exchange* x);
/* device pointer to the [world][n_per_rank] results of the latest step (valid behind the step on the context's stream);
 * status (may be NULL; synchronises): 0, or 1 if a wait for a peer timed out */
int prsd_exchange_received(prsd_exchange* x, void** dev_ptr, int* status);
/* WindowFunctionTransfer  [rsd/transfers/window.py:9-132] on the vendored mcfit transforms
 * [extern/mcfit/mcfit.py:86-176, cosmology.py:15-35, kernels.py:14-17].  k [nk]: the log-spaced grid of the
 * multipoles INCLUDING the zero-padded extension (window.py:84-93); ells [nell] even; q the mcfit tilt (1.5);
 * N the FFT length (0 = mcfit's default 2^(ceil(log2 nk) + 1)); must be a power of two in [1024, 8192]. */
int prsd_window_create(prsd_ctx* ctx, int nk, const double* k, int nell, const int* ells, double q, int N,
                       prsd_window** out);
int prsd_window_destroy(prsd_window* w);
/* info [4] = {N, left padding of the input, first output slot, nk}; rr [nk]: the separation grid on which the
 * reference evaluates the window kernels (mcfit's y of the LAST multipole, window.py:98-110) */
int prsd_window_info(prsd_window* w, int* info, double* rr);
/* kern [nell][nell][nk]: kern[i][j][n] multiplies xi_{ells[j]}(rr[n]) into the convolved xi_{ells[i]}
 * (WindowConvolution._get_kernel, rsd/window.py:147-163) */
int prsd_window_set_kernel(prsd_window* w, const double* kern);
/* poles [nw][nell][nk] -> out [nw][nell][nk] (on the same k grid, as the reference assumes); mode 0 convolves,
 * mode 1 is the reference's dry_run (P -> xi -> P); xi / xi_conv (may be NULL, host, [nw][nell][nk]) receive the
 * configuration-space multipoles when device_ptrs == 0 */
int prsd_window_apply(prsd_window* w, int nw, const double* poles, double* out, int mode, int device_ptrs,
                      double* xi, double* xi_conv);
/* cubic spline through (x, y[row]) evaluated at xq: the k_out resampling of window.py:112-124 and
 * RSDSpline; natural end conditions (Spline.cpp:211-273) -- identical to FITPACK's interpolating spline to
 * rounding a few knots away from the ends.  Only the window of knots within 64 of the queried range is solved (a
 * far knot enters the second derivatives through 0.268^distance).  x [n], xq [nq] host; y [nrows][n] -> out
 * [nrows][nq] DEVICE */
int prsd_spline_resample_device(prsd_ctx* ctx, int nrows, int n, const double* x, const double* d_y, int nq,
                                const double* xq, double* d_out);

/* FittingDriver.chi2  [rsdfit/driver.py:780-802] of nw model vectors model [nw][n] (host, or DEVICE when
 * model_on_device != 0) against data [n] with precision matrix cinv [n][n] (host); chi2 [nw] host */
int prsd_chi2(prsd_ctx* ctx, int nw, int n, const double* model, const double* data, const double* cinv,
              double* chi2, int model_on_device);

/* ---- Gaussian-process simulation fits  [rsd/simulation.py:16-256: GeorgeSimulationData on george.GP, ExpSquaredKernel]:
 * mean prediction  y(x*) = y_mean + y_scale sum_i alpha_i amp exp(-1/2 sum_d ((x*_d - x_id) / x_scale_d)^2 / metric_d)
 * x [n][ndim] raw training inputs, x_mean / x_scale [ndim] their StandardScaler, metric [ndim] the kernel metric in scaled
 * units, alpha [n] = (K + diag(yerr^2))^-1 y_scaled (solved once per training set, pyrsd_b200/data/make_simulation_fits.py).
 * prsd_gp_predict: x [npred][ndim] -> out [npred]; device_ptrs != 0: both on the device, no synchronisation. */
typedef struct prsd_gp prsd_gp;
int prsd_gp_create(prsd_ctx* ctx, int n, int ndim, const double* x, const double* x_mean, const double* x_scale,
                   const double* metric, double amp, const double* alpha, double y_mean, double y_scale, prsd_gp** out);
int prsd_gp_destroy(prsd_gp* gp);
int prsd_gp_predict(prsd_gp* gp, int npred, const double* x, double* out, int device_ptrs);

/* ---- BiasToMassRelation / BiasToSigmaRelation  [rsd/tools.py:393-649; the `sigma_sB` constraint of the default fit,
 * rsdfit/theory/gal_defaults.py:174].  mass_out[w] in M_sun/h = root of bias_Tinker(rescale[w] Sigma(R(M)), delta_halo)
 * = b1[w] for M in [mass_lo, mass_hi] (the reference brackets [1e5, 1e16]), R(M) = (3 M / (4 pi rho_bar))^(1/3),
 * rescale = sigma8_z / cosmo.sigma8(), rho_bar = cosmo.rho_bar_z(0) in M_sun/h / (Mpc/h)^3.  One thread per (sigma8_z,
 * b1) pair, Sigma(R) of every cosmology from one operator application; cmap NULL: walker w uses cosmology w (or 0 when
 * the batch holds one).  sigma_out (may be NULL): Sigma(R) on numpy.logspace(-3, 4, 1000), [ncosmo][1000].  Host pointers. */
int prsd_bias_to_mass(prsd_spectra* sp, int nw, const int* cmap, const double* rescale, const double* b1, double rho_bar,
                      double delta_halo, double mass_lo, double mass_hi, double* mass_out, double* sigma_out);

/* ---- measurement support (bench.py) -------------------------------------------------- */
/* per-kernel-family CUDA-event timing on the context's stream; families: 0 I/K operator (DMMA contraction),
 * 1 linear operators, 2 Zel'dovich, 3 galaxy assembly, 4 spline builds, 5 FFTLog, 6 ImnOneLoop operator
 * (k_nodal_apply<SpecML>), 7 one-loop-table operator (k_nodal_apply<SpecDiag>) */
int prsd_ctx_profile(prsd_ctx* ctx, int enable);      /* also resets the counters */
int prsd_ctx_profile_read(prsd_ctx* ctx, double* ms8, long long* count8);
/* FP64 pipe microbenchmarks: mode 0 DFMA, mode 1 DMMA m8n8k4 (TFLOP/s, best of 4) */
int prsd_fp64_peak(prsd_ctx* ctx, int mode, double* tflops);
/* rebuild the linear spectra of a batch from DEVICE-resident k, T [ncosmo][nspl], pars [ncosmo][8] */
int prsd_spectra_reload_device(prsd_spectra* sp, const double* d_k, const double* d_T, const double* d_pars);

/* ---- streaming (asynchronous) path -----------------------------------------------------
 * The reference evaluates one model per process and blocks; a batch pipeline wants step i + 1 enqueued while the
 * device still runs step i and while P(k, mu) of step i crosses PCIe.  These entry points enqueue on the context's
 * stream and return without waiting.  Contract: every host buffer they are given (k, T, pars, par, stoch, cmap, k,
 * mu, out) is PAGE-LOCKED (prsd_host_alloc, cudaHostAlloc, torch pin_memory) and stays untouched until
 * prsd_galaxy_wait / prsd_galaxy_wait_copy / prsd_ctx_sync says the stream has passed it -- a pageable buffer is
 * still correct but makes the call wait for the stream.  One step of the pipeline is
 *     prsd_spectra_reload(sp, k, T, pars);  prsd_plan_run_async(plan, sp, &res);
 *     prsd_galaxy_power_async(g, sp, res, ..., out_i, &ticket_i);
 * and out_i is complete after prsd_galaxy_wait_copy(g, ticket_i); its copy to the host runs on a second stream
 * under the PT kernels of step i + 1, and the input copies of step i + 1 (k, T, pars, par, stoch) run on a third stream
 * under the kernels of step i.  prsd_galaxy_power_async with k == mu == NULL reuses the evaluation points of the previous
 * asynchronous call (same npts).  Same results as the blocking calls, bit for bit. */
int prsd_spectra_reload(prsd_spectra* sp, const double* k, const double* T, const double* pars);   /* shapes of prsd_spectra_create */
int prsd_plan_run_async(prsd_plan* plan, prsd_spectra* sp, prsd_result** res);
int prsd_galaxy_power_async(prsd_galaxy* g, prsd_spectra* sp, prsd_result* res, int nw, const int* cmap,
                            const double* par, const double* stoch, int npts, const double* k, const double* mu,
                            double* out, long long* ticket /* may be NULL */);
int prsd_galaxy_wait(prsd_galaxy* g);                        /* everything enqueued so far is done and on the host */
int prsd_galaxy_wait_copy(prsd_galaxy* g, long long ticket); /* the P(k, mu) of that call is on the host */
int prsd_host_alloc(size_t bytes, void** out);               /* page-locked host memory */
int prsd_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PYRSD_B200_H */
