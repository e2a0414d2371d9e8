// galaxy.cu -- GalaxySpectrum P(k, mu) assembly (see galaxy.cuh).
#include "galaxy.cuh"
#include "pcr.cuh"

#include <algorithm>

#include <cmath>

namespace prsd {

void galaxy_config_default(GalaxyConfig& c)
{
    c.max_mu = 4;
    c.fog_model = 0;
    c.use_P00_model = c.use_P01_model = c.use_P11_model = c.use_Pdv_model = c.use_Pvv_model = c.use_Phm_model = 1;
    c.use_tidal_bias = 0;
    c.use_so_correction = 0;
    c.k0_low = 5e-3;
    c.interpolate = 0;
    // rsd/hzpt/__init__.py:66-92
    const double h00[10] = {707.8, 3.654, 31.76, 0.1257, 3.243, 0.3729, 3.768, -0.1043, 1.699, 0.4224};
    // rsd/hzpt/P11.py:33-61
    const double h11[9] = {658.9, 3.91, 1.917, 18.95, -0.3657, -0.2585, 0.8473, -0.1524, 0.7769};
    // rsd/hzpt/Phm.py:24-68 : A0, R1, R1h, R2h, R  x (amp, alpha, beta, run)
    const double hhm[20] = {751.9, 1.657, 3.752, 0., 5.192, -0.5739, 0.1616, 0., 8.253, -0.8421, -0.1345, 0.,
                            3.053, -1.03, -0.3596, 0., 16.9, -0.1185, -1.067, 0.};
    for (int i = 0; i < 10; i++) c.hz00[i] = h00[i];
    for (int i = 0; i < 9; i++) c.hz11[i] = h11[i];
    for (int i = 0; i < 20; i++) c.hzhm[i] = hhm[i];
}

// ---- row layout of the PT-table spline buffer: [cosmo][PT_NROWS][5 nk]
enum { PT_I = 0, PT_J = 16, PT_K = 22, PT_ML = 35, PT_SIGK = 56, PT_NROWS = 57 };
// ---- base arrays on the model grid: [walker][NB][nkm]
enum {
    B_PLIN = 0, B_P00, B_P01, B_P11MU4, B_PDV, B_PVV, B_PDD, B_P02NV2, B_P02NV4, B_P12NV4, B_P22NV4,
    B_K00, B_K00S, B_K10, B_K10S, B_K11, B_K11S, B_K20A, B_K20B, B_K20SA, B_K20SB,
    B_C11MU2, B_C11MU4, B_I03F3, B_SIGK2, B_Z00, B_MU6A, B_I30F3, NB
};
// ---- walker parameter slots
enum {
    W_S8Z = 0, W_F, W_APAR, W_APERP, W_ADRAG, W_B1 = 5 /* cA cB sA sB */, W_FS = 9, W_FCB, W_FSB, W_SIGC, W_SIGSA,
    W_SIGSB, W_NCBS, W_NSBSB, W_FSO, W_SIGSO, W_SIGV, W_D, W_S80, W_B2 = 22 /* [6 kinds][4 biases] */,
    W_SIGVH = 46 /* sigma_v of the four biases (vel_disp_from_sims), <= 0: sigma_v */, W_END = 50
};

// ------------------------------------------------- PT tables on the model grid
// The reference splines every PT table on k_interp (not-a-knot, rsd/_cache.py:344-389) and evaluates the spline on
// the model grid.  Both grids are fixed for a GalaxyModel and a spline is linear in its ordinates, so
// build-and-evaluate is ONE fixed matrix E [nkm][nk_interp] (column j = the not-a-knot spline through the unit
// vector e_j, evaluated on the model grid; built on the host by the constructor with nak_spline_build /
// pw_cubic_eval_log themselves).  A far knot enters a spline value through 0.268^distance: a band of RS_W = 64
// columns around the interval of k_i reproduces the full product to < 1e-17, and the 14 592 tridiagonal solves
// per batch of 256 cosmologies (185 us of parallel cyclic reduction, 146 MB of coefficients) become a banded
// matrix product.
struct NakSrc { const double* p[5]; int rows[5]; int off[5]; };
static const int RS_W = 64;          // band width
static const int RS_ROWS = 8;        // table rows per CTA

__global__ void __launch_bounds__(256)
k_pt_resample(int nk, int nkm, int band, const double* __restrict__ Eb /*[band][nkm]*/, const int* __restrict__ jlo,
              NakSrc S, double* __restrict__ V /*[nc][PT_NROWS][nkm]*/)
{
    extern __shared__ double sm_rs[];            // [RS_ROWS][nk]
    const int r0 = blockIdx.x*RS_ROWS, cc = blockIdx.y, tid = threadIdx.x;
    const int nr = min(RS_ROWS, PT_NROWS - r0);
    for (int r = 0; r < nr; r++) {
        const int row = r0 + r;
        int g = 0;
#pragma unroll
        for (int j = 1; j < 5; j++) if (row >= S.off[j]) g = j;
        const double* src = S.p[g] + ((size_t)cc*S.rows[g] + (row - S.off[g]))*nk;
        for (int i = tid; i < nk; i += blockDim.x) sm_rs[r*nk + i] = src[i];
    }
    __syncthreads();
    for (int i = tid; i < nkm; i += blockDim.x) {
        const int j0 = jlo[i];
        double acc[RS_ROWS];
#pragma unroll
        for (int r = 0; r < RS_ROWS; r++) acc[r] = 0.0;
        for (int jj = 0; jj < band; jj++) {
            const double e = __ldg(Eb + (size_t)jj*nkm + i);
#pragma unroll
            for (int r = 0; r < RS_ROWS; r++) if (r < nr) acc[r] = fma(e, sm_rs[r*nk + j0 + jj], acc[r]);
        }
        for (int r = 0; r < nr; r++) V[((size_t)cc*PT_NROWS + r0 + r)*nkm + i] = acc[r];
    }
}

// not-a-knot system of nak_spline_build (pt_math.cuh) for the interior second derivatives M_1 .. M_{n-2}, written
// into the PCR arrays at offset `off` (unknown j = i - 1); x, y in shared memory
__device__ __forceinline__ void nak_system_row(int n, const double* x, const double* y, int i,
                                               double* a, double* b, double* c, double* r, int off)
{
    const int m = n - 2;
    const double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
    double bb = 2.0*(h0 + h1), cc = h1, aa = h0;
    if (i == 1) { bb += h0*(1.0 + h0/h1); cc -= h0*h0/h1; }
    if (i == m) { bb += h1*(1.0 + h1/h0); aa -= h1*h1/h0; }
    a[off + i - 1] = aa; b[off + i - 1] = bb; c[off + i - 1] = cc;
    r[off + i - 1] = 6.0*((y[i + 1] - y[i])/h1 - (y[i] - y[i - 1])/h0);
}

// --------------------------------------------------------------- HZPT broadband
__device__ __forceinline__ double pade_bb(double k, double A0, double R, double R1, double R1h, double R2h)
{
    const double k2 = k*k;
    const double F = 1.0 - 1.0/(1.0 + k2*R*R);
    const double r2 = k*R2h, r22 = r2*r2;
    return F*A0*(1.0 + k2*R1*R1)/(1.0 + k2*R1h*R1h + r22*r22);
}


__global__ void k_gal_base(GalDev G, const double* __restrict__ km, const int* __restrict__ cmap,
                           const double* __restrict__ par, const double* __restrict__ ptspl,
                           const double* __restrict__ plin_m, const double* __restrict__ olspl,
                           const double* __restrict__ scal, const double* __restrict__ zel1, const double* __restrict__ zel2,
                           const double* __restrict__ loaded /*[GAL_NLOAD][nkm] or NULL*/, unsigned loaded_mask,
                           double* __restrict__ base)
{
    const int w = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
    const int nkm = G.nkm;
    if (i >= nkm) return;
    const int c = cmap[w];
    const double* p = par + (size_t)w*GAL_NPAR;
    const double k = km[i], k2 = k*k;
    const double s8z = p[W_S8Z], f = p[W_F], D = p[W_D];
    const double norm = (s8z/p[W_S80])*(s8z/p[W_S80]);
    const double norm2 = norm*norm;
    // PT table `row` of cosmology c at k_i: resampled from k_interp by k_pt_resample
    auto T = [&](int row) { return ptspl[((size_t)c*PT_NROWS + row)*nkm + i]; };
    auto OL = [&](int which) {
        const double* s = olspl + (size_t)(c*4 + which)*5*ONELOOP_N;
        return (k < ONELOOP_KMIN || k > ONELOOP_KMAX) ? 0.0 : cubic_spline_eval_log(ONELOOP_N, s, G.lnk1l0, G.inv_dlnk1l, k);
    };
    auto I = [&](int m, int n) { return norm2*T(PT_I + 4*m + n); };
    // J order: J00 J01 J02 J10 J11 J20
    auto J = [&](int j) { return norm*T(PT_J + j); };
    auto K = [&](int j) { return norm2*T(PT_K + j); };
    auto ML = [&](int m) { return norm2*T(PT_ML + 3*m) + norm2*norm*T(PT_ML + 3*m + 1) + norm2*norm2*T(PT_ML + 3*m + 2); };

    const double PL0 = plin_m[(size_t)c*nkm + i];
    const double Plin = norm*PL0;
    const double* z1 = zel1 + (size_t)w*3*nkm;
    const double* z2 = zel2 + (size_t)w*3*nkm;
    const double* h0 = G.cfg.hz00;
    auto bb00 = [&](double s8) {
        const double x = s8/0.8;
        return pade_bb(k, h0[0]*pow(x, h0[1]), h0[2]*pow(x, h0[3]), h0[4]*pow(x, h0[5]), h0[6]*pow(x, h0[7]), h0[8]*pow(x, h0[9]));
    };
    // ---- P00 (dm/P00.py:9-26)
    double P00;
    if (G.cfg.use_P00_model) P00 = bb00(s8z) + z1[i];
    else P00 = Plin + 2.0*I(0, 0) + 6.0*k2*J(0)*Plin;
    // ---- P01 (dm/P01.py:9-31; broadband hzpt/P01.py:57-90)
    double P01;
    if (G.cfg.use_P01_model) {
        const double x = s8z/0.8;
        const double A0 = h0[0]*pow(x, h0[1]), R = h0[2]*pow(x, h0[3]), R1 = h0[4]*pow(x, h0[5]);
        const double R1h = h0[6]*pow(x, h0[7]), R2h = h0[8]*pow(x, h0[9]);
        const double dA0 = f*h0[1]*A0, dR = f*h0[3]*R, dR1 = f*h0[5]*R1, dR1h = f*h0[7]*R1h, dR2h = f*h0[9]*R2h;
        const double k4 = k2*k2;
        const double N = 1.0 + k2*R*R, F = 1.0 - 1.0/N;
        const double num = 1.0 + k2*R1*R1, den = 1.0 + k2*R1h*R1h + k4*R2h*R2h*R2h*R2h;
        const double bb = F*num/den*dA0 + 2.0*A0*k2*R*num/(N*N)/den*dR + 2.0*A0*k4*R*R*R1/N/den*dR1
                        - 2.0*A0*k4*R*R*num*R1h/N/(den*den)*dR1h - 4.0*A0*k4*k2*R*R*num*R2h*R2h*R2h/N/(den*den)*dR2h;
        P01 = bb + 2.0*f*z1[nkm + i];
    } else P01 = 2.0*f*(Plin + 4.0*(I(0, 0) + 3.0*k2*J(0)*Plin));
    // ---- P11[mu4] total (dm/P11.py:67-82)
    double P11mu4;
    if (G.cfg.use_P11_model) {
        const double* h = G.cfg.hz11;
        const double x = s8z/0.8, y = f/0.5;
        const double A0 = h[0]*pow(x, h[1])*pow(y, h[2]), R = h[3]*pow(x, h[4])*pow(y, h[5]), R1h = h[6]*pow(x, h[7])*pow(y, h[8]);
        const double F = 1.0 - 1.0/(1.0 + k2*R*R);
        P11mu4 = F*A0/(1.0 + k2*R1h*R1h) + f*f*z1[2*nkm + i] - f*f*I(3, 1);
    } else {
        const double part2 = 2.0*I(1, 1) + 4.0*I(2, 2) + 6.0*k2*(J(4) + 2.0*J(3))*Plin;
        P11mu4 = f*f*(Plin + part2 + I(1, 3));
    }
    // ---- Pdd, Pdv, Pvv (power_dm.py:676-718, 848-884)
    const double Pdd = norm*(PL0 + norm*OL(0));
    double Pdv, Pvv;
    {
        double P00_z0 = 0.0, zs2 = 1.0, P00_hz = 0.0;
        if (G.cfg.use_Pdv_model || G.cfg.use_Pvv_model) {
            P00_z0 = bb00(s8z/D) + z2[i];
            P00_hz = bb00(s8z) + z1[i];          // self.hzpt.P00(k) regardless of use_P00_model
            const double zs = 3.0/(D + D*D + D*D*D);
            zs2 = zs*zs;
        }
        if (G.cfg.use_Pdv_model) {
            const double g = (-12483.8*sqrt(P00_z0) + 2.554*P00_z0*P00_z0)/(1381.29 + 2.540*P00_z0);
            Pdv = -f*((g - P00_z0)/zs2 + P00_hz);
        } else Pdv = -f*norm*(PL0 + norm*OL(1));
        if (G.cfg.use_Pvv_model) {
            const double g = (-12480.5*sqrt(P00_z0) + 1.824*P00_z0*P00_z0)/(2165.87 + 1.796*P00_z0);
            Pvv = f*f*((g - P00_z0)/zs2 + P00_hz);
        } else Pvv = f*f*norm*(PL0 + norm*OL(2));
    }
    // user-loaded terms replace the model's (SimLoaderMixin, rsd/sim_loader.py:44-75; dm/P00.py, P01.py, P11.py, power_dm.py:862-863)
    if (loaded_mask) {
        if (loaded_mask & 1u) Pdv = loaded[0*nkm + i];
        if (loaded_mask & 2u) P00 = loaded[1*nkm + i];
        if (loaded_mask & 4u) P01 = loaded[2*nkm + i];
        if (loaded_mask & 8u) P11mu4 = loaded[3*nkm + i];
    }
    double* o = base + (size_t)w*NB*nkm + i;
    const double f2 = f*f, f3 = f2*f, f4 = f2*f2;
    o[B_PLIN*nkm] = Plin;
    o[B_P00*nkm] = P00;
    o[B_P01*nkm] = P01;
    o[B_P11MU4*nkm] = P11mu4;
    o[B_PDV*nkm] = Pdv;
    o[B_PVV*nkm] = Pvv;
    o[B_PDD*nkm] = Pdd;
    o[B_P02NV2*nkm] = f2*(I(0, 2) + 2.0*k2*J(2)*Plin);          // dm/P02.py:8-21
    o[B_P02NV4*nkm] = f2*(I(2, 0) + 2.0*k2*J(5)*Plin);          // dm/P02.py:49-56
    o[B_P12NV4*nkm] = f3*(I(1, 2) - I(0, 3) + 2.0*k2*J(2)*Plin);// dm/P12.py:8-20
    o[B_P22NV4*nkm] = (1.0/16.0)*f4*I(2, 3);                    // dm/P22.py:8-21
    for (int j = 0; j < 10; j++) {
        // K00 K00s K10 K10s K11 K11s K20_a K20_b K20s_a K20s_b from the K row order
        const int rowmap[10] = {0, 2, 5, 7, 6, 8, 9, 10, 11, 12};
        o[(B_K00 + j)*nkm] = K(rowmap[j]);
    }
    o[B_C11MU2*nkm] = f2*(ML(0) + ML(2));                        // biased/P11.py:8-17
    o[B_C11MU4*nkm] = f2*(ML(1) + ML(3));                        // biased/P11.py:36-39
    o[B_I03F3*nkm] = f3*I(0, 3);
    { const double sk = norm*T(PT_SIGK); o[B_SIGK2*nkm] = sk*sk; }
    o[B_Z00*nkm] = z1[i];
    o[B_MU6A*nkm] = f3*(I(2, 1) + 2.0*k2*J(5)*Plin) + 0.125*f4*I(3, 2);   // biased/P12.py:33-45 + power_biased.py:525
    o[B_I30F3*nkm] = f3*I(3, 0);
    (void)scal;
}

// pair table: bias indices (0 cA, 1 cB, 2 sA, 3 sB)
__constant__ int c_pair_a[GAL_NPAIR] = {0, 0, 1, 0, 0, 1, 1, 2, 2, 3};
__constant__ int c_pair_b[GAL_NPAIR] = {0, 1, 1, 2, 3, 2, 3, 2, 3, 3};

// ----------------------------------------------------------- halo moments + splines
// (One CTA per (walker, pair), 128 threads.  Five pairs per CTA -- one load of a spline-operator entry feeding 20 splines
// held in shared memory instead of 4 -- measured 0.175 -> 0.28 ms: the moment formulae, not the banded products, are the
// cost, and 512 CTAs of 182 registers run them five pairs in sequence.  Rejected, round 2.)
__global__ void __launch_bounds__(128)
k_gal_moments(GalDev G, const double* __restrict__ km, const double* __restrict__ Es /*[GAL_NSTOCH][nkm]*/,
              int mband, const double* __restrict__ Mb /*[mband][nkm]*/, const int* __restrict__ mlo,
              const int* __restrict__ cmap,
              const double* __restrict__ par, const double* __restrict__ stoch, const double* __restrict__ scal,
              const double* __restrict__ base, const double* __restrict__ corr /*[nw][10][2][nkm] or NULL*/,
              const double* __restrict__ pairovr /*[nw][10][GAL_NPAIROVR] or NULL*/,
              double* __restrict__ pairspl)
{
    extern __shared__ double sm_gm[];
    const int nkm = G.nkm;
    double* X = sm_gm;                       // [nkm]
    double* Y = X + nkm;                     // [4][nkm]
    double* W = Y + 4*nkm;                   // [4][nkm] second derivatives
    double* S = W + 4*nkm;                   // stochasticity ordinates [20]
    const int pr = blockIdx.x, w = blockIdx.y, tid = threadIdx.x;
    const int c = cmap[w];
    const double* p = par + (size_t)w*GAL_NPAR;
    const double* B = base + (size_t)w*NB*nkm;
    const int ia = c_pair_a[pr], ib = c_pair_b[pr];
    // use_mean_bias (power_biased.py:219-237): both tracers of a pair carry the geometric-mean bias, and every quantity the
    // reference evaluates at _ib1 / _ib1_bar (nonlinear biases, halo velocity dispersion) is handed in per pair
    const double* ov = pairovr ? pairovr + ((size_t)w*GAL_NPAIR + pr)*GAL_NPAIROVR : nullptr;
    const double b1 = ov ? ov[0] : p[W_B1 + ia], b1b = ov ? ov[0] : p[W_B1 + ib];
    // kind: 00_a 00_b 00_c 00_d 01_a 01_b; which: bias index of the tracer
    auto b2 = [&](int kind, int which) { return ov ? ov[1 + kind] : p[W_B2 + kind*4 + which]; };
    const double f = p[W_F], s8z = p[W_S8Z];
    const double norm = (s8z/p[W_S80])*(s8z/p[W_S80]);
    const double sig_lin = sqrt(norm*scal[(size_t)c*4 + 0]);
    const double sigv = p[W_SIGV] > 0.0 ? p[W_SIGV] : sig_lin;   // sigma_v defaults to sigma_lin (power_dm.py:454-460)
    // sigmav_halo / sigmav_halo_bar (power_biased.py:268-286): sigma_v, or -- vel_disp_from_sims -- the simulation fit at the
    // bias of each tracer, handed in per bias
    const double sva0 = ov ? ov[7] : p[W_SIGVH + ia], svb0 = ov ? ov[7] : p[W_SIGVH + ib];
    const double sva = sva0 > 0.0 ? sva0 : sigv, svb = svb0 > 0.0 ? svb0 : sigv;
    const double sig2 = sva*sva, sig2b = svb*svb;
    const double vkurt = norm*norm*scal[(size_t)c*4 + 1];
    const double bs = G.cfg.use_tidal_bias ? -2.0/7.0*(b1 - 1.0) : 0.0;
    const double bsb = G.cfg.use_tidal_bias ? -2.0/7.0*(b1b - 1.0) : 0.0;

    for (int i = tid; i < nkm; i += blockDim.x) X[i] = km[i];
    if (tid < GAL_NSTOCH) S[tid] = stoch[((size_t)w*GAL_NPAIR + pr)*GAL_NSTOCH + tid];
    __syncthreads();
    const double* hm = G.cfg.hzhm;
    // HaloMatterBase._powerlaw (hzpt/Phm.py:231-236): the five Pade parameters depend on (sigma8_z, b1) only --
    // evaluated once per (walker, bias), not per k
    double hp1[5] = {0, 0, 0, 0, 0}, hp2[5] = {0, 0, 0, 0, 0};
    if (G.cfg.use_Phm_model) {
        const double x = s8z/0.8;
        for (int side = 0; side < 2; side++) {
            const double bb = side ? b1b : b1, lb = log(bb);
            double* hp = side ? hp2 : hp1;
            for (int j = 0; j < 5; j++) hp[j] = hm[4*j]*pow(bb, hm[4*j + 1] + hm[4*j + 3]*lb)*pow(x, hm[4*j + 2]);
        }
    }
    auto phm = [&](double bb, const double* hp, int which, double k, double P00, double Z00, double K00, double K00s, double bsx) {
        if (G.cfg.use_Phm_model) return pade_bb(k, hp[0], hp[4], hp[1], hp[2], hp[3]) + bb*Z00;
        return bb*P00 + b2(0, which)*K00 + bsx*K00s;      // power_biased.py:325-340
    };
    for (int i = tid; i < nkm; i += blockDim.x) {
        const double k = X[i], fk2 = f*f*k*k, fk4 = fk2*fk2;
        const double P00 = B[B_P00*nkm + i], P01 = B[B_P01*nkm + i], Pdv = B[B_PDV*nkm + i], Pvv = B[B_PVV*nkm + i];
        const double K00 = B[B_K00*nkm + i], K00s = B[B_K00S*nkm + i], K10 = B[B_K10*nkm + i], K10s = B[B_K10S*nkm + i];
        const double K11 = B[B_K11*nkm + i], K11s = B[B_K11S*nkm + i];
        const double P02nv2 = B[B_P02NV2*nkm + i], P02nv4 = B[B_P02NV4*nkm + i];
        // P00_ss (biased/P00.py:19-43) with the stochasticity spline (power_biased.py:342-366)
        // the 20-point not-a-knot spline evaluated at k_i is a fixed linear map of its ordinates (Es, built on the host)
        double st = 0.0;
#pragma unroll 4
        for (int j = 0; j < GAL_NSTOCH; j++) st = fma(__ldg(Es + (size_t)j*nkm + i), S[j], st);
        const double Phm1 = phm(b1, hp1, ia, k, P00, B[B_Z00*nkm + i], K00, K00s, bs);
        const double Phm2 = (ia == ib) ? Phm1 : phm(b1b, hp2, ib, k, P00, B[B_Z00*nkm + i], K00, K00s, bsb);
        const double mu0 = Phm1*Phm2/P00 + st;
        // P01_mu2 with a given b2_01 kind (biased/P01.py:3-24)
        auto P01mu2 = [&](int kind) {
            const double c1 = b2(kind, ia), c2 = b2(kind, ib);
            return b1*b1b*P01 - Pdv*(b1*(1.0 - b1b) + b1b*(1.0 - b1)) + f*((c1 + c2)*K10 + (bs + bsb)*K10s)
                 + f*((b1b*c1 + b1*c2)*K11 + (b1b*bs + b1*bsb)*K11s);
        };
        // Phh_mu0 with a given b2_00 kind (biased/P00.py:3-17)
        auto Phh = [&](int kind) {
            const double c1 = b2(kind, ia), c2 = b2(kind, ib);
            return b1*b1b*P00 + (b1*c2 + b1b*c1)*K00 + (bs*c2 + bsb*c1)*K00s;
        };
        const double P11ss_mu2 = b1*b1b*B[B_C11MU2*nkm + i];
        // P[mu^2] (power_biased.py:494-505; biased/P02.py:5-33)
        const double P02ss_mu2 = 0.5*(b1 + b1b)*P02nv2 - 0.5*fk2*(sig2 + sig2b)*Phh(2)
                               + 0.5*f*f*((b2(2, ia) + b2(2, ib))*B[B_K20A*nkm + i] + (bs + bsb)*B[B_K20SA*nkm + i]);
        const double mu2 = P01mu2(4) + P11ss_mu2 + P02ss_mu2;
        // P[mu^4] (power_biased.py:507-520)
        const double P11ss_mu4 = 0.5*(b1 + b1b)*B[B_P11MU4*nkm + i] - 0.5*((b1 - 1.0) + (b1b - 1.0))*Pvv
                               + B[B_C11MU4*nkm + i]*(b1*b1b - 0.5*(b1 + b1b));                    // biased/P11.py:19-44
        const double P02ss_mu4 = 0.5*(b1 + b1b)*P02nv4
                               + 0.5*f*f*((b2(1, ia) + b2(1, ib))*B[B_K20B*nkm + i] + (bs + bsb)*B[B_K20SB*nkm + i]);  // P02.py:35-54
        const double P01b = P01mu2(5);
        const double P12ss_mu4 = B[B_P12NV4*nkm + i] - 0.5*fk2*0.5*(sig2 + sig2b)*P01
                               - 0.5*((b1 - 1.0) + (b1b - 1.0))*B[B_I03F3*nkm + i]
                               - 0.25*fk2*(sig2 + sig2b)*(P01b - P01);                                    // biased/P12.py:4-22
        const double P03ss_mu4 = -0.25*fk2*(sig2 + sig2b)*P01b;                                            // biased/P03.py:4-16
        const double PhhD = Phh(3);
        const double P22ss_mu4 = B[B_P22NV4*nkm + i] + 0.5*fk4*(b1*b1b*B[B_PDD*nkm + i])*B[B_SIGK2*nkm + i]
                               - 0.25*fk2*(sig2 + sig2b)*(0.5*(b1 + b1b)*P02nv2)
                               + 0.125*fk4*(sig2*sig2 + sig2b*sig2b)*PhhD;                                  // biased/P22.py:4-34
        const double P13ss_mu4 = -0.5*fk2*(sig2 + sig2b)*P11ss_mu2;                                        // biased/P13.py:3-19
        const double P04ss_mu4 = -0.125*(b1 + b1b)*fk2*(sig2 + sig2b)*P02nv2
                               + (1.0/12.0)*fk4*PhhD*(1.5*(sig2*sig2 + sig2b*sig2b) + vkurt);               // biased/P04.py:4-27
        const double mu4 = P11ss_mu4 + P02ss_mu4 + P12ss_mu4 + P03ss_mu4 + P22ss_mu4 + P13ss_mu4 + P04ss_mu4;
        // P[mu^6] (power_biased.py:522-528; biased/P12.py:24-45)
        const double mu6 = B[B_MU6A*nkm + i] - 0.5*(b1 + b1b)*B[B_I30F3*nkm + i];
        // correct_mu2 / correct_mu4 (power_biased.py:462-520): simulation-calibrated residuals added to the knots
        double cm2 = 0.0, cm4 = 0.0;
        if (corr) {
            const double* cr = corr + ((size_t)w*GAL_NPAIR + pr)*2*nkm;
            cm2 = cr[i];
            cm4 = cr[nkm + i];
        }
        Y[i] = mu0; Y[nkm + i] = mu2 + cm2; Y[2*nkm + i] = mu4 + cm4; Y[3*nkm + i] = mu6;
    }
    __syncthreads();
    // not-a-knot splines of the four moments on the model grid (the reference's RSDSpline on self.k).  The second
    // derivatives are a fixed linear map of the ordinates (the inverse of the tridiagonal not-a-knot system, built on
    // the host); a far knot enters through 0.268^distance, so a band of `mband` columns reproduces it to < 1e-17 --
    // no tridiagonal solve, no barrier-separated reduction sweeps
    double* Mm = W;                          // [4][nkm] second derivatives
    for (int idx = tid; idx < 4*nkm; idx += blockDim.x) {
        const int m = idx/nkm, i = idx - m*nkm;
        const double* y = Y + m*nkm + mlo[i];
        // four independent partial sums with their loads in flight together (the single chain waited on L2 latency)
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int jj = 0;
        for (; jj + 3 < mband; jj += 4) {
            const double e0 = __ldg(Mb + (size_t)jj*nkm + i), e1 = __ldg(Mb + (size_t)(jj + 1)*nkm + i);
            const double e2 = __ldg(Mb + (size_t)(jj + 2)*nkm + i), e3 = __ldg(Mb + (size_t)(jj + 3)*nkm + i);
            a0 = fma(e0, y[jj], a0);
            a1 = fma(e1, y[jj + 1], a1);
            a2 = fma(e2, y[jj + 2], a2);
            a3 = fma(e3, y[jj + 3], a3);
        }
        for (; jj < mband; jj++) a0 = fma(__ldg(Mb + (size_t)jj*nkm + i), y[jj], a0);
        Mm[idx] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    // store [walker][pair][moment][Y | b | c | d]
    double* o = pairspl + ((size_t)w*GAL_NPAIR + pr)*16*nkm;
    for (int idx = tid; idx < 4*nkm; idx += blockDim.x) {
        const int m = idx/nkm, i = idx - m*nkm;
        const int j = min(i, nkm - 2);       // the last knot repeats the last interval's coefficients
        const double* y = Y + m*nkm;
        const double* M = Mm + m*nkm;
        const double h = X[j + 1] - X[j];
        o[(m*4 + 0)*nkm + i] = y[i];
        o[(m*4 + 1)*nkm + i] = (y[j + 1] - y[j])/h - h*(2.0*M[j] + M[j + 1])/6.0;
        o[(m*4 + 2)*nkm + i] = 0.5*M[j];
        o[(m*4 + 3)*nkm + i] = (M[j + 1] - M[j])/(6.0*h);
    }
}

// ------------------------------------------------------------------ P(k, mu)
__device__ __forceinline__ double fog(int model, double x)
{
    const double h = 1.0 + 0.5*x*x;
    if (model == 0) return 1.0/(h*h);            // fog_kernels.py:40-52
    if (model == 1) return 1.0/h;                // :57-69
    return exp(-0.5*x*x);                        // :72-84
}

// The ten sub-spectra enter P_gal only through six distinct FOG products (gal/Pcc, Pcs, Pss coefficient tables,
// SURVEY Appendix A):  G_c^2 {cAcA, cAcB, cBcB},  G_c G_sA {cAsA, cBsA},  G_c G_sB {cAsB, cBsB},  G_sA^2 {sAsA},
// G_sA G_sB {sAsB},  G_sB^2 {sBsB}.  Splines are linear in their ordinates, so the weighted sums of the moment
// splines inside a group are formed once per walker; a (k, mu) point then reads 6 x 3 cubics instead of 10 x 3.
// Group g keeps [moment][Y | b | c | d][nkm] like pairspl, plus its 1-halo constant in grpconst[w][6].
static const int GAL_NGRP = 6;
static const int GAL_NCONST = 12;    // per walker: 6 one-halo constants + {1/alpha_perp, 1/F, 1/F^2 - 1, alpha_drag^3/(alpha_perp^2 alpha_par)}

__global__ void k_gal_groups(int nkm, const double* __restrict__ par, const double* __restrict__ pairspl,
                             double* __restrict__ grpspl, double* __restrict__ grpconst)
{
    const int w = blockIdx.y;
    const double* p = par + (size_t)w*GAL_NPAR;
    const double fcB = p[W_FCB], fsB = p[W_FSB];
    // weight of pair pr inside its group
    const double wt[GAL_NPAIR] = {(1.0 - fcB)*(1.0 - fcB), 2.0*fcB*(1.0 - fcB), fcB*fcB,
                                  (1.0 - fcB)*(1.0 - fsB), (1.0 - fcB)*fsB, fcB*(1.0 - fsB), fcB*fsB,
                                  (1.0 - fsB)*(1.0 - fsB), 2.0*fsB*(1.0 - fsB), fsB*fsB};
    const int grp[GAL_NPAIR] = {0, 0, 0, 1, 2, 1, 2, 3, 4, 5};
    // pairspl [w][pair][moment][Y | b | c | d][nkm]  ->  grpspl [w][knot][group][moment][Y b c d]: the 96 numbers a
    // (k, mu) point needs are contiguous, addressed from one base pointer with constant offsets
    const int n = 16*nkm;
    for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < n; idx += gridDim.x*blockDim.x) {
        const int mc = idx/nkm, i = idx - mc*nkm;          // mc = 4 moment + coefficient
        double acc[GAL_NGRP] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int pr = 0; pr < GAL_NPAIR; pr++) acc[grp[pr]] += wt[pr]*pairspl[((size_t)w*GAL_NPAIR + pr)*n + idx];
#pragma unroll
        for (int g = 0; g < GAL_NGRP; g++) grpspl[(((size_t)w*nkm + i)*GAL_NGRP + g)*16 + mc] = acc[g];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double* c = grpconst + (size_t)w*GAL_NCONST;
        const double apar = p[W_APAR], aperp = p[W_APERP], adrag = p[W_ADRAG], F = apar/aperp;
        c[6] = 1.0/aperp;
        c[7] = 1.0/F;
        c[8] = 1.0/(F*F) - 1.0;
        c[9] = adrag*adrag*adrag/(aperp*aperp*apar);          // rsd/tools.py:211-214
        c[10] = c[11] = 0.0;
        c[0] = 0.0;
        c[1] = wt[5]*p[W_NCBS];          // cBsA
        c[2] = wt[6]*p[W_NCBS];          // cBsB
        c[3] = 0.0;
        c[4] = 0.0;
        c[5] = wt[9]*p[W_NSBSB];         // sBsB
    }
}

// P_gal(k_obs, mu_obs) of walker parameters p from its group splines
// AP remapping and interval search: (k_obs, mu_obs) -> (k', mu', interval i of the model grid); false outside the grid
__device__ __forceinline__ bool gal_locate(const GalDev& G, const double* __restrict__ km, const double* __restrict__ gc,
                                           double kobs, double muobs, double& k, double& mu, int& i)
{
    const int nkm = G.nkm;
    // rsd/tools.py:54-75: F = alpha_par/alpha_perp, k' = (k/alpha_perp) sqrt(1 + mu^2 (F^-2 - 1)), mu' = (mu/F)/sqrt(...)
    k = kobs*gc[6]; mu = muobs;
    if (gc[8] != 0.0) {
        const double y = 1.0 + muobs*muobs*gc[8];
        const double inv = rsqrt(y);
        k = kobs*gc[6]*(y*inv);
        mu = muobs*gc[7]*inv;
    }
    i = 0;
    if (!(k >= km[0] && k <= km[nkm - 1])) return false;     // the reference raises InterpolationDomainError (_cache.py:327-339)
    i = (int)((log(k) - G.lnkm0)*G.inv_dlnkm);
    if (i < 0) i = 0;
    if (i > nkm - 2) i = nkm - 2;
    while (i > 0 && k < km[i]) --i;
    while (i < nkm - 2 && k > km[i + 1]) ++i;
    return true;
}

// P(k', mu') from the coefficient block `knot` [6][4][4] of interval i (global memory, or staged in shared memory)
template <bool STAGED>
__device__ __forceinline__ double gal_eval(const GalDev& G, const double* __restrict__ p, const double* knot,
                                           const double* __restrict__ gc, double k, double mu, double t)
{
    const double mu2 = mu*mu;
    const int nmom = G.cfg.max_mu/2 + 1;
    const double fs = p[W_FS];
    const int fm = G.cfg.fog_model;
    const double kmu = k*mu;
    const bool so = G.cfg.use_so_correction != 0;
    double Gc, GsA, GsB;
    if (fm == 0) {
        // modified Lorentzian 1/(1 + x^2/2)^2 for the three dispersions with ONE division
        const double xc = kmu*p[W_SIGC], xa = kmu*p[W_SIGSA], xb = kmu*p[W_SIGSB];
        const double hc = 1.0 + 0.5*xc*xc, ha = 1.0 + 0.5*xa*xa, hb = 1.0 + 0.5*xb*xb;
        const double inv = 1.0/(hc*ha*hb);
        const double ic = inv*ha*hb, ia = inv*hc*hb, ib = inv*hc*ha;
        Gc = ic*ic; GsA = ia*ia; GsB = ib*ib;
    } else {
        Gc = fog(fm, kmu*p[W_SIGC]); GsA = fog(fm, kmu*p[W_SIGSA]); GsB = fog(fm, kmu*p[W_SIGSB]);
    }
    auto grp = [&](int g) {
        const double2* q = reinterpret_cast<const double2*>(knot + g*16);
        double tot = gc[g], mpow = 1.0;
#pragma unroll 4
        for (int m = 0; m < nmom; m++) {
            const double2 yb = STAGED ? q[2*m] : __ldg(q + 2*m), cd = STAGED ? q[2*m + 1] : __ldg(q + 2*m + 1);
            tot += mpow*(yb.x + t*(yb.y + t*(cd.x + t*cd.y)));
            mpow *= mu2;
        }
        return tot;
    };
    double Pcc;
    if (so) {
        // SOCorrection sets sigma_c = 0 inside Pcc and re-damps the group (gal/Pcc/__init__.py:39-60)
        const double fso = p[W_FSO], fcB = p[W_FCB];
        const double G1 = Gc, G2 = fog(fm, kmu*p[W_SIGSO]);
        Pcc = ((G1*(1.0 - fso))*(G1*(1.0 - fso)) + 2.0*fso*(1.0 - fso)*G1*G2 + (G2*fso)*(G2*fso))*grp(0)
            + 2.0*G1*G2*fso*fcB*p[W_NCBS];
    } else {
        Pcc = Gc*Gc*grp(0);
    }
    const double Pcs = Gc*GsA*grp(1) + Gc*GsB*grp(2);
    const double Pss = GsA*GsA*grp(3) + GsA*GsB*grp(4) + GsB*GsB*grp(5);
    const double res = (1.0 - fs)*(1.0 - fs)*Pcc + 2.0*fs*(1.0 - fs)*Pcs + fs*fs*Pss;
    return res*gc[9];
}


__device__ __forceinline__ double gal_pkmu_point(const GalDev& G, const double* __restrict__ km, const double* __restrict__ p,
                                                 const double* __restrict__ gs /*[nkm][6][4][4]*/, const double* __restrict__ gc,
                                                 double kobs, double muobs)
{
    double k, mu;
    int i;
    if (!gal_locate(G, km, gc, kobs, muobs, k, mu, i)) return nan("");
    return gal_eval<false>(G, p, gs + (size_t)i*(GAL_NGRP*16), gc, k, mu, k - km[i]);
}

// One thread per (k, mu) point.  The 128 points of a CTA are neighbours in the caller's list, so their intervals of
// the model grid usually span a few knots: those coefficient blocks (768 B each) are staged in shared memory once,
// and the 36 16-byte reads of every point hit shared memory instead of waiting on L2 (ncu before: long-scoreboard
// stalls 16 per issue, L1 hit rate 64 %).  CTAs whose points span more than PK_SPAN knots read global memory.
static const int PK_SPAN = 8;

__global__ void __launch_bounds__(128)
k_gal_pkmu(GalDev G, const double* __restrict__ km, const double* __restrict__ par, const double* __restrict__ grpspl,
           const double* __restrict__ grpconst, int npts, const double* __restrict__ kk, const double* __restrict__ mm,
           double* __restrict__ out)
{
    __shared__ __align__(16) double s_knots[PK_SPAN*GAL_NGRP*16];
    __shared__ int s_lo, s_hi;
    const int w = blockIdx.y, j = blockIdx.x*blockDim.x + threadIdx.x;
    const double* p = par + (size_t)w*GAL_NPAR;
    const double* gs = grpspl + (size_t)w*GAL_NGRP*16*G.nkm;
    const double* gc = grpconst + (size_t)w*GAL_NCONST;
    if (threadIdx.x == 0) { s_lo = G.nkm; s_hi = -1; }
    __syncthreads();
    double k = 0.0, mu = 0.0;
    int i = 0;
    const bool live = j < npts;
    const bool ok = live && gal_locate(G, km, gc, kk[j], mm[j], k, mu, i);
    if (ok) { atomicMin(&s_lo, i); atomicMax(&s_hi, i); }
    __syncthreads();
    const int lo = s_lo, hi = s_hi;
    const bool staged = hi >= lo && hi - lo < PK_SPAN;
    if (staged) {
        const int n2 = (hi - lo + 1)*(GAL_NGRP*16/2);
        const double2* src = reinterpret_cast<const double2*>(gs + (size_t)lo*(GAL_NGRP*16));
        double2* dst = reinterpret_cast<double2*>(s_knots);
        for (int t = threadIdx.x; t < n2; t += blockDim.x) dst[t] = __ldg(src + t);
    }
    __syncthreads();
    if (!live) return;
    double v;
    if (!ok) v = nan("");
    else if (staged) v = gal_eval<true>(G, p, s_knots + (size_t)(i - lo)*(GAL_NGRP*16), gc, k, mu, k - km[i]);
    else v = gal_eval<false>(G, p, gs + (size_t)i*(GAL_NGRP*16), gc, k, mu, k - km[i]);
    // streaming store: the 82 MB of results per batch would otherwise evict the walkers' coefficient tables from L2
    __stcs(out + (size_t)w*npts + j, v);
}

// P_ell(k) = (2 ell + 1) \int_0^1 dmu L_ell(mu) P(k, mu) with Simpson's rule on nmu (odd) equally spaced mu:
// MultipoleTransfer (rsd/transfers/poles.py:8-69; SimpsIntegrate of DiscreteQuad.cpp:41-66 on a uniform grid).
// One thread per (walker, k); out [walker][ell index][k].
__device__ __forceinline__ double legendre_even(int ell, double x)
{
    const double x2 = x*x;
    switch (ell) {
        case 0: return 1.0;
        case 2: return 0.5*(3.0*x2 - 1.0);
        case 4: return 0.125*((35.0*x2 - 30.0)*x2 + 3.0);
        case 6: return 0.0625*(((231.0*x2 - 315.0)*x2 + 105.0)*x2 - 5.0);
        default: {      // generic recurrence (odd ell or ell > 6)
            double p0 = 1.0, p1 = x;
            if (ell == 1) return x;
            for (int n = 2; n <= ell; n++) { const double pn = ((2.0*n - 1.0)*x*p1 - (n - 1.0)*p0)/n; p0 = p1; p1 = pn; }
            return p1;
        }
    }
}

__global__ void __launch_bounds__(128)
k_gal_poles(GalDev G, const double* __restrict__ km, const double* __restrict__ par, const double* __restrict__ grpspl,
            const double* __restrict__ grpconst, int nk, const double* __restrict__ kk, int nell, const int* __restrict__ ells,
            int nmu, double* __restrict__ out)
{
    const int w = blockIdx.y, j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= nk) return;
    const double* p = par + (size_t)w*GAL_NPAR;
    const double* gs = grpspl + (size_t)w*GAL_NGRP*16*G.nkm;
    const double* gc = grpconst + (size_t)w*GAL_NCONST;
    const double k = kk[j];
    const double h = 1.0/(nmu - 1);
    double acc[8];
    for (int e = 0; e < nell; e++) acc[e] = 0.0;
    for (int m = 0; m < nmu; m++) {
        const double mu = (m == nmu - 1) ? 1.0 : m*h;
        const double sw = (m == 0 || m == nmu - 1) ? 1.0 : ((m & 1) ? 4.0 : 2.0);       // Simpson weights x 3/h
        const double P = gal_pkmu_point(G, km, p, gs, gc, k, mu);
        for (int e = 0; e < nell; e++) acc[e] += sw*legendre_even(ells[e], mu)*P;
    }
    for (int e = 0; e < nell; e++) out[((size_t)w*nell + e)*nk + j] = (2.0*ells[e] + 1.0)*acc[e]*h*(1.0/3.0);
}

__global__ void k_plin_model(int nspl, const PlinParams* __restrict__ params, const double* __restrict__ spl,
                             int nq, const double* __restrict__ q, double* __restrict__ out)
{
    const int c = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= nq) return;
    out[(size_t)c*nq + j] = plin_eval(params[c], spl + (size_t)c*5*nspl, q[j]);
}

// RegularGridInterpolator._find_indices (rsd/_interpolate.py:185-203): i = searchsorted(grid, x) - 1 clipped to
// [0, n - 2], distance (x - grid[i]) / (grid[i + 1] - grid[i])
__device__ __forceinline__ int grid_locate(const double* __restrict__ g, int n, double x, double& y)
{
    int lo = 0, hi = n;                    // first index with g[idx] >= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] < x) lo = mid + 1; else hi = mid;
    }
    int i = lo - 1;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    y = (x - g[i])/(g[i + 1] - g[i]);
    return i;
}

// Zel'dovich terms of every walker from the sigma8 x k tables: which = 0 -> P00, P01, P11 at sigma8_z; which = 1 ->
// P00 at sigma8_z / D.  A walker whose sigma8 lies outside the table keeps what `out` already holds (the direct value).
__global__ void k_zel_table_read(int nkm, const double* __restrict__ km, int ns, const double* __restrict__ s8g, int nkt,
                                 const double* __restrict__ kt, const double* __restrict__ ztab, const int* __restrict__ cmap,
                                 const double* __restrict__ par, int which, double* __restrict__ out /*[n][3][nkm]*/)
{
    const int w = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= nkm) return;
    const double* p = par + (size_t)w*GAL_NPAR;
    const double s8 = which == 0 ? p[W_S8Z] : p[W_S8Z]/p[W_D];
    if (!(s8 >= s8g[0] && s8 <= s8g[ns - 1])) return;
    const double k = km[i];
    if (!(k >= kt[0] && k <= kt[nkt - 1])) return;
    double ys, yk;
    const int is = grid_locate(s8g, ns, s8, ys), ik = grid_locate(kt, nkt, k, yk);
    const double* T = ztab + (size_t)cmap[w]*ns*3*nkt;
    const int nspec = which == 0 ? 3 : 1;
    for (int t = 0; t < nspec; t++) {
        const double* a = T + ((size_t)is*3 + t)*nkt + ik;
        const double* b = a + (size_t)3*nkt;
        out[((size_t)w*3 + t)*nkm + i] = (1.0 - ys)*(1.0 - yk)*a[0] + (1.0 - ys)*yk*a[1] + ys*(1.0 - yk)*b[0] + ys*yk*b[1];
    }
}

// normalisation ratios of every walker: (sigma8_z / sigma8)^2 and the same at sigma8_z / D (Jennings fits)
__global__ void k_gal_ratios(int nw, const double* __restrict__ par, double* __restrict__ r1, double* __restrict__ r2)
{
    const int w = blockIdx.x*blockDim.x + threadIdx.x;
    if (w >= nw) return;
    const double* p = par + (size_t)w*GAL_NPAR;
    const double n1 = p[W_S8Z]/p[W_S80], n2 = n1/p[W_D];
    r1[w] = n1*n1;
    r2[w] = n2*n2;
}

__global__ void k_scale_rows(int n, const double* __restrict__ in, const double* __restrict__ fac, const int* __restrict__ cmap,
                             double* __restrict__ out)
{
    // out[w][i] = fac[w] * in[cmap[w]][i]
    const int w = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[(size_t)w*n + i] = fac[w]*in[(size_t)cmap[w]*n + i];
}

// chi^2 = (M - D)^T C^-1 (M - D) per walker (rsdfit/driver.py:780-802): M = the walker's multipoles in the order
// [ell][k] (the reference's `combined_model`), D the data vector, C^-1 the precision matrix [n][n].
__global__ void __launch_bounds__(256)
k_gal_chi2(int n, const double* __restrict__ model /*[nw][n]*/, const double* __restrict__ data, const double* __restrict__ cinv,
           double* __restrict__ out)
{
    extern __shared__ double sm_chi[];
    double* diff = sm_chi;                 // [n]
    __shared__ double red[8];
    const int w = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) diff[i] = model[(size_t)w*n + i] - data[i];
    __syncthreads();
    double acc = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double* row = cinv + (size_t)i*n;
        double y = 0.0;
        for (int j = 0; j < n; j++) y = fma(row[j], diff[j], y);
        acc = fma(diff[i], y, acc);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += red[i];
        out[w] = t;
    }
}

// ------------------------------------------------------------------ host
GalaxyModel::GalaxyModel(Context& c, const GalaxyConfig& cfg_, const double* kmodel, int nkm_, const double* k_interp, int nki)
    : ctx(&c), cfg(cfg_), nkm(nkm_), nk_interp(nki)
{
    PRSD_REQUIRE(nkm >= 4 && nkm <= 512, "GalaxyModel: model grid size %d outside [4, 512]", nkm);
    PRSD_REQUIRE(cfg.max_mu == 0 || cfg.max_mu == 2 || cfg.max_mu == 4 || cfg.max_mu == 6, "`max_mu` must be one of [0, 2, 4, 6]");
    PRSD_REQUIRE(cfg.fog_model >= 0 && cfg.fog_model <= 2, "unknown FOG model %d", cfg.fog_model);
    km.assign(kmodel, kmodel + nkm);
    for (int i = 0; i < nkm; i++) PRSD_REQUIRE(km[i] > 0 && (i == 0 || km[i] > km[i - 1]), "model k grid must be positive and increasing");
    d_km.upload(km, c.stream);
    // GP_NK log-spaced points between the ends of the model grid (power_biased.py:357)
    std::vector<double> ks(GAL_NSTOCH);
    const double l0 = std::log10(km[0]), l1 = std::log10(km[nkm - 1]);
    for (int i = 0; i < GAL_NSTOCH; i++) ks[i] = std::pow(10.0, l0 + (l1 - l0)*i/(GAL_NSTOCH - 1));
    d_kstoch.upload(ks, c.stream);
    d_kinterp.upload(k_interp, nki, c.stream);
    h_kinterp.assign(k_interp, k_interp + nki);
    PRSD_REQUIRE(nki >= 4, "GalaxyModel: the PT interpolation grid needs at least 4 points");
    zel.build(c, km.data(), nkm);
    if (cfg.interpolate) {
        // InterpolationTable.grid (interpolated.py:27-29): numpy.linspace(0.3, 1.0, 100), numpy.logspace(log10(k_interp[0]),
        // log10(k_interp[-1]), 300) -- the PT interpolation range INTERP_KMIN .. INTERP_KMAX
        h_s8grid.resize(ZT_NS);
        for (int i = 0; i < ZT_NS; i++) h_s8grid[i] = 0.3 + (1.0 - 0.3)/(ZT_NS - 1)*i;
        h_s8grid[ZT_NS - 1] = 1.0;
        h_ktab.resize(ZT_NK);
        const double a0 = std::log10(k_interp[0]), a1 = std::log10(k_interp[nki - 1]);
        for (int i = 0; i < ZT_NK; i++) h_ktab[i] = std::pow(10.0, a0 + (a1 - a0)/(ZT_NK - 1)*i);
        h_ktab[ZT_NK - 1] = std::pow(10.0, a1);
        d_s8grid.upload(h_s8grid, c.stream);
        d_ktab.upload(h_ktab, c.stream);
        zel_tab.build(c, h_ktab.data(), ZT_NK);
    }
    {
        // E[i][j] = (not-a-knot spline through e_j on k_interp)(km[i]); band of rs_band columns around the interval of km[i]
        const int n = nki;
        const double lnx0 = std::log(k_interp[0]), inv = (n - 1)/std::log(k_interp[n - 1]/k_interp[0]);
        std::vector<double> E((size_t)nkm*n), y(n), b(n), cc(n), d(n), M(n), w(n);
        for (int j = 0; j < n; j++) {
            std::fill(y.begin(), y.end(), 0.0);
            y[j] = 1.0;
            nak_spline_build(n, k_interp, y.data(), b.data(), cc.data(), d.data(), M.data(), w.data());
            for (int i = 0; i < nkm; i++)
                E[(size_t)i*n + j] = pw_cubic_eval_log(n, k_interp, y.data(), b.data(), cc.data(), d.data(), lnx0, inv, km[i]);
        }
        rs_band = std::min(RS_W, n);
        std::vector<double> Eb((size_t)rs_band*nkm);
        std::vector<int> jl(nkm);
        for (int i = 0; i < nkm; i++) {
            int ji = (int)(std::upper_bound(k_interp, k_interp + n, km[i]) - k_interp) - 1;
            int lo = std::min(std::max(ji - rs_band/2 + 1, 0), n - rs_band);
            jl[i] = lo;
            for (int jj = 0; jj < rs_band; jj++) Eb[(size_t)jj*nkm + i] = E[(size_t)i*n + lo + jj];
        }
        d_rsE.upload(Eb, c.stream);
        d_rsJ.upload(jl, c.stream);
    }
    {
        // stochasticity: (not-a-knot spline through e_j on the 20 GP points)(km[i])  (power_biased.py:342-366)
        const int n = GAL_NSTOCH;
        const double lnx0 = std::log(ks[0]), inv = (n - 1)/std::log(ks[n - 1]/ks[0]);
        std::vector<double> Es((size_t)n*nkm), y(n), b(n), cc(n), d(n), M(n), w(n);
        for (int j = 0; j < n; j++) {
            std::fill(y.begin(), y.end(), 0.0);
            y[j] = 1.0;
            nak_spline_build(n, ks.data(), y.data(), b.data(), cc.data(), d.data(), M.data(), w.data());
            for (int i = 0; i < nkm; i++)
                Es[(size_t)j*nkm + i] = pw_cubic_eval_log(n, ks.data(), y.data(), b.data(), cc.data(), d.data(), lnx0, inv, km[i]);
        }
        d_stE.upload(Es, c.stream);
    }
    {
        // moments: second derivatives of the not-a-knot spline on the model grid as a banded map of the ordinates
        const int n = nkm;
        std::vector<double> Mop((size_t)n*n), y(n), b(n), cc(n), d(n), M(n), w(n);
        for (int j = 0; j < n; j++) {
            std::fill(y.begin(), y.end(), 0.0);
            y[j] = 1.0;
            nak_spline_build(n, km.data(), y.data(), b.data(), cc.data(), d.data(), M.data(), w.data());
            for (int i = 0; i < n; i++) Mop[(size_t)i*n + j] = M[i];
        }
        mom_band = std::min(RS_W, n);
        std::vector<double> Mb((size_t)mom_band*n);
        std::vector<int> lo(n);
        for (int i = 0; i < n; i++) {
            lo[i] = std::min(std::max(i - mom_band/2, 0), n - mom_band);
            for (int jj = 0; jj < mom_band; jj++) Mb[(size_t)jj*n + i] = Mop[(size_t)i*n + lo[i] + jj];
        }
        d_momE.upload(Mb, c.stream);
        d_momJ.upload(lo, c.stream);
    }
    c.sync();
}

GalaxyModel::~GalaxyModel()
{
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    for (cudaEvent_t e : copy_done) if (e) cudaEventDestroy(e);
    if (copy_event) cudaEventDestroy(copy_event);
    if (ev_par_ready) { if (ctx->upload_stream) cudaStreamSynchronize(ctx->upload_stream); cudaEventDestroy(ev_par_ready); cudaEventDestroy(ev_par_free); }
    if (copy_stream) cudaStreamDestroy(copy_stream);
}

void GalaxyModel::ensure_zel_tables(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par)
{
    Context& c = *ctx;
    cudaStream_t s = c.stream;
    const int nc = sp.ncosmo;
    // sigma8 of every cosmology at the normalisation of its P_L (the walkers carry it)
    std::vector<double> s80(nc, 0.0);
    for (int w = 0; w < nw; w++) s80[cmap ? cmap[w] : w] = par[(size_t)w*GAL_NPAR + W_S80];
    bool same = ztab_uid == sp.uid && ztab_gen == sp.lin_gen && ztab_nc == nc && ztab_s80.size() == s80.size();
    for (int i = 0; same && i < nc; i++) same = s80[i] == 0.0 || s80[i] == ztab_s80[i];
    if (same) return;
    for (int i = 0; i < nc; i++)
        if (s80[i] == 0.0) s80[i] = (ztab_s80.size() == (size_t)nc && ztab_s80[i] > 0) ? ztab_s80[i] : 1.0;    // cosmology without a walker: never read
    const int rows = nc*ZT_NS;
    std::vector<int> hmap(rows);
    std::vector<double> hratio(rows);
    for (int ci = 0; ci < nc; ci++)
        for (int i = 0; i < ZT_NS; i++) {
            hmap[ci*ZT_NS + i] = ci;
            const double r = h_s8grid[i]/s80[ci];
            hratio[ci*ZT_NS + i] = r*r;
        }
    DevBuf<int> dmap;
    DevBuf<double> dratio, plin_t((size_t)nc*ZT_NK), plin_rows((size_t)rows*ZT_NK);
    dmap.upload(hmap, s);
    dratio.upload(hratio, s);
    k_plin_model<<<dim3((ZT_NK + 127)/128, nc), 128, 0, s>>>(sp.nspl, sp.params.p, sp.spl.p, ZT_NK, d_ktab.p, plin_t.p);
    k_scale_rows<<<dim3((ZT_NK + 127)/128, rows), 128, 0, s>>>(ZT_NK, plin_t.p, dratio.p, dmap.p, plin_rows.p);
    c.launches += 2;
    d_ztab.ensure((size_t)rows*3*ZT_NK);
    // the drivers of the tables have the low-k approximation on (HaloZeldovichBase.from_zeldovich, hzpt/__init__.py:60-63)
    zeldovich_eval_map(c, zel_tab, rows, dmap.p, pt.X0.p, nullptr, pt.YY.p, pt.scalars.p, 4, dratio.p, true, cfg.k0_low,
                       plin_rows.p, pt.scalars.p + 2, 4, d_ztab.p);
    PRSD_CUDA(cudaStreamSynchronize(s));           // the host vectors go out of scope
    ztab_uid = sp.uid; ztab_gen = sp.lin_gen; ztab_nc = nc; ztab_s80 = s80;
}

void GalaxyModel::power(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par, const double* stoch,
                        int npts, const double* d_k, const double* d_mu, double* d_out, double* h_out, bool sync)
{
    Context& c = *ctx;
    cudaStream_t s = c.stream;
    PRSD_REQUIRE(nw > 0 && npts >= 0, "GalaxyModel::power: empty request");
    PRSD_REQUIRE(pt.nk == nk_interp, "GalaxyModel::power: PT tables are on %d k, model expects %d", pt.nk, nk_interp);
    PRSD_REQUIRE(sp.have_oneloop, "GalaxyModel::power: the PT stage has not been run for this batch");
    PRSD_REQUIRE(sp.ctx == ctx, "GalaxyModel::power: the model and the spectra batch belong to different contexts");
    PRSD_REQUIRE(pt.ncosmo == sp.ncosmo && pt.src_uid == sp.uid && pt.src_gen == sp.lin_gen,
                 "GalaxyModel::power: the PT result was not computed from this spectra batch in its current state "
                 "(result: %d cosmologies of batch %llu.%llu, spectra: %d of batch %llu.%llu); run the plan on these spectra first",
                 pt.ncosmo, (unsigned long long)pt.src_uid, (unsigned long long)pt.src_gen, sp.ncosmo,
                 (unsigned long long)sp.uid, (unsigned long long)sp.lin_gen);
    const int nc = sp.ncosmo;
    for (int w = 0; w < nw; w++) {
        const int cm = cmap ? cmap[w] : w;
        PRSD_REQUIRE(cm >= 0 && cm < nc, "walker %d maps to cosmology %d (batch has %d)", w, cm, nc);
        const double* p = par + (size_t)w*GAL_NPAR;
        PRSD_REQUIRE(p[W_S8Z] > 0 && p[W_S80] > 0 && p[W_D] > 0, "walker %d: sigma8_z, sigma8 and D must be positive", w);
        PRSD_REQUIRE(p[W_APAR] > 0 && p[W_APERP] > 0, "walker %d: AP parameters must be positive", w);
    }
    // No pageable staging on the per-call path (a copy from pageable memory waits for the stream first): the identity
    // map is uploaded once, a caller's map / parameters go up straight from the caller's (ideally pinned) memory, and
    // the two normalisation ratios are formed on the device.
    if (cmap) d_cmap.upload(cmap, nw, s);
    else if (ident_n < nw) {
        std::vector<int> id(std::max(nw, 1024));
        for (size_t i = 0; i < id.size(); i++) id[i] = (int)i;
        d_ident.upload(id, s);
        PRSD_CUDA(cudaStreamSynchronize(s));      // the host vector goes out of scope (once per size)
        ident_n = (int)id.size();
    }
    const int* d_map = cmap ? d_cmap.p : d_ident.p;
    if (!sync) {
        // one batch ahead of the kernels on the upload stream (see SpectraBatch::reload_host)
        cudaStream_t up = c.uploads();
        if (!ev_par_ready) {
            PRSD_CUDA(cudaEventCreateWithFlags(&ev_par_ready, cudaEventDisableTiming));
            PRSD_CUDA(cudaEventCreateWithFlags(&ev_par_free, cudaEventDisableTiming));
            PRSD_CUDA(cudaEventRecord(ev_par_free, s));
        }
        if (d_par.n < (size_t)nw*GAL_NPAR || d_stoch.n < (size_t)nw*GAL_NPAIR*GAL_NSTOCH) {
            PRSD_CUDA(cudaStreamSynchronize(s));      // growing: the old blocks may still be read (once per size)
            d_par.ensure((size_t)nw*GAL_NPAR);
            d_stoch.ensure((size_t)nw*GAL_NPAIR*GAL_NSTOCH);
        }
        PRSD_CUDA(cudaStreamWaitEvent(up, ev_par_free, 0));
        d_par.upload(par, (size_t)nw*GAL_NPAR, up);
        d_stoch.upload(stoch, (size_t)nw*GAL_NPAIR*GAL_NSTOCH, up);
        PRSD_CUDA(cudaEventRecord(ev_par_ready, up));
        PRSD_CUDA(cudaStreamWaitEvent(s, ev_par_ready, 0));
    } else {
        d_par.upload(par, (size_t)nw*GAL_NPAR, s);
        d_stoch.upload(stoch, (size_t)nw*GAL_NPAIR*GAL_NSTOCH, s);
    }
    ratio1.ensure(nw);
    ratio2.ensure(nw);
    k_gal_ratios<<<(nw + 127)/128, 128, 0, s>>>(nw, d_par.p, ratio1.p, ratio2.p);
    c.launches++;
    // a previous call's copy to the host may still be reading d_out: its end is ordered before this call's writes
    // (in stream order this wait sits AFTER the PT stage of the current batch, which therefore overlaps that copy)
    if (copy_ticket > 0 && copy_done[(copy_ticket - 1) & 1]) PRSD_CUDA(cudaStreamWaitEvent(s, copy_done[(copy_ticket - 1) & 1], 0));

    bool zel_direct = true;
    if (cfg.interpolate) {
        ensure_zel_tables(sp, pt, nw, cmap, par);
        // the direct evaluation is only needed for walkers whose sigma8_z or sigma8_z / D leaves the table
        zel_direct = false;
        for (int w = 0; w < nw && !zel_direct; w++) {
            const double* p = par + (size_t)w*GAL_NPAR;
            const double a = p[W_S8Z], b = p[W_S8Z]/p[W_D];
            zel_direct = !(a >= 0.3 && a <= 1.0 && b >= 0.3 && b <= 1.0) || !(km[0] >= h_ktab[0] && km[nkm - 1] <= h_ktab[ZT_NK - 1]);
        }
    }
    // ---- per-cosmology: every PT table splined on k_interp and read on the model grid (rsd/_cache.py:344-389)
    const int nk = nk_interp;
    ptspl.ensure((size_t)nc*PT_NROWS*nkm);
    {
        ProfScope prof(c, PROF_SPLINE);
        const size_t smem = (size_t)RS_ROWS*nk*sizeof(double);
        if (smem > 48*1024) PRSD_CUDA(cudaFuncSetAttribute(k_pt_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NakSrc S = {{pt.I.p, pt.J.p, pt.K.p, pt.ML.p, pt.sigmasq_k.p}, {16, 6, 13, 21, 1}, {PT_I, PT_J, PT_K, PT_ML, PT_SIGK}};
        k_pt_resample<<<dim3((PT_NROWS + RS_ROWS - 1)/RS_ROWS, nc), 256, smem, s>>>(nk, nkm, rs_band, d_rsE.p, d_rsJ.p, S, ptspl.p);
        c.launches++;
        PRSD_CUDA(cudaGetLastError());
    }
    // ---- P_L on the model grid (z = 0 normalisation of the batch)
    plin_m.ensure((size_t)nc*nkm);
    k_plin_model<<<dim3((nkm + 127)/128, nc), 128, 0, s>>>(sp.nspl, sp.params.p, sp.spl.p, nkm, d_km.p, plin_m.p);
    c.launches++;
    GalDev G;
    G.cfg = cfg;
    G.nkm = nkm; G.nk = nk; G.nspl = sp.nspl;
    G.lnkm0 = std::log(km[0]); G.inv_dlnkm = (nkm - 1)/std::log(km[nkm - 1]/km[0]);
    G.lnki0 = std::log(h_kinterp[0]); G.inv_dlnki = (nk - 1)/std::log(h_kinterp[nk - 1]/h_kinterp[0]);
    G.lnk1l0 = std::log(ONELOOP_KMIN); G.inv_dlnk1l = (ONELOOP_N - 1)/std::log(ONELOOP_KMAX/ONELOOP_KMIN);
    last_G = G;
    last_nw = nw;

    DevBuf<double> plin_w((size_t)nw*nkm), plin_w2((size_t)nw*nkm);
    zel1.ensure((size_t)nw*3*nkm);
    zel2.ensure((size_t)nw*3*nkm);
    base.ensure((size_t)nw*NB*nkm);
    pairspl.ensure((size_t)nw*GAL_NPAIR*16*nkm);
    grpspl.ensure((size_t)nw*GAL_NGRP*16*nkm);
    grpconst.ensure((size_t)nw*GAL_NCONST);
    const size_t smem_mom = (size_t)(9*nkm + GAL_NSTOCH)*sizeof(double);
    if (smem_mom > 48*1024) PRSD_CUDA(cudaFuncSetAttribute(k_gal_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mom));

    // ---- everything below is per walker: walkers [w0, w0 + n) of the batch
    auto run_walkers = [&](int w0, int n) {
        const int* cm = d_map + w0;
        const double* pr = d_par.p + (size_t)w0*GAL_NPAR;
        const double* st = d_stoch.p + (size_t)w0*GAL_NPAIR*GAL_NSTOCH;
        double* pw1 = plin_w.p + (size_t)w0*nkm;
        double* pw2 = plin_w2.p + (size_t)w0*nkm;
        double* z1 = zel1.p + (size_t)w0*3*nkm;
        double* z2 = zel2.p + (size_t)w0*3*nkm;
        double* bs = base.p + (size_t)w0*NB*nkm;
        double* ps = pairspl.p + (size_t)w0*GAL_NPAIR*16*nkm;
        double* gs = grpspl.p + (size_t)w0*GAL_NGRP*16*nkm;
        double* gc = grpconst.p + (size_t)w0*GAL_NCONST;
        // Zel'dovich terms on the model grid at sigma8_z and at sigma8_z / D (Jennings fits)
        k_scale_rows<<<dim3((nkm + 127)/128, n), 128, 0, s>>>(nkm, plin_m.p, ratio1.p + w0, cm, pw1);
        k_scale_rows<<<dim3((nkm + 127)/128, n), 128, 0, s>>>(nkm, plin_m.p, ratio2.p + w0, cm, pw2);
        c.launches += 2;
        if (zel_direct) {
            zeldovich_eval_map(c, zel, n, cm, pt.X0.p, nullptr, pt.YY.p, pt.scalars.p, 4, ratio1.p + w0, true, cfg.k0_low,
                               pw1, pt.scalars.p + 2, 4, z1);
            zeldovich_eval_map(c, zel, n, cm, pt.X0.p, nullptr, pt.YY.p, pt.scalars.p, 4, ratio2.p + w0, true, cfg.k0_low,
                               pw2, pt.scalars.p + 2, 4, z2, true);             // only P00 is read at sigma8_z / D
        }
        if (cfg.interpolate) {
            for (int which = 0; which < 2; which++)
                k_zel_table_read<<<dim3((nkm + 127)/128, n), 128, 0, s>>>(nkm, d_km.p, ZT_NS, d_s8grid.p, ZT_NK, d_ktab.p, d_ztab.p,
                                                                         cm, pr, which, which == 0 ? z1 : z2);
            c.launches += 2;
        }
        ProfScope prof_gal(c, PROF_GALAXY);
        k_gal_base<<<dim3((nkm + 127)/128, n), 128, 0, s>>>(G, d_km.p, cm, pr, ptspl.p, plin_m.p, sp.oneloop_spl.p,
                                                            pt.scalars.p, z1, z2, loaded_mask ? d_loaded.p : nullptr,
                                                            loaded_mask, bs);
        k_gal_moments<<<dim3(GAL_NPAIR, n), 128, smem_mom, s>>>(G, d_km.p, d_stE.p, mom_band, d_momE.p, d_momJ.p, cm, pr, st,
                                                                pt.scalars.p, bs,
                                                                (corr_nw == nw && d_corr.p) ? d_corr.p + (size_t)w0*GAL_NPAIR*2*nkm : nullptr,
                                                                (ovr_nw == nw && d_ovr.p) ? d_ovr.p + (size_t)w0*GAL_NPAIR*GAL_NPAIROVR : nullptr, ps);
        k_gal_groups<<<dim3((16*nkm + 255)/256, n), 256, 0, s>>>(nkm, pr, ps, gs, gc);
        c.launches += 3;
        if (npts > 0) {
            k_gal_pkmu<<<dim3((npts + 127)/128, n), 128, 0, s>>>(G, d_km.p, pr, gs, gc, npts, d_k, d_mu, d_out + (size_t)w0*npts);
            c.launches++;
        }
        PRSD_CUDA(cudaGetLastError());
    };

    // With a host destination the walkers go in chunks: the device-to-host copy of a chunk (PCIe, the slow leg when
    // full P(k, mu) grids leave the device) runs on the copy stream under the kernels of the next chunk.
    // (only the blocking call chunks: in the streaming mode the whole copy hides under the PT stage of the NEXT batch, and
    // one launch per kernel for all walkers avoids five kernel tails per step -- measured 6.30 -> see profiles/r02_summary.md)
    const int chunk = (sync && h_out && npts > 0 && nw >= 2*GAL_PIPE_CHUNK) ? GAL_PIPE_CHUNK : nw;
    for (int w0 = 0, n = 0; w0 < nw; w0 += n) {
        // the last chunk is halved: its copy is the only one nothing hides
        const int left = nw - w0;
        n = left > chunk ? chunk : ((chunk < nw && left > chunk/2) ? chunk/2 : left);
        run_walkers(w0, n);
        if (h_out && npts > 0) {
            if (!copy_stream) {
                PRSD_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
                PRSD_CUDA(cudaEventCreateWithFlags(&copy_event, cudaEventDisableTiming));
            }
            PRSD_CUDA(cudaEventRecord(copy_event, s));
            PRSD_CUDA(cudaStreamWaitEvent(copy_stream, copy_event, 0));
            PRSD_CUDA(cudaMemcpyAsync(h_out + (size_t)w0*npts, d_out + (size_t)w0*npts, sizeof(double)*(size_t)n*npts,
                                      cudaMemcpyDeviceToHost, copy_stream));
        }
    }
    if (h_out && npts > 0) {
        cudaEvent_t& done = copy_done[copy_ticket & 1];
        if (!done) PRSD_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        PRSD_CUDA(cudaEventRecord(done, copy_stream));
        copy_ticket++;
    }
    if (!sync && ev_par_free) PRSD_CUDA(cudaEventRecord(ev_par_free, s));
    if (sync) wait();
}

void GalaxyModel::set_mu_corrections(int nw, const double* h_corr)
{
    if (!h_corr || nw <= 0) { corr_nw = 0; return; }
    d_corr.upload(h_corr, (size_t)nw*GAL_NPAIR*2*nkm, ctx->stream);
    ctx->sync();
    corr_nw = nw;
}

void GalaxyModel::set_pair_overrides(int nw, const double* h_ovr)
{
    if (!h_ovr || nw <= 0) { ovr_nw = 0; return; }
    d_ovr.upload(h_ovr, (size_t)nw*GAL_NPAIR*GAL_NPAIROVR, ctx->stream);
    ctx->sync();
    ovr_nw = nw;
}

void GalaxyModel::set_loaded(unsigned mask, const double* h_loaded)
{
    if (!h_loaded || !mask) { loaded_mask = 0; return; }
    d_loaded.upload(h_loaded, (size_t)GAL_NLOAD*nkm, ctx->stream);
    ctx->sync();
    loaded_mask = mask & 15u;
}

void GalaxyModel::wait()
{
    ctx->sync();
    if (copy_stream) PRSD_CUDA(cudaStreamSynchronize(copy_stream));
}

void GalaxyModel::wait_copy(long long ticket)
{
    PRSD_REQUIRE(ticket >= 1 && ticket <= copy_ticket, "GalaxyModel::wait_copy: unknown ticket %lld (last %lld)", ticket, copy_ticket);
    if (ticket + 2 <= copy_ticket) return;          // two later copies have been issued on the same stream since: long done
    PRSD_CUDA(cudaEventSynchronize(copy_done[(ticket - 1) & 1]));
}

void GalaxyModel::poles(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par, const double* stoch,
                        int nk, const double* d_k, int nell, const int* ells, int nmu, double* d_out)
{
    PRSD_REQUIRE(nk > 0 && nell > 0 && nell <= 8, "GalaxyModel::poles: need 1..8 multipoles and a non-empty k list");
    PRSD_REQUIRE(nmu >= 3 && (nmu & 1), "GalaxyModel::poles: Simpson's rule needs an odd number (>= 3) of mu samples");
    for (int e = 0; e < nell; e++) PRSD_REQUIRE(ells[e] >= 0 && ells[e] <= 16, "multipole %d unsupported", ells[e]);
    // every table, moment spline and group spline of the walkers (no (k, mu) evaluation)
    power(sp, pt, nw, cmap, par, stoch, 0, nullptr, nullptr, nullptr);
    Context& c = *ctx;
    DevBuf<int> d_ells;
    d_ells.upload(ells, nell, c.stream);
    ProfScope prof_gal(c, PROF_GALAXY);
    k_gal_poles<<<dim3((nk + 127)/128, nw), 128, 0, c.stream>>>(last_G, d_km.p, d_par.p, grpspl.p, grpconst.p, nk, d_k, nell,
                                                                d_ells.p, nmu, d_out);
    PRSD_CUDA(cudaGetLastError());
    c.launches++;
    c.sync();
}

void GalaxyModel::chi2(SpectraBatch& sp, PTResult& pt, int nw, const int* cmap, const double* par, const double* stoch,
                       int nk, const double* d_k, int nell, const int* ells, int nmu, const double* d_data,
                       const double* d_cinv, double* d_poles, double* d_out)
{
    poles(sp, pt, nw, cmap, par, stoch, nk, d_k, nell, ells, nmu, d_poles);
    chi2_rows(*ctx, nw, nell*nk, d_poles, d_data, d_cinv, d_out);
    ctx->sync();
}

void chi2_rows(Context& c, int nw, int n, const double* d_model, const double* d_data, const double* d_cinv, double* d_out)
{
    const size_t smem = (size_t)n*sizeof(double);
    PRSD_REQUIRE(smem <= 200*1024, "chi2: data vector of %d entries too long", n);
    if (smem > 48*1024) PRSD_CUDA(cudaFuncSetAttribute(k_gal_chi2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof_gal(c, PROF_GALAXY);
    k_gal_chi2<<<nw, 256, smem, c.stream>>>(n, d_model, d_data, d_cinv, d_out);
    PRSD_CUDA(cudaGetLastError());
    c.launches++;
}

void GalaxyModel::base_to_host(int nw, double* out)
{
    PRSD_REQUIRE(nw > 0 && nw <= last_nw, "GalaxyModel::base_to_host: %d walkers requested, the last call had %d", nw, last_nw);
    base.download(out, (size_t)nw*NB*nkm, ctx->stream);
    ctx->sync();
}

int GalaxyModel::n_base() { return NB; }

void GalaxyModel::moments_to_host(int nw, double* out)
{
    std::vector<double> tmp((size_t)nw*GAL_NPAIR*16*nkm);
    pairspl.download(tmp.data(), tmp.size(), ctx->stream);
    ctx->sync();
    for (int w = 0; w < nw; w++)
        for (int p = 0; p < GAL_NPAIR; p++)
            for (int m = 0; m < 4; m++)
                for (int i = 0; i < nkm; i++)
                    out[(((size_t)w*GAL_NPAIR + p)*4 + m)*nkm + i] = tmp[((size_t)w*GAL_NPAIR + p)*16*nkm + (size_t)(m*4)*nkm + i];
}

} // namespace prsd
