// quad.h -- host-side (cosmology-independent) quadrature geometry for the PT hot path.
//
// The reference integrates every mode-coupling integral adaptively (Genz-Malik cubature in
// (ln u, v), pyRSD/_gcl/include/Quadrature.inl:310-523; limits in Imn.cpp:120-140,
// Kmn.cpp:118-138, ImnOneLoop.cpp:82-93).  Here the same integrals are written in the
// variables (ln a, b), a = min(q, |k-q|), b = max(q, |k-q|):
//
//   \int dlnu \int_0^1 dv  u q r P(q) P(r) f(u, v)
//        = \int dln a \int_{max(a, k-a)}^{a+k} db  (2/k^2) a^2 b P(a) P(b) f(u, +-v'),
//   u = (a+b)/k, v' = (b-a)/k,
//
// in which the integrable corner q -> 0 of the reference's variables is regular, so a
// fixed composite Gauss-Legendre rule converges geometrically.  Panels are refined where
// the linear spectrum has structure (turn-over, BAO wiggles) in BOTH a and b.  The node
// set depends only on k and the configuration, never on the cosmology, which is what
// allows the per-node kernel values to be folded into a precomputed operator.
#pragma once
#include <cstdint>
#include <vector>

namespace prsd {

// Gauss-Legendre orders as a function of k.  Parity is measured relative to max(|value|, P_L(k)) (the reference
// itself stops on epsabs = epsrel P_L(k)/V, Imn.cpp:129): below k ~ 0.02 every mode-coupling integral is < 1e-3 of
// P_L(k), so a low-order rule already has an error < 1e-6 P_L(k); the orders grow with k up to the range where
// 2 I + 6 k^2 J P_L cancels to a few per cent (k > 0.3).  A tier applies to k < k_below.
struct QuadTier { double k_below; int n_outer, n_inner, n_thin; };

struct QuadConfig {
    double qmin     = 1e-8;   // lower cut of the outer variable [h/Mpc]
    double ln_far   = 2.0;    // max panel width in ln q, featureless power-law regions
    double ln_mid   = 0.5;    // max panel width in ln q for mid_lo <= q <= mid_hi
    double mid_lo   = 1e-3;
    double mid_hi   = 10.0;
    double bao_lo   = 0.01;   // panels at most bao_dq wide in q (not ln q) in the BAO range
    double bao_hi   = 0.7;
    double bao_dq   = 0.05;
    double ln_inner = 0.5;    // max panel width in ln b of the inner integral (any region)
    double ln_thin  = 0.2;    // inner panels narrower than this in ln b use n_thin points: far from q ~ k the
                              // strip |q - r| <= k is thin and the integrand nearly polynomial across it
    // I_mn, K_mn and the one-loop tables (q_max = 1e5)
    // The top orders exist because 2 I + 6 k^2 J P_L cancels to a few per cent above k ~ 0.3; they are kept up to the
    // end of the one-loop grid (k = 10): k_interp runs to 1.0 and with use_P00_model / use_Pdv_model off the galaxy
    // assembly reads P_dd / P_dv / P_vv there, so the 1e-5 parity bound holds on the whole grid (round 1 dropped to
    // (8, 6, 5) above k = 0.5 for 5 % of the step and was 2e-4 there).  Above the last finite k_below of a
    // user-supplied configuration the LAST tier applies.
    // 0.5 <= k < 1: (16, 7, 5).  The one-loop tables in that band are the integrands that dominate the f32 / f33
    // cross parts of ImnOneLoop, which change sign near k = 0.5 and 0.86: with (12, 7, 5) the tables are good to
    // 5e-6 but those integrals are off by 1.4e-4 of |value| at the crossing (dense fixtures, profiles/r02_parity.md).
    static const int NTIER = 6;
    QuadTier tier[NTIER]    = {{0.02, 5, 4, 3}, {0.08, 6, 5, 4}, {0.3, 8, 6, 5}, {0.5, 12, 7, 5}, {1.0, 16, 7, 5}, {1e30, 12, 7, 5}};
    // ImnOneLoop (q_max = 10): the one-loop spectra are large at the hard cut and the h02-type kernels cancel to
    // leading order across v, so above k = 0.06 that operator keeps full-order panels everywhere
    // 0.2 <= k < 0.45 uses 9 inner points: the f32 cross part (P_L x P_vv) is the worst-conditioned integral of the
    // set there (8.7e-6 ... 1.2e-5 of its scale with (10, 8, 8) on the five dense-fixture cosmologies, 7e-6 with 9)
    QuadTier ml_tier[NTIER] = {{0.02, 4, 4, 3}, {0.06, 6, 5, 4}, {0.2, 10, 7, 7}, {0.45, 10, 9, 9}, {1e30, 10, 8, 8}, {1e30, 10, 8, 8}};

    const QuadTier& tier_of(double k) const
    {
        for (int i = 0; i < NTIER; i++) if (k < tier[i].k_below) return tier[i];
        return tier[NTIER - 1];
    }
    QuadConfig ml_variant() const
    {
        QuadConfig q = *this;
        for (int i = 0; i < NTIER; i++) q.tier[i] = ml_tier[i];
        return q;
    }
    void set_uniform(int n_outer, int n_inner, int n_thin)
    {
        for (int i = 0; i < NTIER; i++) { tier[i].n_outer = ml_tier[i].n_outer = n_outer; tier[i].n_inner = ml_tier[i].n_inner = n_inner;
                                      tier[i].n_thin = ml_tier[i].n_thin = n_thin; }
    }
    bool operator==(const QuadConfig& o) const
    {
        bool same = qmin == o.qmin && ln_far == o.ln_far && ln_mid == o.ln_mid && mid_lo == o.mid_lo && mid_hi == o.mid_hi &&
                    bao_lo == o.bao_lo && bao_hi == o.bao_hi && bao_dq == o.bao_dq && ln_inner == o.ln_inner && ln_thin == o.ln_thin;
        for (int i = 0; i < NTIER && same; i++)
            for (const QuadTier* p : {tier, ml_tier}) {
                const QuadTier* q = (p == tier) ? o.tier : o.ml_tier;
                same = same && p[i].k_below == q[i].k_below && p[i].n_outer == q[i].n_outer && p[i].n_inner == q[i].n_inner && p[i].n_thin == q[i].n_thin;
            }
        return same;
    }
};

// Tabulation grid of every spectrum: piecewise-uniform in x = ln q, zone edges at the hard
// cut-offs used by the reference (1e-5 = QMIN / KMIN, 10 = KMAX of the one-loop splines,
// 100 = upper limit of the sigma_v / X / Y integrals).
struct KnotZone { double x0, h; int n; int start; };   // n cells, n+1 knots from `start`
struct KnotGrid {
    std::vector<double> x;          // ln q of every knot
    std::vector<KnotZone> zones;
    int size() const { return (int)x.size(); }
    // stencil start index j0 and fractional coordinate s in [0,3] for a 4-point Lagrange read
    void locate(double lnq, int& j0, double& s) const;
    // the cell (zone z, cell c) -> 4-point stencil start used for any x inside it
    int stencil_start(int z, int c) const;
};
KnotGrid make_knot_grid();

void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w);

// Panel edges between lo and hi (in q), honouring `extra` break points.
std::vector<double> panel_edges(double lo, double hi, const std::vector<double>& extra,
                                const QuadConfig& cfg, double width_cap = 1e30);

// One half (b >= a) of the 2-D integration domain for a list of k.
struct HalfNodes {
    // per k
    std::vector<double>  k;
    std::vector<int64_t> outer_off;     // nk+1
    std::vector<int64_t> node_off;      // nk+1
    // per outer node
    std::vector<double>  a;
    std::vector<int32_t> a_j0;          // table stencil start for P(a)
    std::vector<double>  a_coef;        // 4 per outer node
    std::vector<int32_t> o_kid;         // k index of the outer node
    // per node
    std::vector<double>  b;
    std::vector<double>  wgt;           // (2/k) a^2 b dln(a) dv'   (0 beyond u_max)
    std::vector<int32_t> b_j0;
    std::vector<double>  b_coef;        // 4 per node
    std::vector<int32_t> n_outer;       // GLOBAL outer index of the node
    int64_t n_nodes() const { return (int64_t)b.size(); }
    int64_t n_outer_nodes() const { return (int64_t)a.size(); }
};

// qmax: the reference's QMAX (1e5 for Imn/Kmn, 10 for ImnOneLoop); u <= 2 qmax / k.
void build_half_nodes(const double* k, int nk, double qmax, const QuadConfig& cfg,
                      const KnotGrid& grid, HalfNodes& out);

// ---- thinned k lists (PTPlan): which points of a list an operator is integrated on, and the 8-point Lagrange map (in
// ln k) that fills the others; see quad.cpp
void lagrange_fill_map(const std::vector<double>& kfull, const std::vector<char>& keep, std::vector<int>& idx,
                       std::vector<double>& w, std::vector<double>& ksub);
void thin_graded(std::vector<char>& keep, int from, int dir, int max_gap);
std::vector<char> thin_k_list(const std::vector<double>& k);

} // namespace prsd
