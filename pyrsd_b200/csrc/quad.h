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

struct QuadConfig {
    double qmin     = 1e-8;   // lower cut of the outer variable [h/Mpc]
    int    n_outer  = 12;     // Gauss-Legendre points per outer panel
    int    n_inner  = 8;      // Gauss-Legendre points per inner panel
    double ln_far   = 2.0;    // max panel width in ln q, featureless power-law regions
    double ln_mid   = 0.5;    // max panel width in ln q for mid_lo <= q <= mid_hi
    double mid_lo   = 1e-3;
    double mid_hi   = 10.0;
    double bao_lo   = 0.01;   // panels at most bao_dq wide in q (not ln q) in the BAO range
    double bao_hi   = 0.7;
    double bao_dq   = 0.05;
    double ln_inner = 0.5;    // max panel width in ln b of the inner integral (any region)
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

} // namespace prsd
