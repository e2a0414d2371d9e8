// quad.cpp -- host-side quadrature geometry (see quad.h).
#include "quad.h"
#include "pt_math.cuh"

#include <algorithm>
#include <cmath>

namespace prsd {

// ------------------------------------------------------------------ knots
KnotGrid make_knot_grid()
{
    // (lo, hi, target spacing in ln q).  The BAO zone must resolve the wiggles of P_L with a
    // cubic read (<~ 4e-7), the smooth zones only power laws.
    struct Z { double lo, hi, h; };
    const Z zs[] = {
        {1e-8, 1e-5, 0.03}, {1e-5, 5e-3, 0.02}, {5e-3, 1.0, 0.004},
        {1.0, 10.0, 0.01},  {10.0, 100.0, 0.02}, {100.0, 1.2e5, 0.03},
    };
    KnotGrid g;
    for (const Z& z : zs) {
        const double L = std::log(z.hi/z.lo);
        int n = (int)std::ceil(L/z.h - 1e-9);
        if (n < 3) n = 3;
        KnotZone kz;
        kz.x0 = std::log(z.lo);
        kz.h = L/n;
        kz.n = n;
        kz.start = (int)g.x.size();
        for (int i = 0; i <= n; i++) g.x.push_back(kz.x0 + kz.h*i);
        g.x.back() = std::log(z.hi);
        g.zones.push_back(kz);
    }
    return g;
}

int KnotGrid::stencil_start(int z, int c) const
{
    const KnotZone& kz = zones[z];
    int j = c - 1;
    if (j < 0) j = 0;
    if (j > kz.n - 3) j = kz.n - 3;
    return kz.start + j;
}

void KnotGrid::locate(double lnq, int& j0, double& s) const
{
    int z = 0;
    const int nz = (int)zones.size();
    while (z < nz - 1 && lnq >= zones[z + 1].x0) ++z;
    const KnotZone& kz = zones[z];
    int c = (int)std::floor((lnq - kz.x0)/kz.h);
    if (c < 0) c = 0;
    if (c > kz.n - 1) c = kz.n - 1;
    j0 = stencil_start(z, c);
    s = (lnq - kz.x0)/kz.h - (double)(j0 - kz.start);
}

// ------------------------------------------------------- Gauss-Legendre
void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w)
{
    x.assign(n, 0.0);
    w.assign(n, 0.0);
    for (int i = 0; i < (n + 1)/2; i++) {
        double z = std::cos(M_PI*(i + 0.75)/(n + 0.5));
        double pp = 1.0;
        for (int it = 0; it < 100; it++) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; j++) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0*j - 1.0)*z*p2 - (j - 1.0)*p3)/j;
            }
            pp = n*(z*p1 - p2)/(z*z - 1.0);
            double dz = p1/pp;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        x[i] = -z;
        x[n - 1 - i] = z;
        w[i] = w[n - 1 - i] = 2.0/((1.0 - z*z)*pp*pp);
    }
}

// --------------------------------------------------------------- panels
static double local_width(double q, const QuadConfig& c)
{
    double w = c.ln_far;
    if (q >= c.mid_lo && q <= c.mid_hi) w = std::min(w, c.ln_mid);
    if (q >= c.bao_lo && q <= c.bao_hi) w = std::min(w, c.bao_dq/q);
    return w;
}

std::vector<double> panel_edges(double lo, double hi, const std::vector<double>& extra,
                                const QuadConfig& c, double width_cap)
{
    std::vector<double> brk;
    brk.push_back(lo);
    brk.push_back(hi);
    auto add = [&](double e) { if (e > lo*(1 + 1e-12) && e < hi*(1 - 1e-12)) brk.push_back(e); };
    for (double e : extra) add(e);
    add(c.mid_lo); add(c.mid_hi); add(c.bao_lo); add(c.bao_hi);
    std::sort(brk.begin(), brk.end());
    brk.erase(std::unique(brk.begin(), brk.end()), brk.end());

    std::vector<double> out;
    out.push_back(brk[0]);
    std::vector<double> xs;
    for (size_t i = 0; i + 1 < brk.size(); i++) {
        const double p0 = brk[i], p1 = brk[i + 1];
        const double x0 = std::log(p0), x1 = std::log(p1);
        if (!(x1 > x0)) continue;
        // march with the local width, then stretch the marks to land exactly on p1
        xs.clear();
        xs.push_back(x0);
        while (xs.back() < x1 - 1e-12) {
            double q = std::min(std::exp(xs.back())*1.0000001, p1);
            xs.push_back(xs.back() + std::min(local_width(q, c), width_cap));
        }
        const double scale = (x1 - x0)/(xs.back() - x0);
        for (size_t j = 1; j < xs.size(); j++) {
            double e = (j + 1 == xs.size()) ? p1 : std::exp(x0 + (xs[j] - x0)*scale);
            out.push_back(e);
        }
    }
    return out;
}

// ----------------------------------------------------------- 2-D nodes
void build_half_nodes(const double* k, int nk, double qmax, const QuadConfig& cfg,
                      const KnotGrid& grid, HalfNodes& H)
{
    // Gauss-Legendre rules by order (a handful of distinct orders over the tiers)
    std::vector<std::vector<double>> glx(65), glw(65);
    auto rule = [&](int n) { if (glx[n].empty()) gauss_legendre(n, glx[n], glw[n]); };
    for (int i = 0; i < QuadConfig::NTIER; i++) { rule(cfg.tier[i].n_outer); rule(cfg.tier[i].n_inner); rule(cfg.tier[i].n_thin); }
    const std::vector<double> zone_breaks = {1e-5, 10.0, 100.0};

    H = HalfNodes();
    H.k.assign(k, k + nk);
    H.outer_off.push_back(0);
    H.node_off.push_back(0);
    std::vector<double> extra;
    double coef[4];
    for (int ik = 0; ik < nk; ik++) {
        const double kk = k[ik];
        const double umax = 2.0*qmax/kk;
        const QuadTier& T = cfg.tier_of(kk);
        const std::vector<double>& xo = glx[T.n_outer];
        const std::vector<double>& wo = glw[T.n_outer];
        // a = min(q, r) <= (q + r)/2 <= qmax
        extra = zone_breaks;
        extra.push_back(0.5*kk);
        extra.push_back(kk);
        extra.push_back(2.0*kk);
        // the cut u <= umax makes the inner upper limit min(a + k, 2 qmax - a): kink at a = qmax - k/2
        if (qmax - 0.5*kk > 0) extra.push_back(qmax - 0.5*kk);
        // a hard edge of a spectrum at b = z (one-loop tables vanish outside [1e-5, 10]) is crossed by
        // the inner limits k - a and a + k at a = k - z and a = z - k
        for (double z : zone_breaks) {
            if (z - kk > 0) extra.push_back(z - kk);
            if (kk - z > 0) extra.push_back(kk - z);
        }
        std::vector<double> oe = panel_edges(cfg.qmin, qmax, extra, cfg);
        for (size_t p = 0; p + 1 < oe.size(); p++) {
            const double l0 = std::log(oe[p]), l1 = std::log(oe[p + 1]);
            const double hw = 0.5*(l1 - l0), mid = 0.5*(l1 + l0);
            for (size_t io = 0; io < xo.size(); io++) {
                const double x = mid + hw*xo[io];
                const double wx = hw*wo[io];
                const double a = std::exp(x);
                double blo, bhi;
                if (a < 0.5*kk) { blo = kk - a; bhi = kk + a; }
                else            { blo = a;      bhi = a + kk; }
                // u = (a+b)/k <= umax  ->  b <= k umax - a: the inner range ends there
                const double bcap = kk*umax - a;
                if (bcap <= blo) continue;
                if (bcap < bhi) bhi = bcap;
                const int32_t oid = (int32_t)H.a.size();
                int j0; double s;
                grid.locate(x, j0, s);
                lagrange4(s, coef);
                H.a.push_back(a);
                H.a_j0.push_back(j0);
                H.a_coef.insert(H.a_coef.end(), coef, coef + 4);
                H.o_kid.push_back(ik);
                std::vector<double> ie = panel_edges(blo, bhi, zone_breaks, cfg, cfg.ln_inner);
                for (size_t pb = 0; pb + 1 < ie.size(); pb++) {
                    const double b0 = ie[pb], b1 = ie[pb + 1];
                    const double hb = 0.5*(b1 - b0), mb = 0.5*(b1 + b0);
                    const bool thin = std::log(b1/b0) < cfg.ln_thin;
                    const std::vector<double>& xq = glx[thin ? T.n_thin : T.n_inner];
                    const std::vector<double>& wq = glw[thin ? T.n_thin : T.n_inner];
                    for (size_t ii = 0; ii < xq.size(); ii++) {
                        const double b = mb + hb*xq[ii];
                        const double w = wx*(hb*wq[ii]/kk)*(2.0/kk)*a*a*b;
                        grid.locate(std::log(b), j0, s);
                        lagrange4(s, coef);
                        H.b.push_back(b);
                        H.wgt.push_back(w);
                        H.b_j0.push_back(j0);
                        H.b_coef.insert(H.b_coef.end(), coef, coef + 4);
                        H.n_outer.push_back(oid);
                    }
                }
            }
        }
        H.outer_off.push_back((int64_t)H.a.size());
        H.node_off.push_back((int64_t)H.b.size());
    }
}

// Interpolation map from the kept points of a k list to all of it: 8-point Lagrange in ln k for the points left out,
// the identity for the kept ones.  idx / w [n][8]; ksub = the kept k.
// The stencil of a point is chosen QUASI-UNIFORM at the spacing h of the interval that brackets it -- the kept point
// nearest to x_c + (m - 3.5) h, m = 0..7 -- and falls back to the 8 nearest kept points at the ends of the list.  Where
// the kept set changes density abruptly, the 8 nearest kept points would be 4 coarse + 4 fine ones, and Lagrange
// interpolation on such a stencil has a Lebesgue constant of up to 21 (see thin_graded).
void lagrange_fill_map(const std::vector<double>& kfull, const std::vector<char>& keep, std::vector<int>& idx,
                              std::vector<double>& w, std::vector<double>& ksub)
{
    const int NP = 8, n = (int)kfull.size();
    std::vector<double> xsub;
    std::vector<int> pos;
    ksub.clear();
    for (int i = 0; i < n; i++) if (keep[i]) { ksub.push_back(kfull[i]); xsub.push_back(std::log(kfull[i])); pos.push_back(i); }
    const int nsub = (int)ksub.size();
    idx.assign((size_t)n*NP, 0);
    w.assign((size_t)n*NP, 0.0);
    for (int i = 0, j = 0; i < n; i++) {
        while (j < nsub && pos[j] < i) j++;
        if (keep[i]) { idx[(size_t)i*NP] = j; w[(size_t)i*NP] = 1.0; continue; }
        const double x = std::log(kfull[i]);
        int sel[NP];
        bool ok = j >= 1 && j < nsub && nsub >= NP;
        if (ok) {
            const double h = xsub[j] - xsub[j - 1], xc = 0.5*(xsub[j] + xsub[j - 1]);
            ok = xc - 3.5*h >= xsub[0] - 1e-9 && xc + 3.5*h <= xsub[nsub - 1] + 1e-9;
            int last = -1;
            for (int m = 0; m < NP && ok; m++) {
                const double t = xc + (m - 3.5)*h;
                const int c = (int)(std::lower_bound(xsub.begin(), xsub.end(), t) - xsub.begin());
                int best = c;
                if (c >= nsub || (c > 0 && t - xsub[c - 1] <= xsub[c] - t)) best = c - 1;
                if (best <= last) best = last + 1;
                if (best >= nsub) { ok = false; break; }
                sel[m] = last = best;
            }
        }
        if (!ok) {
            const int a = std::min(std::max(j - NP/2, 0), std::max(nsub - NP, 0));
            for (int m = 0; m < NP; m++) sel[m] = std::min(a + m, nsub - 1);
        }
        for (int l = 0; l < NP; l++) {
            double wl = 1.0;
            for (int m = 0; m < NP; m++) if (m != l) wl *= (x - xsub[sel[m]])/(xsub[sel[l]] - xsub[sel[m]]);
            idx[(size_t)i*NP + l] = sel[l];
            w[(size_t)i*NP + l] = wl;
        }
    }
}

// Thin a uniformly (log-)spaced stretch of a k list away from index `from` (kept) in direction dir = -1 / +1 down to one
// point in `max_gap`.  The gaps GROW gradually (1 2 2 3 3 4 5 6 7 8 10 12): where the kept set jumps from every point to
// every 6th, the 8 nearest kept points of a left-out point are 4 fine + 4 coarse ones, and Lagrange interpolation on such a
// stencil has a Lebesgue constant of 21 -- it amplified the point-to-point quadrature noise of the computed values (4e-7)
// to 9e-6 in K01 at k = 0.0086.  Graded (and with the stencil choice of lagrange_fill_map) the constant is 3.2 and the
// interpolation error of the tight reference values themselves is 1.4e-7 of max(|value|, P_L): the fine points sit where
// the functions start to bend.  (A quasi-uniform stencil on the ungraded set: constant 1.5, but 1.7e-6.)
void thin_graded(std::vector<char>& keep, int from, int dir, int max_gap)
{
    static const int gaps[] = {1, 2, 2, 3, 3, 4, 5, 6, 7, 8, 10, 12};
    const int n = (int)keep.size();
    for (int i = from + dir; i >= 0 && i < n; i += dir) keep[i] = 0;
    int i = from;
    for (int s = 0;; s++) {
        const int g = std::min(max_gap, gaps[std::min(s, 11)]);
        i += dir*g;
        if (i < 0 || i >= n) break;
        keep[i] = 1;
    }
    keep[0] = keep[n - 1] = 1;
}

// The points of a long ascending k list on which the mode-coupling operators are integrated (PTPlan): below k = 0.01 one
// point in 0.29 of ln k, graded; everything for short, unsorted or non-uniform lists.
std::vector<char> thin_k_list(const std::vector<double>& k)
{
    const int nk = (int)k.size();
    std::vector<char> keep(nk, 1);
    bool ascending = nk >= 64;
    for (int i = 1; i < nk && ascending; i++) ascending = k[i] > k[i - 1];
    if (!ascending) return keep;
    int lo = 0;
    while (lo < nk - 1 && k[lo] < 0.01) lo++;
    const double dln = lo > 1 ? std::log(k[lo]/k[0])/lo : 1.0;
    const int max_gap = std::max(1, (int)std::ceil(0.29/dln - 1e-9));
    bool uniform = lo > 8;
    for (int i = 1; i <= lo && uniform; i++) uniform = std::fabs(std::log(k[i]/k[i - 1])/dln - 1.0) < 1e-6;
    if (uniform && max_gap > 1) thin_graded(keep, lo, -1, max_gap);
    int kept = 0;
    for (int i = 0; i < nk; i++) kept += keep[i];
    if (kept < 16) keep.assign(nk, 1);
    return keep;
}

} // namespace prsd
