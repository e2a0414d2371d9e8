// plan.cuh -- the cosmology-independent "plan" of the PT stage and its per-batch results.
//
// A plan owns, for a given k_interp grid (rsd: 250 log-spaced k in [5e-6, 1],
// pyRSD/rsd/power/dm/power_dm.py:22) and the reference's fixed one-loop grid
// (logspace(1e-5, 10, 1000), OneLoopPS.cpp:11-13) and Zel'dovich r grid (ZeldovichPS.cpp:41-53):
//   * the 2-D quadrature operators (bilinear.cuh) for I_mn, K_mn, the one-loop tables and
//     the ImnOneLoop convolutions;
//   * the 1-D operators (linear.cuh) for J_mn, sigma_v^2, sigma_v^2(k), X(r), Y(r), Q3,
//     velocity kurtosis.
// run() evaluates everything the reference's PTIntegralsMixin (rsd/pt_integrals.py:42-450)
// tabulates, for a whole batch of cosmologies, unnormalised (z = 0) exactly as there.
#pragma once
#include "bilinear.cuh"
#include "linear.cuh"
#include "spectra.cuh"

namespace prsd {

static const int ZEL_NR = 1024;     // NUM_PTS of ZeldovichPS.cpp:7
static const int N_ML = 7;          // Ivvdd_h01, Ivvdd_h02, Idvdv_h03, Idvdv_h04, Ivvvv_f23, f32, f33

std::vector<double> oneloop_kgrid();
std::vector<double> zeldovich_rgrid();

struct PTResult {
    int ncosmo = 0, nk = 0;
    uint64_t src_uid = 0, src_gen = 0;     // SpectraBatch::uid / lin_gen the tables were computed from (0: none yet)
    // all device buffers, layouts [cosmo][...]
    DevBuf<double> I;          // [c][16][nk]   I_mn, index 4 m + n
    DevBuf<double> J;          // [c][6][nk]    order J00 J01 J02 J10 J11 J20
    DevBuf<double> K;          // [c][13][nk]   order K00 K01 K00s K01s K02s K10 K11 K10s K11s K20_a K20_b K20s_a K20s_b
    DevBuf<double> ML;         // [c][7][3][nk] (linear, cross, one-loop) of the N_ML ImnOneLoop integrals
    DevBuf<double> oneloop;    // [c][4][1000]  Pdd, Pdv, Pvv, P22bar one-loop tables
    DevBuf<double> sigmasq_k;  // [c][nk]
    DevBuf<double> scalars;    // [c][4]  sigma_v^2, velocity kurtosis, Q3/k^4, (unused)
    DevBuf<double> X0;         // [c][1024]  X_Zel(r)
    DevBuf<double> YY;         // [c][1024]  Y_Zel(r)
    void ensure(int ncosmo_, int nk_);
};

struct PTPlan {
    Context* ctx;
    QuadConfig cfg;
    std::vector<double> k_interp, k_1loop, r_zel;
    int nk = 0;
    bool with_ml = true, with_zel = true, only_oneloop = false;
    BilinearOp op_1loop, op_ik, op_ml;
    LinearOp lin_all;          // [J00, J01, J11 on the one-loop subset (n_lin1 outputs) | J x6 on k_sub, sigmasq_k (nk), sigma_v^2,
                               //  X (1024), Y (1024)]   (P_L table)
    int n_lin1 = 0;
    LinearOp lin_invq2_100;    // \int S/q^2 dq over [1e-5, 100] x 1/(10 pi^2)  (P_L^2 table)  -> Q3 / k^4
    LinearOp lin_kurt;         // 0.25/(2 pi^2) \int S/q^2 dq over [1e-5, 10]   (P22bar table)
    DevBuf<double> d_k1loop, d_kinterp;
    // the subset of the one-loop grid on which the 2-D operator is integrated, and the interpolation map to all 1000 points
    std::vector<double> k_1loop_sub;
    int n_1loop_sub = 0;
    DevBuf<int> d_olmap_idx;        // [1000][8]
    DevBuf<double> d_olmap_w;       // [1000][8]
    // ... and the (larger) subset on which J00, J01, J11 are integrated
    std::vector<double> k_1loop_j;
    DevBuf<int> d_jmap_idx;         // [1000][8]
    DevBuf<double> d_jmap_w;        // [1000][8]
    // the same for the I / K and ImnOneLoop operators on k_interp (PTPlan ctor): nk_sub == nk when the list is not thinned
    std::vector<double> k_sub;
    int nk_sub = 0;
    DevBuf<int> d_kmap_idx;         // [nk][8]
    DevBuf<double> d_kmap_w;        // [nk][8]
    DevBuf<double> s_ml;            // [c][21][nk_sub]
    // scratch
    DevBuf<double> s_op1, s_lin, s_ik, s_mlp, s_mln, s_small, s_ol, s_plk;
    DevBuf<int> s_slot, s_rowA, s_rowB, s_tabidx;
    // index arrays of the current batch size (rebuilt when it changes)
    DevBuf<int> d_tabL, d_tabSq, d_tab22;
    int idx_nc_a = -1, idx_nc_b = -1;
    double build_seconds = 0.0;
    double build_stage_ms[8] = {0};    // host wall clock: 2 enqueue of the 1-D operator builds, 0 node geometry of the three 2-D
                                       // operators (host threads, concurrent with those builds) + uploads + W builds, 4 drain

    PTPlan(Context& c, const double* k, int nk_, const QuadConfig& q, bool ml, bool zel, bool only1l = false);
    // stage A only: the four one-loop tables (left in sp and in res.oneloop)
    void run_oneloop(SpectraBatch& sp, PTResult& res, bool sync = true);
    // sync = false: everything is enqueued on the context's stream and the call returns (asynchronous pipeline)
    void run(SpectraBatch& sp, PTResult& res, bool sync = true);
    size_t device_bytes() const;
    long long nodes_total() const;
};

} // namespace prsd
