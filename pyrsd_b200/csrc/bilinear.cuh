// bilinear.cuh -- the mode-coupling integrals as a precomputed quadrature operator.
//
// Every 2-D integral of the reference's Imn / Kmn / ImnOneLoop drivers
// (Imn.cpp:113-140, Kmn.cpp:109-138, ImnOneLoop.cpp:74-93) has the form
//     O_c(k) = sum_nodes  W_c[node] * P_1(a_node) * P_2(b_node)
// once a fixed quadrature is chosen (quad.h).  W (measure x kernel x quadrature weight)
// depends only on k and the kernel, so it is built ONCE on the device and every cosmology /
// spectrum pair is then a row of a (rows x nodes) . (nodes x kernels) contraction, executed
// with FP64 tensor-core MMAs (DMMA m8n8k4) fed from spectra tables staged in shared memory
// by TMA bulk copies.
#pragma once
#include "common.h"
#include "quad.h"

#include <memory>

namespace prsd {

static const int TAB_M    = 2457;   // knots of make_knot_grid() (checked at start-up)
static const int TAB_MPAD = 2468;   // row stride of every spectra table: = 4 (mod 16) doubles

struct ColSpec {
    int kid;        // KernelId
    int sign;       // +1: f(u, +v') with P_1(a) P_2(b);  -1: f(u, -v') (the r <= q half);  0: their sum
    double fac;     // coefficient c in the prefactor c * k  (k/(8 pi^2) etc.)
};

// node geometry on the device, shared between operators on the same k list
struct NodeGeometry {
    int nk = 0;
    int max_outer = 0;
    int64_t n_nodes = 0, n_outer = 0;
    std::vector<double> k;
    std::vector<int64_t> h_node_off, h_outer_off;
    DevBuf<int64_t> outer_off, node_off;
    DevBuf<int> a_j0, o_kid;
    DevBuf<double> a, a_coef;
    DevBuf<int> b_j0, n_outer_local;
    DevBuf<double> b, wgt, b_coef;
    // fractional stencil coordinate s in [0, 3] of every node / outer node (the four coefficients are lagrange4(s)):
    // the nodal kernels read 8 bytes instead of 32 and spend 13 FP64 operations
    DevBuf<double> b_s, a_s;
    // compact index stream of the one-lane-per-node kernels: (outer index << 12) | stencil start
    DevBuf<int> b_pack;
    // the two per-node words of the nodal kernels in ONE 8-byte load: .x = b_pack, .y = stencil coordinate as binary32
    // (a coordinate error of 1e-7 of a knot spacing moves the interpolated value by < 1e-9)
    DevBuf<int2> b_node;
    DevBuf<double> kdev;
    size_t bytes() const;
};

// bank_mod: the node order makes the stencil starts of `bank_mod` consecutive nodes equal or distinct modulo
// bank_mod -- 16 for 8-byte table reads (half-warp wavefronts), 32 for the 4-byte reads of the nodal kernels
std::shared_ptr<NodeGeometry> make_geometry(Context& ctx, const double* k, int nk, double qmax,
                                            const QuadConfig& cfg, const KnotGrid& grid, int bank_mod = 16);

// the two halves of make_geometry: the host part touches no CUDA state and may run on any thread
struct GeometryHost;
std::shared_ptr<GeometryHost> make_geometry_host(const double* k, int nk, double qmax, const QuadConfig& cfg,
                                                 const KnotGrid& grid, int bank_mod = 16);
std::shared_ptr<NodeGeometry> upload_geometry(Context& ctx, const GeometryHost& h);
uint64_t geometry_checksum(const GeometryHost& h);

struct BilinearOp {
    std::shared_ptr<NodeGeometry> geo;
    int ncol = 0;       // padded to a multiple of 8
    int ncol_used = 0;
    bool soa = false;   // false: W [n_nodes][ncol] in fragment-major column order (DMMA kernel);
                        // true : W [ncol_used][n_nodes] (one-lane-per-node kernels, nodal.cuh)
    DevBuf<double> W;
    // Nodal (SoA) operators stream their entries as binary32 packed four columns to a float4, [ceil(ncol / 4)][n_nodes]:
    // the LSU issues one global load per ~1.8 cycles whatever its width, so 9 columns cost 3 LDG.128 instead of 9 LDG.64.
    // An entry is weight x kernel value; its rounding error (6e-8, independent from node to node) averages out over the
    // ~2000 nodes of an integral -- measured on the dense fixtures: no table or ImnOneLoop integral moves by more than 2e-8.
    DevBuf<float4> Wf;
    int nvec = 0;
    // DMMA operator: the fragments of W as binary32 in the lane order of w_index (build() converts and releases W)
    DevBuf<float> Wh;
    void to_float(Context& ctx);        // SoA only: W -> Wf, releases W
    void build(Context& ctx, std::shared_ptr<NodeGeometry> g, const std::vector<ColSpec>& cols, bool soa_layout = false);
    // SoA operators only: column `dst` += sum_i coef_i * column src_i, then keep the first `ncol_keep` columns
    void fold_columns(Context& ctx, int dst, const std::vector<std::pair<int, double>>& terms, int ncol_keep);

    // out[row][k][ncol] (device) = sum_nodes W * TA_row(a) * TB_row(b)
    //   tables   : device pool [ntab][TAB_MPAD]
    //   slot_tab : device [ntiles][8] global table index of each local slot
    //   rowA/rowB: device [ntiles][8] local slot (0..7) of the a-side / b-side table per row
    void apply(Context& ctx, const double* tables, const int* slot_tab, const int* rowA,
               const int* rowB, int ntiles, double* out) const;
};

} // namespace prsd
