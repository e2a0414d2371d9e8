// nodal.cuh -- "one lane per quadrature node" evaluation of the mode-coupling operators.
//
// The DMMA contraction of bilinear.cu maps (row, node) pairs onto lanes, which replicates every per-node
// load over the rows and computes all rows x columns products.  Two operators have few columns and a
// sparse row/column pattern, and are cheaper on the plain FP64 pipe with one lane per node:
//
//   * the ImnOneLoop convolutions (ImnOneLoop.cpp:96-216): per cosmology the four spectra P_L, P_dd, P_dv,
//     P_vv enter 8 (P_1, P_2) pairs x 7 kernels, of which the reference only ever uses 23; both halves of
//     the v domain are fused (the r <= q half reads the same nodes with the spectra swapped), and the
//     (linear, cross, one-loop) combinations of rsd/pt_integrals.py are formed in the epilogue;
//   * the one-loop tables (OneLoopPS.cpp:56-185): 6 symmetric kernels, P_L x P_L only, 4 cosmologies per CTA.
//
// One CTA = (tile of spectra, chunk of k); the tile's tables are staged once with TMA bulk copies; per k the
// outer values P(a) are tabulated in shared memory, then every thread walks nodes with coalesced loads of
// the node stream {packed outer index and stencil start, stencil coordinate, W columns} (structure-of-arrays; the four
// Lagrange coefficients are formed from the coordinate in registers), reads the
// four table values per spectrum from shared memory and accumulates in registers; a shuffle reduction and
// a cross-warp pass finish each k.  CTAs of the same k chunk are adjacent in launch order so that the node
// stream is read from DRAM once and from L2 afterwards.
#pragma once
#include "bilinear.cuh"

namespace prsd {

// out[c][7][3][nk] (final ImnOneLoop layout of PTResult::ML); slot_tab [ncosmo][4] = table index of
// P_L, P_dd, P_dv, P_vv of every cosmology; `op` must have been built with ml_columns()
std::vector<ColSpec> ml_columns();
void nodal_apply_ml(Context& ctx, const BilinearOp& op, const double* tables, const int* slot_tab, int ncosmo,
                    double* out);

// out[c][nk][8] (columns = the operator's kernels); slot_tab [ntiles][R] = P_L table of the R = nodal_diag_rows()
// cosmologies of every tile, ntiles = ceil(ncosmo / R) (pad the last tile with any valid table)
int nodal_diag_rows();
void nodal_apply_diag(Context& ctx, const BilinearOp& op, const double* tables, const int* slot_tab, int ncosmo,
                      double* out);

} // namespace prsd
