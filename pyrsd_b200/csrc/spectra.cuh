// spectra.cuh -- device-resident spectra of a batch of cosmologies.
//
// The linear transfer function T(k) comes from the host (CLASS stays on the host, as in
// the reference where Cosmology owns the Boltzmann solver, Cosmology.cpp:251-281).  Per
// cosmology the device keeps
//   * the reference's cubic spline of T(k) (Spline.cpp:211-257) and the Eisenstein-Hu
//     no-wiggle extrapolation constants (Cosmology.cpp:187-246), so P_L(k) can be evaluated
//     exactly as LinearPS::Evaluate does (LinearPS.cpp:11-15);
//   * tables of every spectrum the PT operators read (P_L, the three one-loop spectra,
//     P22bar, P_L^2) on the common knot grid (quad.h), in one pool so that the bilinear
//     kernels can address any (cosmology, kind) by a single index.
#pragma once
#include "common.h"
#include "pt_math.cuh"
#include "bilinear.cuh"

namespace prsd {

enum SpectrumKind { SP_LIN = 0, SP_DD1, SP_DV1, SP_VV1, SP_P22BAR, SP_LIN_SQ, SP_NKIND };

static const int ONELOOP_N = 1000;      // NUM_PTS of OneLoopPS.cpp:13
static const double ONELOOP_KMIN = 1e-5, ONELOOP_KMAX = 1e1;

struct SpectraBatch {
    Context* ctx = nullptr;
    int ncosmo = 0;
    int nspl = 0;
    std::vector<PlinParams> h_params;
    DevBuf<PlinParams> params;     // [ncosmo]
    DevBuf<double> spl;            // [ncosmo][5*nspl]   X | Y | b | c | d
    DevBuf<double> tables;         // [ncosmo*SP_NKIND][TAB_MPAD]
    DevBuf<double> oneloop_spl;    // [ncosmo][4][5*ONELOOP_N] splines of the one-loop tables
    DevBuf<double> oneloop_raw;    // [ncosmo][4][ONELOOP_N] the tables themselves
    bool have_oneloop = false;
    // identity of the batch and of its current linear spectra: PTPlan::run stamps its result with both, and the
    // galaxy stage refuses a result computed from another batch or from spectra that have since been replaced
    uint64_t uid = 0, lin_gen = 0;

    SpectraBatch(Context& c, int ncosmo_, int nspl_, const double* k, const double* T, const double* pars);
    // rebuild every linear-spectrum table from DEVICE-resident inputs: d_k, d_T [ncosmo][nspl],
    // d_pars [ncosmo][8] (same meaning as the host constructor); asynchronous on the context's stream
    void reload_device(const double* d_k, const double* d_T, const double* d_pars);
    // the same from HOST inputs (same shapes as the constructor's), asynchronous: the inputs are copied into
    // persistent device staging and nothing waits for the stream.  For the call to return before the copies have
    // run the host buffers must be page-locked, and they must stay unchanged until the stream has passed them.
    void reload_host(const double* k, const double* T, const double* pars);
    DevBuf<double> in_k, in_T, in_pars;
    // reload_host copies on the context's upload stream: ev_in_ready = the staging holds the new inputs (the main stream
    // waits for it), ev_in_free = the kernels reading the staging have run (the next upload waits for it)
    cudaEvent_t ev_in_ready = nullptr, ev_in_free = nullptr;
    ~SpectraBatch();
    int table_index(int cosmo, int kind) const { return cosmo*SP_NKIND + kind; }
    const double* table_ptr(int cosmo, int kind) const { return tables.p + (size_t)table_index(cosmo, kind)*TAB_MPAD; }
    // multiply P_L (spline amplitude and tables) of cosmology c by fac
    void rescale_linear(const double* fac_host);
    void eval_linear(int nq, const double* q_host, double* out_host);    // [ncosmo][nq]
    // install the four one-loop tables (device, [ncosmo][4][ONELOOP_N]), build their splines
    // and tabulate them on the knot grid
    void set_oneloop_tables(const double* d_tables);
    // replace ONE table (host, [ncosmo][ONELOOP_N]); the other three keep their values (0 if never set)
    void set_oneloop_table(int which, const double* h_table);
    void get_oneloop_tables(double* h_out);      // [ncosmo][4][ONELOOP_N]
    void eval_oneloop(int which, bool full, int nq, const double* q_host, double* out_host);
};

const KnotGrid& knot_grid();             // the process-wide knot grid (host)
const double* knot_grid_dev(Context&);   // ln q of every knot on the device [TAB_M]

// batched netlib cubic-spline build: Y [nspline][n] -> spl [nspline][5n]; X is [n] if
// x_stride == 0 (shared abscissae) else [nspline][n]
// d_invb (optional): factors of the abscissae from spline_factor(), [1][n] if x_stride == 0 else [nspline][n]
void spline_build_batched(Context& ctx, int nspline, int n, const double* dX, int x_stride,
                          const double* dY, double* dspl, const double* d_invb = nullptr);
void spline_factor(Context& ctx, int nfac, int n, const double* dX, int x_stride, double* d_invb);

} // namespace prsd
