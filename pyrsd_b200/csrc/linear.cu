// linear.cu -- build and apply the 1-D product-integration operators (see linear.cuh).
#include "linear.cuh"
#include "linear_math.cuh"
#include "spectra.cuh"

namespace prsd {

__constant__ double c_glx[GLN];
__constant__ double c_glw[GLN];

__global__ void k_linear_build(int nout, const LinOutDev* __restrict__ outs, ZonesDev Z, double* __restrict__ Omega)
{
    const int o = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= TAB_MPAD) return;
    double total = 0.0;
    if (j < TAB_M) total = linear_weight(outs[o], Z, j, c_glx, c_glw);
    Omega[(size_t)o*TAB_MPAD + j] = total;
}

void LinearOp::build(Context& ctx, const std::vector<LinOutSpec>& outs)
{
    nout = (int)outs.size();
    PRSD_REQUIRE(nout > 0, "LinearOp: no outputs");
    std::vector<double> x, w;
    gauss_legendre(GLN, x, w);
    PRSD_CUDA(cudaMemcpyToSymbolAsync(c_glx, x.data(), GLN*sizeof(double), 0, cudaMemcpyHostToDevice, ctx.stream));
    PRSD_CUDA(cudaMemcpyToSymbolAsync(c_glw, w.data(), GLN*sizeof(double), 0, cudaMemcpyHostToDevice, ctx.stream));
    const KnotGrid& kg = knot_grid();
    ZonesDev Z;
    Z.nz = (int)kg.zones.size();
    PRSD_REQUIRE(Z.nz <= 8, "too many knot zones");
    for (int i = 0; i < Z.nz; i++) {
        Z.x0[i] = kg.zones[i].x0; Z.h[i] = kg.zones[i].h; Z.n[i] = kg.zones[i].n; Z.start[i] = kg.zones[i].start;
    }
    std::vector<LinOutDev> ho(nout);
    for (int i = 0; i < nout; i++) {
        const LinOutSpec& s = outs[i];
        PRSD_REQUIRE(s.lo > 0 && s.hi >= s.lo, "LinearOp: bad limits [%g, %g]", s.lo, s.hi);
        ho[i] = {s.type, s.sub, s.param, s.lo, s.hi, s.scale};
    }
    DevBuf<LinOutDev> dout;
    dout.upload(ho, ctx.stream);
    Omega.alloc((size_t)nout*TAB_MPAD);
    dim3 grid((TAB_MPAD + 63)/64, nout);
    k_linear_build<<<grid, 64, 0, ctx.stream>>>(nout, dout.p, Z, Omega.p);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    ctx.sync();
}

// ---------------------------------------------------------------- apply
// One warp computes 4 outputs for 8 rows; lanes stride over the knots.
static const int LA_OUT = 4, LA_ROW = 8;

__global__ void __launch_bounds__(128)
k_linear_apply(int nout, int nrow, const double* __restrict__ Omega, const double* __restrict__ tables,
               const int* __restrict__ tab_index, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o0 = (blockIdx.x*4 + warp)*LA_OUT;
    const int r0 = blockIdx.y*LA_ROW;
    if (o0 >= nout) return;
    const double* om[LA_OUT];
    const double* tb[LA_ROW];
#pragma unroll
    for (int i = 0; i < LA_OUT; i++) om[i] = Omega + (size_t)min(o0 + i, nout - 1)*TAB_MPAD;
#pragma unroll
    for (int r = 0; r < LA_ROW; r++) tb[r] = tables + (size_t)tab_index[min(r0 + r, nrow - 1)]*TAB_MPAD;
    double acc[LA_OUT][LA_ROW];
#pragma unroll
    for (int i = 0; i < LA_OUT; i++)
#pragma unroll
        for (int r = 0; r < LA_ROW; r++) acc[i][r] = 0.0;
    for (int j = lane; j < TAB_M; j += 32) {
        double w[LA_OUT], t[LA_ROW];
#pragma unroll
        for (int i = 0; i < LA_OUT; i++) w[i] = __ldg(om[i] + j);
#pragma unroll
        for (int r = 0; r < LA_ROW; r++) t[r] = __ldg(tb[r] + j);
#pragma unroll
        for (int i = 0; i < LA_OUT; i++)
#pragma unroll
            for (int r = 0; r < LA_ROW; r++) acc[i][r] = fma(w[i], t[r], acc[i][r]);
    }
#pragma unroll
    for (int i = 0; i < LA_OUT; i++)
#pragma unroll
        for (int r = 0; r < LA_ROW; r++) {
            double v = acc[i][r];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if (lane == 0 && o0 + i < nout && r0 + r < nrow) out[(size_t)(r0 + r)*nout + o0 + i] = v;
        }
}

void LinearOp::apply(Context& ctx, const double* tables, const int* tab_index, int nrow, double* out) const
{
    dim3 grid((nout + 4*LA_OUT - 1)/(4*LA_OUT), (nrow + LA_ROW - 1)/LA_ROW);
    ProfScope prof(ctx, PROF_LINEAR);
    k_linear_apply<<<grid, 128, 0, ctx.stream>>>(nout, nrow, Omega.p, tables, tab_index, out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
