// linear.cu -- build and apply the 1-D product-integration operators (see linear.cuh).
#include "linear.cuh"
#include "linear_math.cuh"
#include "spectra.cuh"
#include "ptx.cuh"

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
// out[row][o] = sum_j T[row][j] Omega[o][j] is a (rows x knots) . (knots x outputs) product: FP64 tensor-core MMAs
// (DMMA m8n8k4).  One warp owns 8 outputs and 32 rows (4 row tiles), the 8 warps of a CTA share the same rows so
// the table fragments come from L1; Omega, the large operand (nout x 2468 doubles), is read once per 32 rows.
static const int LA_MT = 4;

__global__ void __launch_bounds__(256)
k_linear_apply(int nout, int nrow, const double* __restrict__ Omega, const double* __restrict__ tables,
               const int* __restrict__ tab_index, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o0 = (blockIdx.x*8 + warp)*8;
    const int r0 = blockIdx.y*(8*LA_MT);
    if (o0 >= nout) return;
    const int sub = lane & 3, grp = lane >> 2;
    const double* bp = Omega + (size_t)min(o0 + grp, nout - 1)*TAB_MPAD + sub;
    const double* ap[LA_MT];
#pragma unroll
    for (int m = 0; m < LA_MT; m++) ap[m] = tables + (size_t)tab_index[min(r0 + 8*m + grp, nrow - 1)]*TAB_MPAD + sub;
    double acc[LA_MT][2];
#pragma unroll
    for (int m = 0; m < LA_MT; m++) acc[m][0] = acc[m][1] = 0.0;
#pragma unroll 4
    for (int j = 0; j < TAB_MPAD; j += 4) {
        const double b = __ldg(bp + j);
#pragma unroll
        for (int m = 0; m < LA_MT; m++) dmma_m8n8k4(acc[m][0], acc[m][1], __ldg(ap[m] + j), b);
    }
#pragma unroll
    for (int m = 0; m < LA_MT; m++) {
        const int row = r0 + 8*m + grp;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int col = o0 + 2*sub + e;
            if (row < nrow && col < nout) out[(size_t)row*nout + col] = acc[m][e];
        }
    }
}

void LinearOp::apply(Context& ctx, const double* tables, const int* tab_index, int nrow, double* out) const
{
    static_assert(TAB_MPAD % 4 == 0, "table rows must be a whole number of DMMA k-steps");
    dim3 grid((nout + 63)/64, (nrow + 8*LA_MT - 1)/(8*LA_MT));
    ProfScope prof(ctx, PROF_LINEAR);
    k_linear_apply<<<grid, 256, 0, ctx.stream>>>(nout, nrow, Omega.p, tables, tab_index, out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
