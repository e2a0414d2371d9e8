// linear.cu -- build and apply the 1-D product-integration operators (see linear.cuh).
#include "linear.cuh"
#include "linear_math.cuh"
#include "spectra.cuh"
#include "ptx.cuh"

namespace prsd {

__constant__ double c_glx[GLN];
__constant__ double c_glw[GLN];

// One WARP per (output, knot): the lanes share the (sub-panel, Gauss point) pairs of every cell piece.  The outputs
// with an oscillatory kernel (X(r), Y(r) at r up to 1e5 Mpc/h: omega dq / 10 = 2e4 sub-panels per cell near q = 100)
// cost 1e6 kernel evaluations per weight; with one thread per weight (round 1) a handful of threads ran for 1.6 s
// while the rest of the device idled.  Same arithmetic as linear_weight() (linear_math.cuh, the CPU emulation's
// copy); only the order of the partial sums differs.
__device__ double linear_weight_warp(const LinOutDev& S, const ZonesDev& Z, int j, int lane)
{
    double total = 0.0;
    int z = 0;
    while (z < Z.nz - 1 && j >= Z.start[z + 1]) ++z;
    const int jl = j - Z.start[z];
    const int n = Z.n[z];
    const double x0 = Z.x0[z], h = Z.h[z];
    const double xlo = log(S.lo), xhi = log(S.hi);
    const bool has_break = (S.type == LK_JMN || S.type == LK_SIGMA3);
    const double omega = lin_omega(S.type, S.param);
    const int c_lo = jl - 3 > 0 ? jl - 3 : 0;
    const int c_hi = jl + 3 < n - 1 ? jl + 3 : n - 1;
    for (int c = c_lo; c <= c_hi; c++) {
        int s0 = c - 1;
        if (s0 < 0) s0 = 0;
        if (s0 > n - 3) s0 = n - 3;
        const int l = jl - s0;
        if (l < 0 || l > 3) continue;
        double xa = x0 + h*c, xb = x0 + h*(c + 1);
        if (xa < xlo) xa = xlo;
        if (xb > xhi) xb = xhi;
        if (!(xb > xa)) continue;
        const double qa = exp(xa), qb = exp(xb);
        double pe[3];
        int np = 1;
        pe[0] = qa;
        if (has_break && S.param > qa && S.param < qb) { pe[1] = S.param; np = 2; }
        pe[np] = qb;
        for (int ip = 0; ip < np; ip++) {
            const double p0 = pe[ip], p1 = pe[ip + 1];
            long long nsub = 1;
            if (omega > 0.0) {
                const double ph = omega*(p1 - p0);
                if (ph > 10.0) nsub = (long long)ceil(ph/10.0);
            }
            const double dq = (p1 - p0)/(double)nsub, hw = 0.5*dq;
            double acc = 0.0;
            const long long npts = nsub*GLN;
            for (long long idx = lane; idx < npts; idx += 32) {
                const long long is = idx/GLN;
                const int g = (int)(idx - is*GLN);
                const double q = p0 + dq*((double)is + 0.5) + hw*c_glx[g];
                const double s = (log(q) - x0)/h - (double)s0;
                double Lg[4];
                lagrange4(s, Lg);
                acc += c_glw[g]*Lg[l]*lin_kernel(S.type, S.sub, S.param, q);
            }
            total += acc*hw;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    return total*S.scale;
}

__global__ void __launch_bounds__(256)
k_linear_build(int nout, const LinOutDev* __restrict__ outs, ZonesDev Z, double* __restrict__ Omega)
{
    const int o = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= TAB_MPAD) return;
    double total = 0.0;
    if (j < TAB_M) total = linear_weight_warp(outs[o], Z, j, lane);
    if (lane == 0) Omega[(size_t)o*TAB_MPAD + j] = total;
}

void LinearOp::build(Context& ctx, const std::vector<LinOutSpec>& outs, bool sync)
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
    dim3 grid((TAB_MPAD + 7)/8, nout);
    k_linear_build<<<grid, 256, 0, ctx.stream>>>(nout, dout.p, Z, Omega.p);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    // sync = false: the caller overlaps host work with the build; `dout` returns to the stream-ordered pool, which hands it
    // out again only behind this kernel
    if (sync) ctx.sync();
}

// ---------------------------------------------------------------- apply
// out[row][o] = sum_j T[row][j] Omega[o][j] is a (rows x knots) . (knots x outputs) product: FP64 tensor-core MMAs
// (DMMA m8n8k4).  One CTA owns 16 outputs x 32 rows; its 8 warps split the knots, so every warp streams a
// contiguous 1/8 of 8 Omega rows and of 32 table rows with 16-byte loads (each lane's two values feed two
// consecutive k-steps: the order of the knots inside a sum is free as long as both operands agree), and a
// shared-memory pass adds the eight partial tiles.  Omega, the large operand, is read once per 32 rows.
static const int LA_MT = 4;                            // row tiles (of 8) per CTA
static const int LA_NT = 2;                            // output tiles (of 8) per CTA
static const int LA_JS = ((TAB_MPAD + 63)/64)*8;      // knots per warp, a multiple of 8

__global__ void __launch_bounds__(256)
k_linear_apply(int nout, int nrow, const double* __restrict__ Omega, const double* __restrict__ tables,
               const int* __restrict__ tab_index, double* __restrict__ out)
{
    __shared__ double red[8][LA_NT*LA_MT][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o0 = blockIdx.x*(8*LA_NT);
    const int r0 = blockIdx.y*(8*LA_MT);
    const int sub = lane & 3, grp = lane >> 2;
    const double* bp[LA_NT];
#pragma unroll
    for (int n = 0; n < LA_NT; n++) bp[n] = Omega + (size_t)min(o0 + 8*n + grp, nout - 1)*TAB_MPAD;
    const double* ap[LA_MT];
#pragma unroll
    for (int m = 0; m < LA_MT; m++) ap[m] = tables + (size_t)tab_index[min(r0 + 8*m + grp, nrow - 1)]*TAB_MPAD;
    double acc[LA_NT][LA_MT][2];
#pragma unroll
    for (int n = 0; n < LA_NT; n++)
#pragma unroll
        for (int m = 0; m < LA_MT; m++) acc[n][m][0] = acc[n][m][1] = 0.0;
    const int j_lo = warp*LA_JS, j_hi = min(j_lo + LA_JS, TAB_MPAD);
#pragma unroll 2
    for (int j = j_lo; j < j_hi; j += 8) {
        const int idx = j + 2*sub;
        const bool ok = idx < TAB_MPAD;
        double2 b[LA_NT], a[LA_MT];
#pragma unroll
        for (int n = 0; n < LA_NT; n++) b[n] = ok ? __ldg(reinterpret_cast<const double2*>(bp[n] + idx)) : make_double2(0.0, 0.0);
#pragma unroll
        for (int m = 0; m < LA_MT; m++) a[m] = ok ? __ldg(reinterpret_cast<const double2*>(ap[m] + idx)) : make_double2(0.0, 0.0);
#pragma unroll
        for (int n = 0; n < LA_NT; n++)
#pragma unroll
            for (int m = 0; m < LA_MT; m++) {
                dmma_m8n8k4(acc[n][m][0], acc[n][m][1], a[m].x, b[n].x);
                dmma_m8n8k4(acc[n][m][0], acc[n][m][1], a[m].y, b[n].y);
            }
    }
#pragma unroll
    for (int n = 0; n < LA_NT; n++)
#pragma unroll
        for (int m = 0; m < LA_MT; m++) {
            red[warp][n*LA_MT + m][grp*8 + 2*sub] = acc[n][m][0];
            red[warp][n*LA_MT + m][grp*8 + 2*sub + 1] = acc[n][m][1];
        }
    __syncthreads();
    // 64 elements per tile, LA_NT x LA_MT tiles
    for (int idx = threadIdx.x; idx < LA_NT*LA_MT*64; idx += 256) {
        const int t = idx >> 6, e = idx & 63;
        const int n = t/LA_MT, m = t - n*LA_MT;
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) v += red[w][t][e];
        const int row = r0 + 8*m + (e >> 3), col = o0 + 8*n + (e & 7);
        if (row < nrow && col < nout) out[(size_t)row*nout + col] = v;
    }
}

// Operators with a handful of outputs (sigma_v^2-type scalars: one output per cosmology) leave the MMA kernel with
// 8 CTAs on 148 SMs (55 us each, latency-bound: ncu, profiles/r01_ncu_others.json): one warp per (row, output) dot
// product instead.
__global__ void __launch_bounds__(256)
k_linear_dot(int nout, int nrow, const double* __restrict__ Omega, const double* __restrict__ tables,
             const int* __restrict__ tab_index, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= nrow*nout) return;
    const int row = item/nout, o = item - row*nout;
    const double2* a = reinterpret_cast<const double2*>(tables + (size_t)tab_index[row]*TAB_MPAD);
    const double2* b = reinterpret_cast<const double2*>(Omega + (size_t)o*TAB_MPAD);
    double s0 = 0.0, s1 = 0.0;
    for (int j = lane; j < TAB_MPAD/2; j += 32) {
        const double2 x = __ldg(a + j), y = __ldg(b + j);
        s0 = fma(x.x, y.x, s0);
        s1 = fma(x.y, y.y, s1);
    }
    double s = s0 + s1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) out[(size_t)row*nout + o] = s;
}

void LinearOp::apply(Context& ctx, const double* tables, const int* tab_index, int nrow, double* out) const
{
    if (nout <= 4) {
        ProfScope prof(ctx, PROF_LINEAR);
        const int items = nrow*nout;
        k_linear_dot<<<(items + 7)/8, 256, 0, ctx.stream>>>(nout, nrow, Omega.p, tables, tab_index, out);
        PRSD_CUDA(cudaGetLastError());
        ctx.launches++;
        return;
    }
    static_assert(TAB_MPAD % 2 == 0, "table rows must be 16-byte multiples");
    dim3 grid((nout + 8*LA_NT - 1)/(8*LA_NT), (nrow + 8*LA_MT - 1)/(8*LA_MT));
    ProfScope prof(ctx, PROF_LINEAR);
    k_linear_apply<<<grid, 256, 0, ctx.stream>>>(nout, nrow, Omega.p, tables, tab_index, out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
