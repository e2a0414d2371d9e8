// linear.cu -- build and apply the 1-D product-integration operators (see linear.cuh).
#include "linear.cuh"
#include "linear_math.cuh"
#include "spectra.cuh"
#include "ptx.cuh"

namespace prsd {

__constant__ double c_glx[GLN];
__constant__ double c_glw[GLN];

// One WARP per (output, CELL of the knot grid): the lanes share the (sub-panel, Gauss point) pairs of every piece of the
// cell and accumulate the integrals of all FOUR cardinal functions of the cell's stencil at once -- the kernel value and
// log q, which are the cost, are computed once per point and not once per knot whose stencil holds the cell (4 x fewer
// evaluations than one warp per (output, knot), which the first version of round 2 used).  The outputs with an
// oscillatory kernel (X(r), Y(r) at r up to 1e5 Mpc/h: omega dq / 10 = 2e4 sub-panels per cell near q = 100) cost 3e5
// kernel evaluations per cell; with one thread per weight (round 1) a handful of threads ran for 1.6 s while the rest of
// the device idled.  Same arithmetic as linear_weight() (linear_math.cuh, the CPU emulation's copy); only the order of
// the partial sums differs.  k_linear_gather then adds the <= 4 cell integrals of every knot in ascending cell order.
struct CellsDev { int ncell; int cstart[8]; };      // cstart[z]: first global cell of zone z

__global__ void __launch_bounds__(256)
k_linear_cells(int o_first, const LinOutDev* __restrict__ outs, ZonesDev Z, CellsDev C, double* __restrict__ part /*[o][cell][4]*/)
{
    const int o = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int gc = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gc >= C.ncell) return;
    const LinOutDev S = outs[o_first + o];
    int z = 0;
    while (z < Z.nz - 1 && gc >= C.cstart[z + 1]) ++z;
    const int c = gc - C.cstart[z];
    const int n = Z.n[z];
    const double x0 = Z.x0[z], h = Z.h[z];
    const double xlo = log(S.lo), xhi = log(S.hi);
    const bool has_break = (S.type == LK_JMN || S.type == LK_SIGMA3);
    const double omega = lin_omega(S.type, S.param);
    int s0 = c - 1;
    if (s0 < 0) s0 = 0;
    if (s0 > n - 3) s0 = n - 3;
    double total[4] = {0.0, 0.0, 0.0, 0.0};
    double xa = x0 + h*c, xb = x0 + h*(c + 1);
    if (xa < xlo) xa = xlo;
    if (xb > xhi) xb = xhi;
    if (xb > xa) {
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
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const long long npts = nsub*GLN;
            for (long long idx = lane; idx < npts; idx += 32) {
                const long long is = idx/GLN;
                const int g = (int)(idx - is*GLN);
                const double q = p0 + dq*((double)is + 0.5) + hw*c_glx[g];
                const double s = (log(q) - x0)/h - (double)s0;
                double Lg[4];
                lagrange4(s, Lg);
                const double wk = c_glw[g]*lin_kernel(S.type, S.sub, S.param, q);
#pragma unroll
                for (int l = 0; l < 4; l++) acc[l] += wk*Lg[l];
            }
#pragma unroll
            for (int l = 0; l < 4; l++) total[l] += acc[l]*hw;
        }
    }
#pragma unroll
    for (int l = 0; l < 4; l++)
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) total[l] += __shfl_xor_sync(0xffffffffu, total[l], sh);
    if (lane == 0) {
        double2* dst = reinterpret_cast<double2*>(part + ((size_t)o*C.ncell + gc)*4);
        dst[0] = make_double2(total[0], total[1]);
        dst[1] = make_double2(total[2], total[3]);
    }
}

__global__ void __launch_bounds__(256)
k_linear_gather(int o_first, const LinOutDev* __restrict__ outs, ZonesDev Z, CellsDev C, const double* __restrict__ part,
                double* __restrict__ Omega)
{
    const int o = blockIdx.y;
    const int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= TAB_MPAD) return;
    double total = 0.0;
    if (j < TAB_M) {
        int z = 0;
        while (z < Z.nz - 1 && j >= Z.start[z + 1]) ++z;
        const int jl = j - Z.start[z];
        const int n = Z.n[z];
        const int c_lo = jl - 3 > 0 ? jl - 3 : 0;
        const int c_hi = jl + 3 < n - 1 ? jl + 3 : n - 1;
        for (int c = c_lo; c <= c_hi; c++) {
            int s0 = c - 1;
            if (s0 < 0) s0 = 0;
            if (s0 > n - 3) s0 = n - 3;
            const int l = jl - s0;
            if (l < 0 || l > 3) continue;
            total += part[((size_t)o*C.ncell + C.cstart[z] + c)*4 + l];
        }
        total *= outs[o_first + o].scale;
    }
    Omega[(size_t)(o_first + o)*TAB_MPAD + j] = total;
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
    CellsDev C;
    C.ncell = 0;
    for (int i = 0; i < Z.nz; i++) { C.cstart[i] = C.ncell; C.ncell += Z.n[i]; }
    // the cell integrals of <= 2048 outputs at a time (160 MB of scratch)
    const int o_batch = std::min(nout, 2048);
    DevBuf<double> part;
    part.alloc((size_t)o_batch*C.ncell*4);
    for (int o0 = 0; o0 < nout; o0 += o_batch) {
        const int no = std::min(o_batch, nout - o0);
        k_linear_cells<<<dim3((C.ncell + 7)/8, no), 256, 0, ctx.stream>>>(o0, dout.p, Z, C, part.p);
        PRSD_CUDA(cudaGetLastError());
        k_linear_gather<<<dim3((TAB_MPAD + 255)/256, no), 256, 0, ctx.stream>>>(o0, dout.p, Z, C, part.p, Omega.p);
        PRSD_CUDA(cudaGetLastError());
        ctx.launches += 2;
    }
    // sync = false: the caller overlaps host work with the build; `dout` and `part` return to the stream-ordered pool,
    // which hands them out again only behind these kernels
    if (sync) ctx.sync();
}

// ---------------------------------------------------------------- apply
// out[row][o] = sum_j T[row][j] Omega[o][j] is a (rows x knots) . (knots x outputs) product: FP64 tensor-core MMAs
// (DMMA m8n8k4) in a shared-memory tiled GEMM.  A CTA owns 16 MT rows x 64 outputs (MT = 4, 2 or 1 by batch size), its 8
// warps (2 x 4) own 8 MT x 16 each (MT x 2 MMA tiles); the knots are walked in chunks of 32 through a three-stage cp.async
// pipeline (16-byte copies, zero-filled past the end of a row).  With MT = 4 each operand element is read from L2 once per
// 64 x 64 tile -- 8 flop per byte against 2.7 for round 1's kernel (32 x 16 outputs per CTA, fragments straight from
// global memory), whose 7.6 TB/s of L2 -> SM traffic was its limit -- and every fragment read from shared memory feeds
// 2 - 4 MMAs.  Row stride 36 doubles (= 4 mod 16): the 16 fragment reads of a half-warp (4 rows x 4 knots) fall on distinct
// bank pairs.  Every output element accumulates its knots in the same order whatever MT: results do not depend on the
// batch size, bit for bit.
static const int GM_BN = 64, GM_BK = 32, GM_LD = GM_BK + 4, GM_STAGES = 3;

template <int MT>
__global__ void __launch_bounds__(256, 2)
k_linear_gemm(int nout, int nrow, const double* __restrict__ Omega, const double* __restrict__ tables,
              const int* __restrict__ tab_index, double* __restrict__ out)
{
    constexpr int BM = 16*MT;
    extern __shared__ __align__(16) double gm_smem[];       // [STAGES][(BM + BN)][LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane & 3, grp = lane >> 2;
    const int o0 = blockIdx.x*GM_BN, r0 = blockIdx.y*BM;
    const int wm = warp >> 2, wn = warp & 3;                // warp tile: rows 8 MT wm .., outputs 16 wn ..

    // global -> shared: 16 chunks of 16 bytes per row and stage; thread t copies chunk t % 16 of rows t / 16 + 16 i
    const int c_chunk = tid & 15, c_row = tid >> 4;
    const double* src_a[MT];
    const double* src_b[4];
#pragma unroll
    for (int i = 0; i < MT; i++)
        src_a[i] = tables + (size_t)tab_index[min(r0 + c_row + 16*i, nrow - 1)]*TAB_MPAD + 2*c_chunk;
#pragma unroll
    for (int i = 0; i < 4; i++) src_b[i] = Omega + (size_t)min(o0 + c_row + 16*i, nout - 1)*TAB_MPAD + 2*c_chunk;
    auto load_stage = [&](int stage, int k0) {
        double* As = gm_smem + (size_t)stage*(BM + GM_BN)*GM_LD;
        double* Bs = As + BM*GM_LD;
        const int bytes = (k0 + 2*c_chunk < TAB_MPAD) ? 16 : 0;       // TAB_MPAD is even: a chunk is inside a row or outside
#pragma unroll
        for (int i = 0; i < MT; i++) cp_async16(As + (c_row + 16*i)*GM_LD + 2*c_chunk, bytes ? src_a[i] + k0 : tables, bytes);
#pragma unroll
        for (int i = 0; i < 4; i++) cp_async16(Bs + (c_row + 16*i)*GM_LD + 2*c_chunk, bytes ? src_b[i] + k0 : Omega, bytes);
    };
    constexpr int NCH = (TAB_MPAD + GM_BK - 1)/GM_BK;
#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; s++) {
        if (s < NCH) load_stage(s, s*GM_BK);
        cp_async_commit();
    }
    double acc[MT][2][2];
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
        for (int n = 0; n < 2; n++) acc[m][n][0] = acc[m][n][1] = 0.0;

    for (int ch = 0; ch < NCH; ch++) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();                                     // chunk ch has landed; the stage refilled below was read in ch - 1
        const int nx = ch + GM_STAGES - 1;
        if (nx < NCH) load_stage(nx % GM_STAGES, nx*GM_BK);
        cp_async_commit();
        const double* As = gm_smem + (size_t)(ch % GM_STAGES)*(BM + GM_BN)*GM_LD + (8*MT*wm + grp)*GM_LD + sub;
        const double* Bs = gm_smem + (size_t)(ch % GM_STAGES)*(BM + GM_BN)*GM_LD + (BM + 16*wn + grp)*GM_LD + sub;
#pragma unroll
        for (int kk = 0; kk < GM_BK; kk += 4) {
            double a[MT], b[2];
#pragma unroll
            for (int m = 0; m < MT; m++) a[m] = As[m*8*GM_LD + kk];
#pragma unroll
            for (int n = 0; n < 2; n++) b[n] = Bs[n*8*GM_LD + kk];
#pragma unroll
            for (int m = 0; m < MT; m++)
#pragma unroll
                for (int n = 0; n < 2; n++) dmma_m8n8k4(acc[m][n][0], acc[m][n][1], a[m], b[n]);
        }
    }
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
        for (int n = 0; n < 2; n++) {
            const int row = r0 + 8*MT*wm + 8*m + grp, col = o0 + 16*wn + 8*n + 2*sub;
            if (row < nrow) {
                if (col < nout) out[(size_t)row*nout + col] = acc[m][n][0];
                if (col + 1 < nout) out[(size_t)row*nout + col + 1] = acc[m][n][1];
            }
        }
}

template <int MT>
static void launch_gemm(Context& ctx, int nout, int nrow, const double* Omega, const double* tables, const int* tab_index, double* out)
{
    const size_t smem = (size_t)GM_STAGES*(16*MT + GM_BN)*GM_LD*sizeof(double);
    PRSD_CUDA(cudaFuncSetAttribute(k_linear_gemm<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((nout + GM_BN - 1)/GM_BN, (nrow + 16*MT - 1)/(16*MT));
    k_linear_gemm<MT><<<grid, 256, smem, ctx.stream>>>(nout, nrow, Omega, tables, tab_index, out);
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
    ProfScope prof(ctx, PROF_LINEAR);
    if (nrow > 32) launch_gemm<4>(ctx, nout, nrow, Omega.p, tables, tab_index, out);
    else if (nrow > 16) launch_gemm<2>(ctx, nout, nrow, Omega.p, tables, tab_index, out);
    else launch_gemm<1>(ctx, nout, nrow, Omega.p, tables, tab_index, out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
