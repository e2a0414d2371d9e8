// nodal.cu -- one-lane-per-node evaluation of the ImnOneLoop and one-loop-table operators (see nodal.cuh).
#include "nodal.cuh"
#include "pt_math.cuh"
#include "ptx.cuh"

#include <algorithm>

namespace prsd {

static const int NL_THREADS = 256;
static const int NL_WARPS = NL_THREADS/32;
static const int NL_SLOTS = 4;
static const double NL_INV8PI2 = 1.0/(8.0*M_PI*M_PI);

struct NodalArgs {
    int nk, max_outer, kchunk, nrow;
    const int64_t* outer_off;
    const int64_t* node_off;
    const int* a_j0;
    const double* a_s;      // stencil coordinate of every outer node
    const int2* b_node;     // .x = (outer index << 12) | stencil start, .y = stencil coordinate (binary32 bits): the
                            // Lagrange coefficients are lagrange4(s)
    const float4* W;        // [ceil(NW / 4)][n_nodes], binary32, four columns per load (bilinear.cuh)
    int64_t n_nodes;
    const double* tables;
    const int* slot_tab;    // [ntiles][4]
    double* out;
};

// ------------------------------------------------------------------ ImnOneLoop
// slots: 0 P_L, 1 P_dd, 2 P_dv, 3 P_vv.  W columns (ml_columns()):
//   0..6   f(+v)  of h01 h02 h03 h04 f23 f32 f33               -> P_1(a) P_2(b)
//   7, 8   f(-v)  of h01 h02                                   -> P_1(b) P_2(a)
// h03, h04, f23, f32, f33 depend on u^2, v^2 and (u^2 - v^2) only: their f(-v) column IS the f(+v) one, so the
// two halves share it (9 columns streamed instead of 14).
// accumulators:
//   0..5 (L,L): h01 h02 h03 h04 f23 f33 | 6,7 cross of ImnOneLoop(P_vv, P_dd): h01 h02 | 8,9 its one-loop part
//   10,11 (L,dv): h03 h04 | 12,13 (dv,dv) | 14..16 (L,vv): f23 f32 f33 | 17,18 (vv,vv): f23 f33
struct SpecML {
    static const int NSLOT = 4, MINB = 2, NACC = 19, NW = 9, NOUT = 21, THREADS = NL_THREADS;
    static const bool PREFETCH_W = false;    // W one iteration ahead: no gain here (1.453 vs 1.456 ms, 128 registers)
    static const bool DOUBLE_BUF = true;     // P(a) and the per-warp partial sums ping-pong between two buffers: one barrier per k
    __device__ __forceinline__ static void accumulate(double* acc, const double* pa, const double* pb, const double* w)
    {
        const double La = pa[0], Da = pa[1], Va = pa[2], Wa = pa[3];
        const double Lb = pb[0], Db = pb[1], Vb = pb[2], Wb = pb[3];
        const double h01p = w[0], h02p = w[1], h01n = w[7], h02n = w[8];
        // (L, L): both halves of every kernel
        const double ll = La*Lb, ll2 = ll + ll;
        acc[0] = fma(h01p + h01n, ll, acc[0]);
        acc[1] = fma(h02p + h02n, ll, acc[1]);
        acc[2] = fma(w[2], ll2, acc[2]);
        acc[3] = fma(w[3], ll2, acc[3]);
        acc[4] = fma(w[4], ll2, acc[4]);
        acc[5] = fma(w[6], ll2, acc[5]);
        // ImnOneLoop(P_vv, P_dd): cross = (P_L, P_dd) + (P_vv, P_L) (ImnOneLoop.cpp:130-148)
        const double cp = fma(La, Db, Wa*Lb), cn = fma(Lb, Da, Wb*La);
        acc[6] = fma(h01p, cp, fma(h01n, cn, acc[6]));
        acc[7] = fma(h02p, cp, fma(h02n, cn, acc[7]));
        const double op = Wa*Db, on = Wb*Da;
        acc[8] = fma(h01p, op, fma(h01n, on, acc[8]));
        acc[9] = fma(h02p, op, fma(h02n, on, acc[9]));
        // ImnOneLoop(P_dv): h03, h04
        const double dpn = fma(La, Vb, Lb*Va);
        acc[10] = fma(w[2], dpn, acc[10]);
        acc[11] = fma(w[3], dpn, acc[11]);
        const double dd2 = 2.0*(Va*Vb);
        acc[12] = fma(w[2], dd2, acc[12]);
        acc[13] = fma(w[3], dd2, acc[13]);
        // ImnOneLoop(P_vv): f23, f32, f33
        const double vpn = fma(La, Wb, Lb*Wa);
        acc[14] = fma(w[4], vpn, acc[14]);
        acc[15] = fma(w[5], vpn, acc[15]);
        acc[16] = fma(w[6], vpn, acc[16]);
        const double vv2 = 2.0*(Wa*Wb);
        acc[17] = fma(w[4], vv2, acc[17]);
        acc[18] = fma(w[6], vv2, acc[18]);
    }
    // out [c][7][3][nk]; t = output index 0..20 = 3 m + part; every output is one accumulator (x 2 for the cross
    // parts of the equal-spectra drivers).  The reference evaluates the LINEAR and ONE-LOOP parts of (m, n) = (3, 2)
    // with the f23 kernel (ImnOneLoop.cpp:108-109, 199-200) and only the cross part with f32 (:160-165); as is.
    __device__ __forceinline__ static int source(int t, double& fac)
    {
        const int src[21] = {0, 6, 8,  1, 7, 9,  2, 10, 12,  3, 11, 13,  4, 14, 17,  4, 15, 17,  5, 16, 18};
        fac = (t >= 6 && t % 3 == 1) ? 2.0 : 1.0;
        return src[t];
    }
    __device__ __forceinline__ static void store(const NodalArgs& g, int tile, int ik, int t, double v)
    {
        g.out[((size_t)tile*21 + t)*g.nk + ik] = v;
    }
};

std::vector<ColSpec> ml_columns()
{
    std::vector<ColSpec> c;
    for (int id : {KID_H01, KID_H02, KID_K20A, KID_K20B, KID_F23, KID_F32, KID_F33}) c.push_back({id, +1, NL_INV8PI2});
    for (int id : {KID_H01, KID_H02}) c.push_back({id, -1, NL_INV8PI2});
    return c;
}

// ------------------------------------------------------------ one-loop tables
// slots: P_L of DIAG_R cosmologies; W columns: the operator's 4 kernels; accumulator 4 r + c
static const int DIAG_R = 4;      // measured at batch 256: R = 4 (2 CTAs per SM) 1.65 ms, R = 3 (3 CTAs) +9 %, R = 2 (3 CTAs) +23 %,
                                  // R = 8 (one CTA of 512 threads per SM, 128 registers) +23 %
struct SpecDiag {
    static const int NC = 4;      // I00, I01, I11 and P22bar = I23 + 2/3 I32 + 1/5 I33 (combined in W at build time)
    static const int NSLOT = DIAG_R, MINB = DIAG_R <= 3 ? 3 : (DIAG_R <= 4 ? 2 : 1), NACC = NC*DIAG_R, NW = NC, NOUT = NC*DIAG_R;
    static const int THREADS = DIAG_R <= 4 ? NL_THREADS : 2*NL_THREADS;
    static const bool PREFETCH_W = true;     // the float4 of W one iteration ahead: 0.935 -> 0.907 ms (one register quad)
    static const bool DOUBLE_BUF = false;    // 2 x 4 x 752 outer nodes would not leave room for two CTAs per SM
    __device__ __forceinline__ static void accumulate(double* acc, const double* pa, const double* pb, const double* w)
    {
#pragma unroll
        for (int r = 0; r < DIAG_R; r++) {
            const double pp = pa[r]*pb[r];
#pragma unroll
            for (int c = 0; c < NC; c++) acc[NC*r + c] = fma(w[c], pp, acc[NC*r + c]);
        }
    }
    __device__ __forceinline__ static int source(int t, double& fac) { fac = 1.0; return t; }
    // out [cosmology][nk][8]
    __device__ __forceinline__ static void store(const NodalArgs& g, int tile, int ik, int t, double v)
    {
        const int r = t/NC, c = t - NC*r;
        const int cosmo = tile*DIAG_R + r;
        if (cosmo < g.nrow) g.out[((size_t)cosmo*g.nk + ik)*8 + c] = v;
    }
};

// Warp-wide sums of N per-lane accumulators, transposing as they go: at the step with lane mask M the lanes with the
// bit clear keep the lower half of their array and receive the other lanes' lower half, the lanes with the bit set
// the upper half -- N/2 + N/4 + ... + 1 ~ N 64-bit shuffles and additions in all (19 accumulators: 21) against 5 N for
// a butterfly per accumulator and N tensor MMAs + 2 N shuffles + 3 N additions for the MMA-assisted form used before
// (an FP64 MMA m8n8k4 holds the FP64 pipe as long as four DFMA).  On return a[0] of lane l is the total of
// accumulator `idx` if `valid` (odd sizes are padded with zeros, whose sums land on the invalid lanes).
template <int N, int MASK>
struct WarpTransposeSum {
    __device__ __forceinline__ static void run(double* a, int lane, int& idx, int& nv)
    {
        constexpr int H = (N + 1)/2;
        const bool up = (lane & MASK) != 0;
#pragma unroll
        for (int i = 0; i < H; i++) {
            const double lo = a[i], hi = (i + H < N) ? a[i + H] : 0.0;
            const double send = up ? lo : hi, keep = up ? hi : lo;
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
        }
        idx += up ? H : 0;
        nv = up ? max(nv - H, 0) : min(nv, H);
        WarpTransposeSum<H, MASK/2>::run(a, lane, idx, nv);
    }
};
template <int N>
struct WarpTransposeSum<N, 0> {
    __device__ __forceinline__ static void run(double*, int, int&, int&) { static_assert(N == 1, "more than 32 accumulators"); }
};
template <int MASK>
struct WarpTransposeSum<1, MASK> {
    // one value left and lane bits to spare: plain exchange, both lanes end with the sum (the lower one reports it)
    __device__ __forceinline__ static void run(double* a, int lane, int& idx, int& nv)
    {
        a[0] += __shfl_xor_sync(0xffffffffu, a[0], MASK);
        if (lane & MASK) nv = 0;
        WarpTransposeSum<1, MASK/2>::run(a, lane, idx, nv);
    }
};
template <>
struct WarpTransposeSum<1, 0> {
    __device__ __forceinline__ static void run(double*, int, int&, int&) {}
};

// (Tried and rejected, round 2: streaming the Lagrange coefficients c_1..c_3 as binary32 in a 16-byte node word instead of
// rebuilding them from the stencil coordinate -- 13 FP64 operations fewer per node, 8 bytes more: ImnOneLoop unchanged,
// one-loop tables 2.6 % slower.  Reading three knots instead of four (not parity-clean, timing experiment only): -5.7 % /
// -1.2 %.  Neither the FP64 pipe nor the shared-memory pipe alone sets the time of these kernels.)
// (P(a) staged as binary32 -- half the wavefronts of the four P(a) reads, room for the second buffer of the one-loop
// operator: ImnOneLoop 1.006 -> 1.025 ms, one-loop 0.776 -> 0.765 ms, parity unchanged, exact amplitude scaling lost.  Rejected.)
// (Tried and rejected this round: staging the tables in binary32.  Rounding the TABLE is harmless -- the error of
// a knot is common to every node that reads it and cancels wherever the kernels cancel -- and halves the table
// wavefronts, but the kernel then runs into the P(a) reads and the global stream and gains only ~10 %; with the
// interpolation ARITHMETIC in binary32 the h01-type cancellations amplify the independent 6e-8 errors to 1e-4.)
template <class Spec>
__global__ void __launch_bounds__(Spec::THREADS, Spec::MINB)
k_nodal_apply(const NodalArgs g)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NBUF = Spec::DOUBLE_BUF ? 2 : 1;
    constexpr int NWARP = Spec::THREADS/32;
    double* tabs = reinterpret_cast<double*>(smem_raw);                 // [4][TAB_MPAD]
    double* PaS = tabs + Spec::NSLOT*TAB_MPAD;                             // [NBUF][4][max_outer]
    double* red = PaS + NBUF*Spec::NSLOT*g.max_outer;                      // [NBUF][NWARP][NACC]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + NBUF*NWARP*Spec::NACC + ((NBUF*NWARP*Spec::NACC) & 1));

    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // k is dealt round-robin over blockIdx.y: the node count grows ~10x from the lowest to the highest k
    // (quad.h tiers), contiguous chunks would leave a long tail
    const int k_stride = gridDim.y;

    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = TAB_MPAD*sizeof(double);
        mbar_expect_tx(mbar, Spec::NSLOT*bytes);
        for (int s = 0; s < Spec::NSLOT; s++)
            bulk_copy_g2s(tabs + s*TAB_MPAD, g.tables + (size_t)g.slot_tab[tile*Spec::NSLOT + s]*TAB_MPAD, bytes, mbar);
    }

    // the offsets of the NEXT k of this CTA are fetched during the current one, and the first node of a k is
    // requested before the P(a) pass: the k loop starts without a chain of exposed L2 round trips.  The same for the
    // first outer node of this lane (stencil start and coordinate): the P(a) pass of the next k starts from registers.
    int ik = blockIdx.y;
    int64_t o0_next = 0, o1_next = 0, n0_next = 0, n1_next = 0;
    int aj_next = 0;
    double as_next = 0.0;
    if (ik < g.nk) {
        o0_next = g.outer_off[ik]; o1_next = g.outer_off[ik + 1];
        n0_next = g.node_off[ik];  n1_next = g.node_off[ik + 1];
        if (tid < (int)(o1_next - o0_next)) { aj_next = g.a_j0[o0_next + tid]; as_next = g.a_s[o0_next + tid]; }
    }
    mbar_wait(mbar, 0);

    // finish k: the outputs are formed straight from the per-warp partial sums
    auto finish = [&](int k_done, const double* rbuf) {
        if (tid < Spec::NOUT) {
            double fac;
            const int src = Spec::source(tid, fac);
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; w++) s += rbuf[w*Spec::NACC + src];
            Spec::store(g, tile, k_done, tid, fac*s);
        }
    };
    int par = 0, ik_prev = -1;
    for (; ik < g.nk; ik += k_stride, par ^= (NBUF - 1)) {
        const int64_t o0 = o0_next, n0 = n0_next, n1 = n1_next;
        const int n_out = (int)(o1_next - o0);
        const int aj_first = aj_next;
        const double as_first = as_next;
        double* Pa = PaS + par*Spec::NSLOT*g.max_outer;
        double* rd = red + par*NWARP*Spec::NACC;
        // the index word and the stencil coordinate of the NEXT node of this lane are loaded one iteration ahead: the
        // shared-memory reads depend on them, and an exposed L2 round trip per node was the largest stall
        int64_t node = n0 + tid;
        int2 nw_next = make_int2(0, 0);
        constexpr int NV = (Spec::NW + 3)/4;
        float4 w_next[Spec::PREFETCH_W ? NV : 1];
        if (node < n1) {
            nw_next = __ldg(g.b_node + node);
            if (Spec::PREFETCH_W) {
#pragma unroll
                for (int v = 0; v < NV; v++) w_next[v] = __ldg(g.W + (int64_t)v*g.n_nodes + node);
            }
        }
        if (ik + k_stride < g.nk) {
            o0_next = g.outer_off[ik + k_stride]; o1_next = g.outer_off[ik + k_stride + 1];
            n0_next = g.node_off[ik + k_stride];  n1_next = g.node_off[ik + k_stride + 1];
        }
        // ---- P(a) of the four spectra for every outer node of this k
        for (int o = tid; o < n_out; o += Spec::THREADS) {
            const int j0 = (o == tid) ? aj_first : g.a_j0[o0 + o];
            double c[4];
            lagrange4((o == tid) ? as_first : __ldg(g.a_s + o0 + o), c);
#pragma unroll
            for (int s = 0; s < Spec::NSLOT; s++) {
                const double* t = tabs + s*TAB_MPAD + j0;
                Pa[s*g.max_outer + o] = fma(c[3], t[3], fma(c[2], t[2], fma(c[1], t[1], c[0]*t[0])));
            }
        }
        if (ik + k_stride < g.nk && tid < (int)(o1_next - o0_next)) {
            aj_next = g.a_j0[o0_next + tid];
            as_next = g.a_s[o0_next + tid];
        }
        __syncthreads();
        // with two buffers the barrier above is the only one of a k: it also orders every warp's partial sums of the
        // previous k before they are read here, and this k's P(a) / partial sums go to the other buffer
        if (Spec::DOUBLE_BUF && ik_prev >= 0) finish(ik_prev, red + (par ^ 1)*NWARP*Spec::NACC);

        double acc[Spec::NACC];
#pragma unroll
        for (int i = 0; i < Spec::NACC; i++) acc[i] = 0.0;
        for (; node < n1; node += Spec::THREADS) {
            const int pk = nw_next.x;
            double c[4];
            lagrange4((double)__int_as_float(nw_next.y), c);
            float4 wv[NV];
            if (Spec::PREFETCH_W) {
#pragma unroll
                for (int v = 0; v < NV; v++) wv[v] = w_next[v];
            }
            const int64_t nx = node + Spec::THREADS;
            if (nx < n1) {
                nw_next = __ldg(g.b_node + nx);
                if (Spec::PREFETCH_W) {
#pragma unroll
                    for (int v = 0; v < NV; v++) w_next[v] = __ldg(g.W + (int64_t)v*g.n_nodes + nx);
                }
            }
            const int j0 = pk & 4095, ol = pk >> 12;
            if (!Spec::PREFETCH_W) {
#pragma unroll
                for (int v = 0; v < NV; v++) wv[v] = __ldg(g.W + (int64_t)v*g.n_nodes + node);
            }
            double w[4*NV];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                w[4*v] = (double)wv[v].x; w[4*v + 1] = (double)wv[v].y; w[4*v + 2] = (double)wv[v].z; w[4*v + 3] = (double)wv[v].w;
            }
            double pa[Spec::NSLOT], pb[Spec::NSLOT];
#pragma unroll
            for (int s = 0; s < Spec::NSLOT; s++) {
                const double* t = tabs + s*TAB_MPAD + j0;
                pb[s] = fma(c[3], t[3], fma(c[2], t[2], fma(c[1], t[1], c[0]*t[0])));
                pa[s] = Pa[s*g.max_outer + ol];
            }
            Spec::accumulate(acc, pa, pb, w);
        }
        // ---- reduce over the CTA: transposing warp sums (WarpTransposeSum), then one partial per warp in shared memory
        {
            int idx = 0, nv = Spec::NACC;
            WarpTransposeSum<Spec::NACC, 16>::run(acc, lane, idx, nv);
            if (nv > 0) rd[warp*Spec::NACC + idx] = acc[0];
        }
        if (!Spec::DOUBLE_BUF) {
            __syncthreads();
            finish(ik, rd);
            // the next k's writes of PaS and red are ordered after this k's reads by the barrier after the next P(a) pass
        }
        ik_prev = ik;
    }
    if (Spec::DOUBLE_BUF && ik_prev >= 0) {
        __syncthreads();
        finish(ik_prev, red + (par ^ 1)*NWARP*Spec::NACC);
    }
}

template <class Spec>
static void launch_nodal(Context& ctx, const BilinearOp& op, const double* tables, const int* slot_tab, int ntiles,
                         int nrow, int kchunk, double* out, int prof_tag)
{
    PRSD_REQUIRE(op.soa && op.ncol_used == Spec::NW && op.Wf.p, "nodal operator has the wrong layout");
    const NodeGeometry& geo = *op.geo;
    NodalArgs a;
    a.nk = geo.nk; a.max_outer = geo.max_outer; a.kchunk = kchunk; a.nrow = nrow;
    a.outer_off = geo.outer_off.p; a.node_off = geo.node_off.p;
    a.a_j0 = geo.a_j0.p; a.a_s = geo.a_s.p;
    a.b_node = geo.b_node.p;
    a.W = op.Wf.p; a.n_nodes = geo.n_nodes;
    a.tables = tables; a.slot_tab = slot_tab; a.out = out;
    const int nbuf = Spec::DOUBLE_BUF ? 2 : 1;
    const size_t smem = (size_t)(Spec::NSLOT*TAB_MPAD + nbuf*Spec::NSLOT*geo.max_outer + nbuf*(Spec::THREADS/32)*Spec::NACC + 2)*sizeof(double) + 16;
    PRSD_REQUIRE(smem <= ctx.smem_optin, "nodal apply needs %zu B shared memory (> %zu)", smem, ctx.smem_optin);
    PRSD_CUDA(cudaFuncSetAttribute(k_nodal_apply<Spec>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // k chunks: few enough to amortise the table staging, enough CTAs to fill 148 SMs x 2 when the batch is small
    int nchunks = (geo.nk + kchunk - 1)/kchunk;
    nchunks = std::min(geo.nk, std::max(nchunks, (2400 + ntiles - 1)/ntiles));
    dim3 grid(ntiles, nchunks);
    ProfScope prof(ctx, prof_tag);
    k_nodal_apply<Spec><<<grid, Spec::THREADS, smem, ctx.stream>>>(a);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

void nodal_apply_ml(Context& ctx, const BilinearOp& op, const double* tables, const int* slot_tab, int ncosmo, double* out)
{
    launch_nodal<SpecML>(ctx, op, tables, slot_tab, ncosmo, ncosmo, 8, out, PROF_NODAL_ML);
}

int nodal_diag_rows() { return DIAG_R; }

void nodal_apply_diag(Context& ctx, const BilinearOp& op, const double* tables, const int* slot_tab, int ncosmo, double* out)
{
    launch_nodal<SpecDiag>(ctx, op, tables, slot_tab, (ncosmo + DIAG_R - 1)/DIAG_R, ncosmo, 8, out, PROF_NODAL_DIAG);
}

} // namespace prsd
