// bilinear.cu -- build and apply the mode-coupling quadrature operators (see bilinear.cuh).
#include "bilinear.cuh"
#include "pt_math.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>

namespace prsd {

size_t NodeGeometry::bytes() const
{
    // inner nodes: b, wgt, n_outer_local, b_node; outer nodes: a, a_j0, a_coef, a_s
    return n_nodes*(sizeof(double)*2 + sizeof(int) + sizeof(int2)) + n_outer*(sizeof(int) + sizeof(double)*6);
}

// Host part of the geometry for a contiguous range of k: nodes, padding and lane ordering.  Every k is independent,
// so make_geometry runs this on several threads over chunks of the k list and concatenates the pieces.
namespace {
struct GeoChunk {
    HalfNodes H;                                  // outer-node arrays in their final (reordered) order
    std::vector<int> b_j0, n_ol;
    std::vector<double> b, wgt, b_coef;
    std::vector<int64_t> node_off;                // padded node offsets, local, size nk + 1
    int max_outer = 0;
};

void geometry_chunk(const double* k, int nk, double qmax, const QuadConfig& cfg, const KnotGrid& grid, int BM, GeoChunk& out)
{
    HalfNodes& H = out.H;
    build_half_nodes(k, nk, qmax, cfg, grid, H);

    // Outer nodes of one k are tabulated P(a) = sum_m coef_m t[j0 + m] by consecutive lanes: order them so that every
    // run of 16 has stencil starts distinct modulo 16 where the list allows (same wavefront argument as below).
    for (int ik = 0; ik < nk; ik++) {
        const int64_t o0 = H.outer_off[ik], no = H.outer_off[ik + 1] - o0;
        std::vector<std::vector<int>> q(16);
        for (int64_t o = no - 1; o >= 0; o--) q[H.a_j0[o0 + o] & 15].push_back((int)o);     // pop_back = ascending
        std::vector<int> perm, inv(no);
        perm.reserve(no);
        while ((int64_t)perm.size() < no) {
            int order[16];
            for (int r = 0; r < 16; r++) order[r] = r;
            std::sort(order, order + 16, [&](int x, int y) { return q[x].size() > q[y].size(); });
            int room = (int)std::min<int64_t>(16, no - (int64_t)perm.size());
            for (int i = 0; i < 16 && room > 0; i++)
                if (!q[order[i]].empty()) { perm.push_back(q[order[i]].back()); q[order[i]].pop_back(); room--; }
            for (int i = 0; room > 0; i = (i + 1) & 15)
                if (!q[order[i]].empty()) { perm.push_back(q[order[i]].back()); q[order[i]].pop_back(); room--; }
        }
        std::vector<double> a2(no), c2(4*no);
        std::vector<int> j2(no);
        for (int64_t n = 0; n < no; n++) {
            const int64_t o = perm[n];
            inv[o] = (int)n;
            a2[n] = H.a[o0 + o];
            j2[n] = H.a_j0[o0 + o];
            for (int m = 0; m < 4; m++) c2[4*n + m] = H.a_coef[4*(o0 + o) + m];
        }
        std::copy(a2.begin(), a2.end(), H.a.begin() + o0);
        std::copy(j2.begin(), j2.end(), H.a_j0.begin() + o0);
        std::copy(c2.begin(), c2.end(), H.a_coef.begin() + 4*o0);
        for (int64_t j = H.node_off[ik]; j < H.node_off[ik + 1]; j++) H.n_outer[j] = o0 + inv[H.n_outer[j] - o0];
    }

    // Per k: pad the node list to a multiple of 4 (one DMMA k-step) with weightless nodes, then order it for the
    // shared-memory reads of the nodal kernels: a half-warp (16 consecutive nodes) reads t[j0 + m] with 8-byte
    // accesses, which is one wavefront only if the 16 stencil starts are equal or distinct modulo 16.  Nodes
    // sharing a stencil start are kept together (broadcast) and every half-warp is filled from distinct residues.
    std::vector<int>& b_j0 = out.b_j0;
    std::vector<int>& n_ol = out.n_ol;
    std::vector<double>& b = out.b;
    std::vector<double>& wgt = out.wgt;
    std::vector<double>& b_coef = out.b_coef;
    out.node_off.assign(1, 0);
    {
        // room for every padded node up front: the gather below appends ~1e6 times per operator
        const int64_t pad_to = BM >= 16 ? 16 : 4;
        size_t total = 0;
        for (int ik = 0; ik < nk; ik++) total += (size_t)((H.node_off[ik + 1] - H.node_off[ik] + pad_to - 1)/pad_to*pad_to);
        b.reserve(total); wgt.reserve(total); b_j0.reserve(total); n_ol.reserve(total); b_coef.reserve(4*total);
    }
    std::vector<int64_t> src;                     // source node of every slot of this k (-1: padding)
    // an item = the nodes of one k that share the stencil start AND the outer node: every read of theirs is a
    // broadcast.  A group of BM lanes takes items whose stencil starts are distinct modulo BM (table reads) and,
    // where possible, whose outer indices are too (the P(a) reads of the nodal kernels).
    struct Item { int j0, ol; int64_t begin, end; };
    std::vector<int64_t> sorted, sort_tmp;
    std::vector<int64_t> sort_cnt;
    std::vector<Item> items;
    std::vector<std::vector<int>> queue(BM);      // per stencil residue: item indices, ascending
    for (int ik = 0; ik < nk; ik++) {
        const int64_t n0 = H.node_off[ik], n1 = H.node_off[ik + 1];
        const int64_t o0 = H.outer_off[ik];
        out.max_outer = std::max<int>(out.max_outer, (int)(H.outer_off[ik + 1] - o0));
        // nodal kernels (BM = 16) read the node stream 32 nodes per warp: segments that start on a 128-byte
        // boundary cost two cache lines per load instead of three
        const int64_t pad_to = BM >= 16 ? 16 : 4;
        const int64_t cnt = n1 - n0, padded = (cnt + pad_to - 1) / pad_to * pad_to;
        auto j0_of = [&](int64_t j) { return j < 0 ? 0 : H.b_j0[j]; };
        auto ol_of = [&](int64_t j) { return j < 0 ? 0 : (int)(H.n_outer[j] - o0); };
        sorted.clear();
        for (int64_t j = n0; j < n1; j++) sorted.push_back(j);
        for (int64_t j = cnt; j < padded; j++) sorted.push_back(-1);       // padding reads stencil 0, outer node 0
        // stable order by (stencil start, outer node): two stable counting passes, least significant key first (the
        // comparison sort this replaces took a third of the generator's time; same order, geometry_checksum unchanged)
        {
            const int n_ol_keys = (int)(H.outer_off[ik + 1] - o0) + 1;
            int j0_max = 0;
            for (int64_t x : sorted) j0_max = std::max(j0_max, j0_of(x));
            sort_tmp.resize(sorted.size());
            sort_cnt.assign((size_t)n_ol_keys + 1, 0);
            for (int64_t x : sorted) sort_cnt[ol_of(x) + 1]++;
            for (int i = 0; i < n_ol_keys; i++) sort_cnt[i + 1] += sort_cnt[i];
            for (int64_t x : sorted) sort_tmp[sort_cnt[ol_of(x)]++] = x;
            sort_cnt.assign((size_t)j0_max + 2, 0);
            for (int64_t x : sort_tmp) sort_cnt[j0_of(x) + 1]++;
            for (int i = 0; i <= j0_max; i++) sort_cnt[i + 1] += sort_cnt[i];
            for (int64_t x : sort_tmp) sorted[sort_cnt[j0_of(x)]++] = x;
        }
        items.clear();
        for (auto& q : queue) q.clear();
        for (int64_t i = 0; i < padded; ) {
            int64_t e = i + 1;
            while (e < padded && j0_of(sorted[e]) == j0_of(sorted[i]) && ol_of(sorted[e]) == ol_of(sorted[i])) e++;
            queue[j0_of(sorted[i]) & (BM - 1)].push_back((int)items.size());
            items.push_back({j0_of(sorted[i]), ol_of(sorted[i]), i, e});
            i = e;
        }
        size_t head[32] = {0};
        int64_t remaining[32];
        for (int r = 0; r < BM; r++) {
            remaining[r] = 0;
            for (int it : queue[r]) remaining[r] += items[it].end - items[it].begin;
        }
        src.clear();
        int64_t left = padded;
        while (left > 0) {
            int room = (int)std::min<int64_t>(BM, left);
            int used_j0[32], used_ol[32];
            for (int r = 0; r < BM; r++) used_j0[r] = used_ol[r] = -1;
            int order[32];
            for (int r = 0; r < BM; r++) order[r] = r;
            std::sort(order, order + BM, [&](int x, int y) { return remaining[x] > remaining[y]; });
            // strict: 0 = both reads conflict-free, 1 = table reads conflict-free, 2 = anything
            for (int pass = 0; room > 0 && left > 0; pass++) {
                const int strict = std::min(pass, 2);
                for (int i = 0; i < BM && room > 0; i++) {
                    const int r = order[i];
                    auto& q = queue[r];
                    while (head[r] < q.size() && items[q[head[r]]].begin == items[q[head[r]]].end) head[r]++;
                    int pick = -1;
                    for (size_t c = head[r], seen = 0; c < q.size() && seen < 8; c++) {
                        const Item& it = items[q[c]];
                        if (it.begin == it.end) continue;
                        seen++;
                        const bool ok_j0 = used_j0[r] < 0 || used_j0[r] == it.j0;
                        const int ro = it.ol & (BM - 1);
                        const bool ok_ol = used_ol[ro] < 0 || used_ol[ro] == it.ol;
                        if ((strict == 0 && ok_j0 && ok_ol) || (strict == 1 && ok_j0) || strict == 2) { pick = q[c]; break; }
                    }
                    if (pick < 0) continue;
                    Item& it = items[pick];
                    used_j0[r] = it.j0;
                    used_ol[it.ol & (BM - 1)] = it.ol;
                    while (room > 0 && it.begin < it.end) {
                        src.push_back(sorted[it.begin++]);
                        room--; left--; remaining[r]--;
                    }
                }
            }
        }
        for (int64_t j : src) {
            if (j >= 0) {
                b.push_back(H.b[j]);
                wgt.push_back(H.wgt[j]);
                b_j0.push_back(H.b_j0[j]);
                n_ol.push_back((int)(H.n_outer[j] - o0));
                b_coef.insert(b_coef.end(), &H.b_coef[4*j], &H.b_coef[4*j] + 4);
            } else {
                b.push_back(k[ik]);
                wgt.push_back(0.0);
                b_j0.push_back(0);
                n_ol.push_back(0);
                b_coef.insert(b_coef.end(), 4, 0.0);
            }
        }
        out.node_off.push_back((int64_t)b.size());
    }
}
} // namespace

// Host part of make_geometry: everything up to the arrays the device receives (pure CPU, thread-safe: PTPlan builds the
// geometries of its three operators concurrently while the device integrates the 1-D operators).
struct GeometryHost {
    int nk = 0, max_outer = 0;
    std::vector<double> k;
    std::vector<int64_t> node_off, outer_off;
    HalfNodes H;                                   // merged outer-node arrays
    std::vector<int> b_j0, n_ol, pack;
    std::vector<double> b, wgt, b_coef, bs, as;
    std::vector<int2> nodew;
};

std::shared_ptr<GeometryHost> make_geometry_host(const double* k, int nk, double qmax, const QuadConfig& cfg,
                                                 const KnotGrid& grid, int bank_mod)
{
    PRSD_REQUIRE(bank_mod == 4 || bank_mod == 16 || bank_mod == 32, "make_geometry: bank_mod must be 4, 16 or 32");
    const int BM = bank_mod;
    PRSD_REQUIRE(nk > 0, "make_geometry: empty k list");
    for (int i = 0; i < nk; i++) PRSD_REQUIRE(k[i] > 0 && std::isfinite(k[i]), "make_geometry: k[%d] = %g must be > 0", i, k[i]);
    // chunks of k dealt round-robin in blocks to the threads (the node count grows ~10x from the lowest to the highest
    // k): 4 chunks per thread keep the threads busy to the end
    const int hw = (int)std::thread::hardware_concurrency();
    const int nthreads = std::max(1, std::min({hw > 0 ? hw : 4, 16, nk/4}));
    const int nchunk = nthreads == 1 ? 1 : std::min(nk, 4*nthreads);
    std::vector<GeoChunk> chunks(nchunk);
    std::vector<int> lo(nchunk + 1);
    for (int c = 0; c <= nchunk; c++) lo[c] = (int)((int64_t)nk*c/nchunk);
    // the node count per k grows with k: hand out the expensive chunks (high k) first
    std::atomic<int> next{0};
    std::string failure;
    std::mutex fmu;
    auto worker = [&]() {
        for (int t = next++; t < nchunk; t = next++) {
            const int c = nchunk - 1 - t;
            try {
                geometry_chunk(k + lo[c], lo[c + 1] - lo[c], qmax, cfg, grid, BM, chunks[c]);
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lk(fmu);
                failure = e.what();
            }
        }
    };
    if (nthreads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    PRSD_REQUIRE(failure.empty(), "make_geometry: %s", failure.c_str());

    auto g = std::make_shared<GeometryHost>();
    g->nk = nk;
    g->k.assign(k, k + nk);
    HalfNodes& H = g->H;
    std::vector<int>& b_j0 = g->b_j0;
    std::vector<int>& n_ol = g->n_ol;
    std::vector<double>& b = g->b;
    std::vector<double>& wgt = g->wgt;
    std::vector<double>& b_coef = g->b_coef;
    // concatenate the chunks: offsets first, then every chunk copies itself into place on the worker threads, which also
    // derive the packed per-node words (done serially this step was as long as the chunk generation of the largest operator)
    std::vector<size_t> node_base(nchunk + 1, 0), outer_base(nchunk + 1, 0);
    for (int c = 0; c < nchunk; c++) {
        node_base[c + 1] = node_base[c] + chunks[c].b.size();
        outer_base[c + 1] = outer_base[c] + chunks[c].H.a.size();
        g->max_outer = std::max(g->max_outer, chunks[c].max_outer);
    }
    const size_t tot_nodes = node_base[nchunk], tot_outer = outer_base[nchunk];
    g->node_off.assign(nk + 1, 0);
    g->outer_off.assign(nk + 1, 0);
    b_j0.resize(tot_nodes); n_ol.resize(tot_nodes); b.resize(tot_nodes); wgt.resize(tot_nodes); b_coef.resize(4*tot_nodes);
    H.a.resize(tot_outer); H.a_j0.resize(tot_outer); H.a_coef.resize(4*tot_outer); H.o_kid.resize(tot_outer);
    g->bs.resize(tot_nodes); g->as.resize(tot_outer); g->pack.resize(tot_nodes); g->nodew.resize(tot_nodes);
    next = 0;
    auto merger = [&]() {
        for (int ci = next++; ci < nchunk; ci = next++) {
            GeoChunk& c = chunks[ci];
            const size_t nb = node_base[ci], ob = outer_base[ci];
            try {
                for (size_t i = 1; i < c.node_off.size(); i++) g->node_off[lo[ci] + i] = (int64_t)nb + c.node_off[i];
                for (size_t i = 1; i < c.H.outer_off.size(); i++) g->outer_off[lo[ci] + i] = (int64_t)ob + c.H.outer_off[i];
                std::copy(c.b_j0.begin(), c.b_j0.end(), b_j0.begin() + nb);
                std::copy(c.n_ol.begin(), c.n_ol.end(), n_ol.begin() + nb);
                std::copy(c.b.begin(), c.b.end(), b.begin() + nb);
                std::copy(c.wgt.begin(), c.wgt.end(), wgt.begin() + nb);
                std::copy(c.b_coef.begin(), c.b_coef.end(), b_coef.begin() + 4*nb);
                std::copy(c.H.a.begin(), c.H.a.end(), H.a.begin() + ob);
                std::copy(c.H.a_j0.begin(), c.H.a_j0.end(), H.a_j0.begin() + ob);
                std::copy(c.H.a_coef.begin(), c.H.a_coef.end(), H.a_coef.begin() + 4*ob);
                for (size_t i = 0; i < c.H.o_kid.size(); i++) H.o_kid[ob + i] = c.H.o_kid[i] + lo[ci];
                // s = sum_m m c_m exactly reproduces the abscissa of a Lagrange rule of degree >= 1
                for (size_t i = ob; i < ob + c.H.a.size(); i++) g->as[i] = H.a_coef[4*i + 1] + 2.0*H.a_coef[4*i + 2] + 3.0*H.a_coef[4*i + 3];
                for (size_t i = nb; i < nb + c.b.size(); i++) {
                    g->bs[i] = b_coef[4*i + 1] + 2.0*b_coef[4*i + 2] + 3.0*b_coef[4*i + 3];
                    PRSD_REQUIRE(b_j0[i] >= 0 && b_j0[i] < 4096 && n_ol[i] >= 0 && n_ol[i] < (1 << 19), "node index does not fit the packed stream");
                    g->pack[i] = (n_ol[i] << 12) | b_j0[i];
                    const float sf = (float)g->bs[i];
                    int bits;
                    std::memcpy(&bits, &sf, sizeof(int));
                    g->nodew[i] = make_int2(g->pack[i], bits);
                }
                c = GeoChunk();                        // release the chunk's memory
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lk(fmu);
                failure = e.what();
            }
        }
    };
    if (nthreads == 1) merger();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++) pool.emplace_back(merger);
        for (auto& t : pool) t.join();
    }
    PRSD_REQUIRE(failure.empty(), "make_geometry: %s", failure.c_str());
    return g;
}

// FNV-1a over every array the device receives: pins the node order (tests, and any change of the host-side generator)
uint64_t geometry_checksum(const GeometryHost& h)
{
    uint64_t x = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t bytes) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < bytes; i++) { x ^= c[i]; x *= 1099511628211ull; }
    };
    mix(h.node_off.data(), h.node_off.size()*sizeof(int64_t));
    mix(h.outer_off.data(), h.outer_off.size()*sizeof(int64_t));
    mix(h.H.a.data(), h.H.a.size()*sizeof(double));
    mix(h.H.a_j0.data(), h.H.a_j0.size()*sizeof(int32_t));
    mix(h.H.a_coef.data(), h.H.a_coef.size()*sizeof(double));
    mix(h.H.o_kid.data(), h.H.o_kid.size()*sizeof(int32_t));
    mix(h.b.data(), h.b.size()*sizeof(double));
    mix(h.wgt.data(), h.wgt.size()*sizeof(double));
    mix(h.b_coef.data(), h.b_coef.size()*sizeof(double));
    mix(h.pack.data(), h.pack.size()*sizeof(int));
    mix(h.nodew.data(), h.nodew.size()*sizeof(int2));
    return x;
}

std::shared_ptr<NodeGeometry> upload_geometry(Context& ctx, const GeometryHost& h)
{
    auto g = std::make_shared<NodeGeometry>();
    g->nk = h.nk;
    g->k = h.k;
    g->h_node_off = h.node_off;
    g->h_outer_off = h.outer_off;
    g->max_outer = h.max_outer;
    g->n_nodes = (int64_t)h.b.size();
    g->n_outer = (int64_t)h.H.a.size();
    cudaStream_t s = ctx.stream;
    // (the stencil starts, Lagrange coefficients, coordinates and index words of the inner nodes stay on the host: the
    // kernels read them packed in b_node; 48 bytes per node less to upload and to hold)
    g->outer_off.upload(g->h_outer_off, s);
    g->node_off.upload(g->h_node_off, s);
    g->a_j0.upload(h.H.a_j0, s);
    g->a.upload(h.H.a, s);
    g->a_coef.upload(h.H.a_coef, s);
    g->n_outer_local.upload(h.n_ol, s);
    g->b.upload(h.b, s);
    g->wgt.upload(h.wgt, s);
    g->a_s.upload(h.as, s);
    g->b_node.upload(h.nodew, s);
    g->kdev.upload(g->k, s);
    return g;       // copies from pageable memory are staged before cudaMemcpyAsync returns: h may be released
}

std::shared_ptr<NodeGeometry> make_geometry(Context& ctx, const double* k, int nk, double qmax,
                                            const QuadConfig& cfg, const KnotGrid& grid, int bank_mod)
{
    auto h = make_geometry_host(k, nk, qmax, cfg, grid, bank_mod);
    auto g = upload_geometry(ctx, *h);
    ctx.sync();
    return g;
}

// ------------------------------------------------------------------ W build
// Fragment order: the B fragment of n-tile j for the k-step of nodes 4g .. 4g+3 is stored as 32 consecutive
// doubles in LANE order (lane = 4 (col % 8) + node % 4), so that one warp-wide load of a fragment is a single
// 256-byte coalesced access:  index(node, col) = ((node / 4) NT + col / 8) 32 + 4 (col % 8) + node % 4.
__host__ __device__ inline int64_t w_index(int64_t node, int col, int nt)
{
    return ((node >> 2)*nt + (col >> 3))*32 + ((col & 7) << 2) + (node & 3);
}

struct ColSpecDev { int kid; int sign; double fac; };

__global__ void k_build_w(int64_t n_nodes, int ncol, int ncol_used, int soa, const ColSpecDev* __restrict__ cols,
                          const int64_t* __restrict__ outer_off, const int64_t* __restrict__ node_off, int nk,
                          const double* __restrict__ kk, const double* __restrict__ a_arr,
                          const int* __restrict__ n_outer_local, const double* __restrict__ b_arr,
                          const double* __restrict__ wgt, double* __restrict__ W)
{
    const int64_t node = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    // binary search of the k this node belongs to
    int lo = 0, hi = nk - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (node_off[mid] <= node) lo = mid; else hi = mid - 1;
    }
    const double k = kk[lo];
    const double a = a_arr[outer_off[lo] + n_outer_local[node]];
    const double b = b_arr[node];
    const double w = wgt[node];
    const double u = (a + b)/k, vp = (b - a)/k;
    const double ea = 2.0*a/k, eb = 2.0*b/k;
    const int nt = ncol >> 3;
    for (int c = 0; c < ncol; c++) {
        double val = 0.0;
        if (c < ncol_used && w != 0.0) {
            const ColSpecDev cs = cols[c];
            double f = 0.0;
            if (cs.sign >= 0) f += pt_kernel(cs.kid, u, vp, ea, eb);
            if (cs.sign <= 0) f += pt_kernel(cs.kid, u, -vp, eb, ea);
            val = w*k*cs.fac*f;
        }
        if (soa) { if (c < ncol_used) W[(int64_t)c*n_nodes + node] = val; }
        else W[w_index(node, c, nt)] = val;
    }
}

__global__ void k_w_to_float(int64_t n, const double* __restrict__ src, float* __restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

void BilinearOp::build(Context& ctx, std::shared_ptr<NodeGeometry> g, const std::vector<ColSpec>& cols, bool soa_layout)
{
    geo = g;
    soa = soa_layout;
    ncol_used = (int)cols.size();
    ncol = (ncol_used + 7)/8*8;
    PRSD_REQUIRE(ncol_used > 0 && ncol <= 40, "BilinearOp: %d columns unsupported", ncol_used);
    std::vector<ColSpecDev> hc(ncol_used);
    for (int i = 0; i < ncol_used; i++) hc[i] = {cols[i].kid, cols[i].sign, cols[i].fac};
    DevBuf<ColSpecDev> dc;
    dc.upload(hc, ctx.stream);
    W.alloc((size_t)g->n_nodes*(soa ? ncol_used : ncol));
    const int threads = 128;
    const int64_t blocks = (g->n_nodes + threads - 1)/threads;
    k_build_w<<<(unsigned)blocks, threads, 0, ctx.stream>>>(g->n_nodes, ncol, ncol_used, soa ? 1 : 0, dc.p, g->outer_off.p,
        g->node_off.p, g->nk, g->kdev.p, g->a.p, g->n_outer_local.p, g->b.p, g->wgt.p, W.p);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    if (!soa) {
        // the DMMA operator streams its fragments as binary32 too (same lane order): half the bytes of the dominant
        // stream through the LSU; the entries' rounding errors (6e-8, independent from node to node) average out
        Wh.alloc(W.n);
        k_w_to_float<<<(unsigned)((W.n + 255)/256), 256, 0, ctx.stream>>>((int64_t)W.n, W.p, Wh.p);
        PRSD_CUDA(cudaGetLastError());
        ctx.launches++;
    }
    ctx.sync();
    if (!soa) W.release();
}

__global__ void k_fold_column(int64_t n, double* __restrict__ dst, const double* __restrict__ src, double coef)
{
    const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) dst[i] = fma(coef, src[i], dst[i]);
}

__global__ void k_w_to_float4(int64_t n_nodes, int ncol, const double* __restrict__ src /*[ncol][n_nodes]*/, float4* __restrict__ dst)
{
    const int64_t node = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (node >= n_nodes) return;
    float e[4];
    for (int j = 0; j < 4; j++) e[j] = (4*v + j < ncol) ? (float)src[(int64_t)(4*v + j)*n_nodes + node] : 0.f;
    dst[(int64_t)v*n_nodes + node] = make_float4(e[0], e[1], e[2], e[3]);
}

void BilinearOp::to_float(Context& ctx)
{
    PRSD_REQUIRE(soa && W.p, "to_float: only for SoA operators");
    nvec = (ncol_used + 3)/4;
    Wf.alloc((size_t)geo->n_nodes*nvec);
    k_w_to_float4<<<dim3((unsigned)((geo->n_nodes + 255)/256), nvec), 256, 0, ctx.stream>>>(geo->n_nodes, ncol_used, W.p, Wf.p);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
    ctx.sync();
    W.release();
}

void BilinearOp::fold_columns(Context& ctx, int dst, const std::vector<std::pair<int, double>>& terms, int ncol_keep)
{
    PRSD_REQUIRE(soa && dst >= 0 && dst < ncol_keep && ncol_keep <= ncol_used, "fold_columns: bad arguments");
    const int64_t n = geo->n_nodes;
    for (const auto& t : terms) {
        PRSD_REQUIRE(t.first >= 0 && t.first < ncol_used && t.first != dst, "fold_columns: bad source column");
        k_fold_column<<<(unsigned)((n + 255)/256), 256, 0, ctx.stream>>>(n, W.p + (int64_t)dst*n, W.p + (int64_t)t.first*n, t.second);
        PRSD_CUDA(cudaGetLastError());
        ctx.launches++;
    }
    ncol_used = ncol_keep;
    ncol = (ncol_used + 7)/8*8;
    ctx.sync();
}

// -------------------------------------------------------------------- apply
// One CTA (768 threads: 0.514 -> 0.497 ms against 512, 80 registers) per (k, tile of 8 rows).  Shared memory:
//   [ nslots tables (TAB_MPAD doubles each) | PaS[8][max_outer] | mbarrier ]
// and, after the node loop, the table area is reused for the cross-warp reduction.
static const int BL_THREADS = 768;
static const int BL_WARPS = BL_THREADS/32;

template <int NT>
__global__ void __launch_bounds__(BL_THREADS, 1)
k_bilinear_apply(int nk, int max_outer, const int64_t* __restrict__ outer_off, const int64_t* __restrict__ node_off,
                 const int* __restrict__ a_j0, const double* __restrict__ a_coef,
                 const int2* __restrict__ b_node, const float* __restrict__ W,
                 const double* __restrict__ tables, const int* __restrict__ slot_tab,
                 const int* __restrict__ rowA, const int* __restrict__ rowB, double* __restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tabs = reinterpret_cast<double*>(smem_raw);
    double* PaS = tabs + 8*TAB_MPAD;                                      // [8][max_outer], max_outer = padded stride (= 4 mod 16)
    double* red = PaS + 8*max_outer;                                      // [8 warps][NT][64]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + (BL_WARPS/2)*NT*64);

    // tiles of the same k are adjacent in launch order: the node stream of a k is read from DRAM once
    // and from L2 by the other tiles.  A CTA keeps its eight tables for several k (dealt round-robin over
    // blockIdx.y: the node count grows ~10x from the lowest to the highest k).
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage the (up to) 8 spectra tables of this row tile with TMA bulk copies
    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = TAB_MPAD*sizeof(double);
        mbar_expect_tx(mbar, 8*bytes);
        for (int s = 0; s < 8; s++)
            bulk_copy_g2s(tabs + s*TAB_MPAD, tables + (size_t)slot_tab[tile*8 + s]*TAB_MPAD, bytes, mbar);
    }
    mbar_wait(mbar, 0);

    const int row = lane >> 2, sub = lane & 3;
    const double* tb = tabs + rowB[tile*8 + row]*TAB_MPAD;
    const double* pa = PaS + row*max_outer;
    for (int ik = blockIdx.y; ik < nk; ik += gridDim.y) {
    // ---- P_A(a) for every outer node of this k and every row
    const int64_t o0 = outer_off[ik];
    const int n_out = (int)(outer_off[ik + 1] - o0);
    for (int idx = tid; idx < 8*n_out; idx += BL_THREADS) {
        const int o = idx >> 3, r = idx & 7;
        const double* t = tabs + rowA[tile*8 + r]*TAB_MPAD + a_j0[o0 + o];
        const double* c = a_coef + 4*(o0 + o);
        PaS[r*max_outer + o] = c[0]*t[0] + c[1]*t[1] + c[2]*t[2] + c[3]*t[3];
    }
    __syncthreads();

    // ---- node loop: each warp consumes 4 nodes per step; lane = (row = lane/4, node = lane%4)
    double acc[NT][2];
#pragma unroll
    for (int j = 0; j < NT; j++) acc[j][0] = acc[j][1] = 0.0;

    const int64_t n0 = node_off[ik], n1 = node_off[ik + 1];
    // (an explicit one-step-ahead prefetch of the node word and the fragments with the loop unrolled twice: 0.51 -> 0.71 ms;
    // unrolled four times the compiler keeps the loads of four steps in flight)
    // measured: unrolled 2x 0.60 ms, 4x 0.51 ms, 8x 0.53 ms
#pragma unroll 4
    for (int64_t base = n0 + 4*warp; base < n1; base += 4*BL_WARPS) {
        const int64_t node = base + sub;
        // one 8-byte node word (packed outer index | stencil start, stencil coordinate as binary32: the Lagrange
        // coefficients are lagrange4(s), 13 FP64 operations) and four binary32 fragments of W per step: 5 global loads
        // of 24 bytes per lane where round 1 had 8 of 72 (0.537 -> 0.51 ms; the LSU data pipe was at 68 %)
        const int2 nw = __ldg(b_node + node);
        const int j0 = nw.x & 4095, ol = nw.x >> 12;
        double c[4];
        lagrange4((double)__int_as_float(nw.y), c);
        double wf[NT];
        const float* wrow = W + (base >> 2)*(32*NT) + lane;       // base is a multiple of 4 (node lists are padded)
#pragma unroll
        for (int j = 0; j < NT; j++) wf[j] = (double)__ldg(wrow + 32*j);
        const double* t = tb + j0;
        double pb = c[0]*t[0];
        pb = fma(c[1], t[1], pb);
        pb = fma(c[2], t[2], pb);
        pb = fma(c[3], t[3], pb);
        const double pp = pa[ol]*pb;
#pragma unroll
        for (int j = 0; j < NT; j++) dmma_m8n8k4(acc[j][0], acc[j][1], pp, wf[j]);
    }

    // ---- reduce the C fragments of the 16 warps (upper half into the buffer, lower half on top) and write
    //      out[row][k][col]
    if (warp >= BL_WARPS/2) {
#pragma unroll
        for (int j = 0; j < NT; j++) {
            red[((warp - BL_WARPS/2)*NT + j)*64 + row*8 + sub*2 + 0] = acc[j][0];
            red[((warp - BL_WARPS/2)*NT + j)*64 + row*8 + sub*2 + 1] = acc[j][1];
        }
    }
    __syncthreads();
    if (warp < BL_WARPS/2) {
#pragma unroll
        for (int j = 0; j < NT; j++) {
            red[(warp*NT + j)*64 + row*8 + sub*2 + 0] += acc[j][0];
            red[(warp*NT + j)*64 + row*8 + sub*2 + 1] += acc[j][1];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NT*64; idx += BL_THREADS) {
        const int j = idx >> 6, e = idx & 63;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < BL_WARPS/2; w++) s += red[(w*NT + j)*64 + e];
        const int r = e >> 3, c = 8*j + (e & 7);
        out[((size_t)(tile*8 + r)*nk + ik)*(8*NT) + c] = s;
    }
    // the next k rewrites PaS at once (its readers passed the barriers above) and `red` only after its own
    // P(a) barrier, which the threads of the loop above reach after their reads
    }
}

template <int NT>
static void launch_apply(Context& ctx, const BilinearOp& op, const double* tables, const int* slot_tab,
                         const int* rowA, const int* rowB, int ntiles, double* out)
{
    const NodeGeometry& g = *op.geo;
    // row stride of the P(a) buffer = 4 (mod 16) doubles like the tables: rows r and r + 4 share banks, the node
    // order (make_geometry, bank_mod = 4) keeps the four nodes of a k-step on distinct residues modulo 4
    const int pas_stride = ((g.max_outer + 15)/16)*16 + 4;
    const size_t smem = (size_t)(8*TAB_MPAD + 8*pas_stride + (BL_WARPS/2)*NT*64)*sizeof(double) + 16;
    PRSD_REQUIRE(smem <= ctx.smem_optin, "bilinear apply needs %zu B shared memory (> %zu)", smem, ctx.smem_optin);
    PRSD_CUDA(cudaFuncSetAttribute(k_bilinear_apply<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // k chunks: few enough to amortise the 158 kB table staging, enough CTAs to fill 148 SMs when the batch is small
    const int nchunks = std::min(g.nk, std::max(1, (1200 + ntiles - 1)/ntiles));
    dim3 grid(ntiles, nchunks);
    ProfScope prof(ctx, PROF_BILINEAR);
    k_bilinear_apply<NT><<<grid, BL_THREADS, smem, ctx.stream>>>(g.nk, pas_stride, g.outer_off.p, g.node_off.p,
        g.a_j0.p, g.a_coef.p, g.b_node.p, op.Wh.p, tables, slot_tab, rowA, rowB, out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

void BilinearOp::apply(Context& ctx, const double* tables, const int* slot_tab, const int* rowA,
                       const int* rowB, int ntiles, double* out) const
{
    PRSD_REQUIRE(!soa && Wh.p, "BilinearOp::apply: operator was built for the nodal kernels");
    switch (ncol/8) {
        case 1: launch_apply<1>(ctx, *this, tables, slot_tab, rowA, rowB, ntiles, out); break;
        case 2: launch_apply<2>(ctx, *this, tables, slot_tab, rowA, rowB, ntiles, out); break;
        case 3: launch_apply<3>(ctx, *this, tables, slot_tab, rowA, rowB, ntiles, out); break;
        case 4: launch_apply<4>(ctx, *this, tables, slot_tab, rowA, rowB, ntiles, out); break;
        case 5: launch_apply<5>(ctx, *this, tables, slot_tab, rowA, rowB, ntiles, out); break;
        default: throw Error(ERR_INTERNAL, "BilinearOp::apply: unsupported column count");
    }
}

} // namespace prsd
