// exchange.cu -- see exchange.cuh
#include "exchange.cuh"

#include <algorithm>
#include <cstring>

namespace prsd {

Exchange::Exchange(Context& c, int world_, int rank_, size_t n_per_rank_) : ctx(&c), world(world_), rank(rank_), n_per_rank(n_per_rank_)
{
    PRSD_REQUIRE(world >= 1 && world <= EXCH_MAX_RANKS && rank >= 0 && rank < world, "Exchange: bad rank %d of %d", rank, world);
    PRSD_REQUIRE(n_per_rank > 0, "Exchange: empty contribution");
    PRSD_CUDA(cudaMalloc(&recv_local, 2*(size_t)world*n_per_rank*sizeof(double)));
    PRSD_CUDA(cudaMalloc(&flags_local, EXCH_MAX_RANKS*sizeof(unsigned long long)));
    PRSD_CUDA(cudaMalloc(&d_status, sizeof(int)));
    PRSD_CUDA(cudaMemset(recv_local, 0, 2*(size_t)world*n_per_rank*sizeof(double)));
    PRSD_CUDA(cudaMemset(flags_local, 0, EXCH_MAX_RANKS*sizeof(unsigned long long)));
    PRSD_CUDA(cudaMemset(d_status, 0, sizeof(int)));
    PRSD_CUDA(cudaDeviceSynchronize());
    recv_peer[rank] = recv_local;
    flags_peer[rank] = flags_local;
    if (world == 1) opened = true;
}

Exchange::~Exchange()
{
    cudaDeviceSynchronize();
    for (int p = 0; p < world; p++) {
        if (p == rank) continue;
        if (recv_peer[p]) cudaIpcCloseMemHandle(recv_peer[p]);
        if (flags_peer[p]) cudaIpcCloseMemHandle(flags_peer[p]);
    }
    cudaFree(recv_local);
    cudaFree(flags_local);
    cudaFree(d_status);
}

void Exchange::handles(unsigned char* out128) const
{
    static_assert(2*sizeof(cudaIpcMemHandle_t) <= EXCH_HANDLE_BYTES, "handle block too small");
    cudaIpcMemHandle_t h[2];
    PRSD_CUDA(cudaIpcGetMemHandle(&h[0], recv_local));
    PRSD_CUDA(cudaIpcGetMemHandle(&h[1], flags_local));
    std::memset(out128, 0, EXCH_HANDLE_BYTES);
    std::memcpy(out128, h, sizeof(h));
}

void Exchange::open(const unsigned char* all)
{
    PRSD_REQUIRE(!opened, "Exchange: peers are already mapped");
    for (int p = 0; p < world; p++) {
        if (p == rank) continue;
        cudaIpcMemHandle_t h[2];
        std::memcpy(h, all + (size_t)p*EXCH_HANDLE_BYTES, sizeof(h));
        void* a = nullptr;
        void* b = nullptr;
        PRSD_CUDA(cudaIpcOpenMemHandle(&a, h[0], cudaIpcMemLazyEnablePeerAccess));
        PRSD_CUDA(cudaIpcOpenMemHandle(&b, h[1], cudaIpcMemLazyEnablePeerAccess));
        recv_peer[p] = static_cast<double*>(a);
        flags_peer[p] = static_cast<unsigned long long*>(b);
    }
    opened = true;
}

static PeerTable peer_table(const Exchange& x)
{
    PeerTable t;
    t.world = x.world; t.rank = x.rank;
    const size_t parity = (size_t)(x.step & 1)*x.world*x.n_per_rank;
    for (int p = 0; p < EXCH_MAX_RANKS; p++) t.recv[p] = p < x.world ? x.recv_peer[p] + parity + (size_t)x.rank*x.n_per_rank : nullptr;
    return t;
}

double* Exchange::begin_step()
{
    PRSD_REQUIRE(opened, "Exchange: peers have not been mapped (prsd_exchange_open)");
    step++;
    return peer_table(*this).recv[rank];
}

// blockIdx.y = destination peer (the own slot is the source): coalesced 8-byte stores over NVLink, one system-scope fence
// per thread at the end (the flag kernel that follows in the stream publishes them)
__global__ void __launch_bounds__(256)
k_exchange_push(size_t n, const PeerTable t)
{
    const int p = blockIdx.y;
    if (p == t.rank) return;
    const double* __restrict__ src = t.recv[t.rank];
    double* __restrict__ dst = t.recv[p];
    for (size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

struct FlagTable { unsigned long long* peer[EXCH_MAX_RANKS]; };

// thread p: make this rank's stores of the step visible system-wide, raise its flag at peer p, then wait until peer p has
// raised its flag here.  A wait is bounded (2 s): a lost peer sets the status word instead of hanging the device.
__global__ void k_exchange_signal_wait(int world, int rank, unsigned long long step, FlagTable peers,
                                       volatile unsigned long long* mine, int* status)
{
    const int p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();
    if (p != rank) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(peers.peer[p] + rank), "l"(step) : "memory");
    } else {
        mine[rank] = step;
    }
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + p) : "memory");
        if (v >= step) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) { atomicExch(status, 1); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

void Exchange::finish_step()
{
    if (world > 1) {
        const unsigned bx = (unsigned)std::min<size_t>(64, (n_per_rank + 255)/256);
        k_exchange_push<<<dim3(bx, world), 256, 0, ctx->stream>>>(n_per_rank, peer_table(*this));
        PRSD_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    FlagTable f;
    for (int p = 0; p < EXCH_MAX_RANKS; p++) f.peer[p] = p < world ? flags_peer[p] : nullptr;
    k_exchange_signal_wait<<<1, 32, 0, ctx->stream>>>(world, rank, step, f, flags_local, d_status);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
}

int Exchange::status()
{
    int s = 0;
    PRSD_CUDA(cudaMemcpyAsync(&s, d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    return s;
}

} // namespace prsd
