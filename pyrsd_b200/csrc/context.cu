// context.cu -- device context.
#include "common.h"

#include <map>
#include <mutex>
#include <unordered_map>

namespace prsd {

// ------------------------------------------------------------------ caching allocator
// Stream-ordered: a block remembers the stream of the context it was allocated under (the C-ABI entry points set
// it for the calling thread, StreamScope).  Everything a context does runs on that one stream, so a block freed
// by the host while kernels that use it are still queued can be handed out again at once for work on the SAME
// stream -- stream order makes the reuse safe, and no host-side synchronisation is needed (the per-step path
// of an asynchronous pipeline never blocks in the allocator).  Blocks allocated outside any context, or wanted
// on a different stream than they were used on, go through the conservative path: reusable only after a
// cudaDeviceSynchronize() issued by the allocator itself.
namespace {
struct Block { size_t bytes; int device; cudaStream_t stream; };
struct Key {
    int device; cudaStream_t stream; size_t bytes;
    bool operator<(const Key& o) const
    {
        if (device != o.device) return device < o.device;
        if (stream != o.stream) return stream < o.stream;
        return bytes < o.bytes;
    }
};
struct Pool {
    std::mutex mu;
    std::unordered_map<void*, Block> live;                     // every block handed out
    std::multimap<Key, void*> ready;                           // reusable blocks (stream == nullptr: by anyone)
    std::vector<std::pair<void*, Block>> pending;              // freed, untagged: work using them may be in flight
    size_t cached_bytes = 0;                                   // bytes held in `ready` + `pending`
};
Pool& pool() { static Pool* p = new Pool; return *p; }         // never destroyed: static DevBufs outlive main()

size_t round_up(size_t b) { return (b + 511) & ~size_t(511); }
thread_local cudaStream_t tl_stream = nullptr;
}

cudaStream_t pool_set_stream(cudaStream_t s) noexcept
{
    cudaStream_t prev = tl_stream;
    tl_stream = s;
    return prev;
}

void* pool_alloc(size_t bytes)
{
    Pool& P = pool();
    int dev = 0;
    PRSD_CUDA(cudaGetDevice(&dev));
    const size_t want = round_up(bytes);
    const cudaStream_t st = tl_stream;
    std::lock_guard<std::mutex> lk(P.mu);
    auto take = [&](cudaStream_t from) -> void* {
        auto it = P.ready.lower_bound({dev, from, want});
        if (it != P.ready.end() && it->first.device == dev && it->first.stream == from && it->first.bytes <= want + want/4) {
            void* p = it->second;
            P.live[p] = {it->first.bytes, dev, st};
            P.cached_bytes -= it->first.bytes;
            P.ready.erase(it);
            return p;
        }
        return nullptr;
    };
    if (st) if (void* p = take(st)) return p;          // same stream: safe by stream order
    if (void* p = take(nullptr)) return p;             // synchronised blocks: safe for anyone
    if (!P.pending.empty()) {
        // untagged blocks freed since the last synchronisation become reusable by anyone now
        bool mine = false;
        for (auto& e : P.pending) mine = mine || e.second.device == dev;
        if (mine) {
            PRSD_CUDA(cudaDeviceSynchronize());           // the CURRENT device: only its blocks move
            std::vector<std::pair<void*, Block>> keep;
            for (auto& e : P.pending) {
                if (e.second.device == dev) P.ready.insert({{dev, nullptr, e.second.bytes}, e.first});
                else keep.push_back(e);
            }
            P.pending.swap(keep);
            if (void* p = take(nullptr)) return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        // out of memory: synchronise, return THIS device's cache to the driver and retry once
        cudaGetLastError();
        cudaDeviceSynchronize();
        for (auto it = P.ready.begin(); it != P.ready.end();) {
            if (it->first.device == dev) { cudaFree(it->second); P.cached_bytes -= it->first.bytes; it = P.ready.erase(it); }
            else ++it;
        }
        std::vector<std::pair<void*, Block>> keep;
        for (auto& pe : P.pending) {
            if (pe.second.device == dev) { cudaFree(pe.first); P.cached_bytes -= pe.second.bytes; }
            else keep.push_back(pe);
        }
        P.pending.swap(keep);
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess)
        throw Error(ERR_CUDA, strprintf("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)));
    P.live[p] = {want, dev, st};
    return p;
}

void pool_free(void* p) noexcept
{
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) return;
    const Block b = it->second;
    P.live.erase(it);
    P.cached_bytes += b.bytes;
    if (b.stream) P.ready.insert({{b.device, b.stream, b.bytes}, p});
    else P.pending.push_back({p, b});
    // bound the cache (varying batch sizes would otherwise accumulate blocks of every size ever used)
    static const size_t CACHE_LIMIT = size_t(24) << 30;
    if (P.cached_bytes > CACHE_LIMIT) {
        int prev = -1;
        cudaGetDevice(&prev);
        for (auto& r : P.ready) { cudaSetDevice(r.first.device); cudaFree(r.second); P.cached_bytes -= r.first.bytes; }   // cudaFree synchronises
        P.ready.clear();
        if (prev >= 0) cudaSetDevice(prev);
    }
}

void pool_trim(int device) noexcept
{
    // gives the cached blocks of ONE device back to the driver (device < 0: all of them); blocks of other devices
    // may belong to live contexts whose streams are still running
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    int prev = -1;
    cudaGetDevice(&prev);
    std::vector<std::pair<void*, Block>> keep;
    for (auto& e : P.pending) {
        if (device < 0 || e.second.device == device) {
            cudaSetDevice(e.second.device);
            cudaFree(e.first);                       // cudaFree synchronises the owning device
            P.cached_bytes -= e.second.bytes;
        } else keep.push_back(e);
    }
    P.pending.swap(keep);
    for (auto it = P.ready.begin(); it != P.ready.end();) {
        if (device < 0 || it->first.device == device) {
            cudaSetDevice(it->first.device);
            cudaFree(it->second);
            P.cached_bytes -= it->first.bytes;
            it = P.ready.erase(it);
        } else ++it;
    }
    if (prev >= 0) cudaSetDevice(prev);
}

Context::Context(int dev) : device(dev)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(ERR_NO_DEVICE, strprintf("no CUDA device available (%s): pyrsd_b200 has no CPU fallback",
                                             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)));
    if (dev < 0 || dev >= count) throw Error(ERR_INVALID, strprintf("device %d out of range (0..%d)", dev, count - 1));
    PRSD_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    PRSD_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        throw Error(ERR_NO_DEVICE, strprintf("device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                                             dev, prop.major, prop.minor));
    sm_count = prop.multiProcessorCount;
    smem_optin = prop.sharedMemPerBlockOptin;
    PRSD_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
}

Context::~Context()
{
    for (int t = 0; t < PROF_NTAGS; t++)
        for (auto& pr : prof_pending[t]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : prof_pool) cudaEventDestroy(e);
    if (upload_stream) { cudaStreamSynchronize(upload_stream); cudaStreamDestroy(upload_stream); }
    if (stream) cudaStreamDestroy(stream);
    pool_trim(device);
}

cudaStream_t Context::uploads()
{
    if (!upload_stream) PRSD_CUDA(cudaStreamCreateWithFlags(&upload_stream, cudaStreamNonBlocking));
    return upload_stream;
}

cudaEvent_t Context::prof_event()
{
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e;
    PRSD_CUDA(cudaEventCreate(&e));
    return e;
}

void Context::prof_collect()
{
    sync();
    for (int t = 0; t < PROF_NTAGS; t++) {
        for (auto& pr : prof_pending[t]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { prof_ms[t] += ms; prof_count[t]++; }
            prof_pool.push_back(pr.first);
            prof_pool.push_back(pr.second);
        }
        prof_pending[t].clear();
    }
}

void Context::prof_reset()
{
    prof_collect();
    for (int t = 0; t < PROF_NTAGS; t++) { prof_ms[t] = 0; prof_count[t] = 0; }
}

} // namespace prsd
