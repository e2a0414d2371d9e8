// common.h -- error plumbing, device context and small RAII helpers shared by the library.
//
// Error convention of the C-ABI (replaces the reference's SWIG exception translation,
// pyRSD/gcl.i:22-30, and its abort()-on-bad-index, Common.cpp:67-74): every entry point
// returns PRSD_OK (0) or a negative code and never aborts; prsd_last_error() returns the
// message for the calling thread.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace prsd {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

enum {
    ERR_OK = 0,
    ERR_INVALID = -1,     // bad argument (e.g. invalid (m, n) indices)
    ERR_CUDA = -2,        // CUDA runtime failure
    ERR_NO_DEVICE = -3,   // no usable GPU: there is no CPU fallback by design
    ERR_INTERNAL = -4,
};

inline std::string strprintf(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string(buf);
}

#define PRSD_CUDA(expr)                                                                     \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            throw ::prsd::Error(::prsd::ERR_CUDA, ::prsd::strprintf("%s failed: %s (%s:%d)", #expr, \
                                cudaGetErrorString(e_), __FILE__, __LINE__));               \
    } while (0)

#define PRSD_REQUIRE(cond, ...)                                                             \
    do {                                                                                    \
        if (!(cond)) throw ::prsd::Error(::prsd::ERR_INVALID, ::prsd::strprintf(__VA_ARGS__)); \
    } while (0)

// Device memory comes from a small caching allocator (context.cu): cudaMalloc / cudaFree cost 0.1-1 ms each and
// cudaFree synchronises the device, which would serialise every call that owns temporaries.  The allocator is
// stream-ordered: a block freed under a context is reusable at once for work on the same context's stream (safe
// by stream order), and only after an allocator-issued device synchronisation by anyone else.
cudaStream_t pool_set_stream(cudaStream_t s) noexcept;     // allocation stream of the calling thread; returns the previous one
void* pool_alloc(size_t bytes);
void  pool_free(void* p) noexcept;
void  pool_trim(int device = -1) noexcept;   // give the cached blocks of one device (or all) back to the driver

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
        if (p) pool_free(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t n_)
    {
        release();
        n = n_;
        if (n) p = static_cast<T*>(pool_alloc(n*sizeof(T)));
    }
    void ensure(size_t n_) { if (n_ > n) alloc(n_); }
    void upload(const T* h, size_t cnt, cudaStream_t s)
    {
        ensure(cnt);
        if (cnt) PRSD_CUDA(cudaMemcpyAsync(p, h, cnt*sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
    void download(T* h, size_t cnt, cudaStream_t s) const
    {
        if (cnt) PRSD_CUDA(cudaMemcpyAsync(h, p, cnt*sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    void zero(cudaStream_t s) { if (n) PRSD_CUDA(cudaMemsetAsync(p, 0, n*sizeof(T), s)); }
};

enum ProfTag { PROF_BILINEAR = 0, PROF_LINEAR, PROF_ZELDOVICH, PROF_GALAXY, PROF_SPLINE, PROF_FFTLOG, PROF_NODAL_ML, PROF_NODAL_DIAG, PROF_NTAGS };

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    size_t smem_optin = 0;
    long long launches = 0;        // kernels launched through this context (bench bookkeeping)
    // optional per-kernel-family timing with CUDA events on `stream` (bench.py roofline)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pending[PROF_NTAGS];
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[PROF_NTAGS] = {0};
    long long prof_count[PROF_NTAGS] = {0};
    explicit Context(int dev);
    ~Context();
    void sync() { PRSD_CUDA(cudaStreamSynchronize(stream)); }
    // second stream for host -> device input copies of the asynchronous entry points: the inputs of batch i + 1 travel
    // while the kernels of batch i run (created on first use)
    cudaStream_t upload_stream = nullptr;
    cudaStream_t uploads();
    cudaEvent_t prof_event();
    void prof_collect();           // sync + fold pending event pairs into prof_ms / prof_count
    void prof_reset();
};

// records a start event at construction and a stop event at destruction when profiling is on
struct ProfScope {
    Context& c;
    int tag;
    cudaEvent_t e0 = nullptr;
    ProfScope(Context& ctx, int t) : c(ctx), tag(t)
    {
        if (c.profiling) { e0 = c.prof_event(); cudaEventRecord(e0, c.stream); }
    }
    ~ProfScope()
    {
        if (e0) { cudaEvent_t e1 = c.prof_event(); cudaEventRecord(e1, c.stream); c.prof_pending[tag].push_back({e0, e1}); }
    }
};

} // namespace prsd
