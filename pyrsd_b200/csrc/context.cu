// context.cu -- device context.
#include "common.h"

namespace prsd {

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
    if (stream) cudaStreamDestroy(stream);
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
