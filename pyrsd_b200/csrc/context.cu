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
    if (stream) cudaStreamDestroy(stream);
}

} // namespace prsd
