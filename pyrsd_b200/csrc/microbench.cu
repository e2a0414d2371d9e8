// microbench.cu -- FP64 peak microbenchmarks used as roofline denominators.
//
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the PT hot path is FP64 (tensor DMMA for
// the contraction, DFMA for everything else), so bench.py measures both pipes itself, once, outside
// its timed region: dependent-free chains of DFMA / mma.m8n8k4.f64 on every SM.
#include "common.h"
#include "ptx.cuh"

namespace prsd {

__global__ void __launch_bounds__(256) k_peak_dfma(int iters, double* out)
{
    double a0 = threadIdx.x*1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll 8
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x*blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) k_peak_dmma(int iters, double* out)
{
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = 0.0;
    const double a = 1.0 + threadIdx.x*1e-9, b = 1.0 - threadIdx.x*1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma_m8n8k4(c[j][0], c[j][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

// mode 0: DFMA (2 flop per lane-op), mode 1: DMMA m8n8k4 (512 flop per warp instruction)
double fp64_peak_tflops(Context& ctx, int mode)
{
    const int blocks = ctx.sm_count*8, threads = 256, iters = mode == 0 ? 4096 : 2048;
    DevBuf<double> out((size_t)blocks*threads);
    cudaEvent_t e0, e1;
    PRSD_CUDA(cudaEventCreate(&e0));
    PRSD_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        PRSD_CUDA(cudaEventRecord(e0, ctx.stream));
        if (mode == 0) k_peak_dfma<<<blocks, threads, 0, ctx.stream>>>(iters, out.p);
        else k_peak_dmma<<<blocks, threads, 0, ctx.stream>>>(iters, out.p);
        PRSD_CUDA(cudaEventRecord(e1, ctx.stream));
        PRSD_CUDA(cudaEventSynchronize(e1));
        PRSD_CUDA(cudaGetLastError());
        ctx.launches++;
        float ms = 0.f;
        PRSD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops;
        if (mode == 0) flops = (double)blocks*threads*iters*64.0*2.0;
        else flops = (double)blocks*(threads/32)*iters*8.0*512.0;
        if (rep > 0) best = std::max(best, flops/(ms*1e-3)/1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

} // namespace prsd
