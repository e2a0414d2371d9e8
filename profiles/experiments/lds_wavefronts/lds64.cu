// How many shared-memory wavefronts does a warp-wide 64-bit load cost when the 32 addresses cluster?
// One warp per CTA issues batches of independent LDS.64 with a per-lane base index pattern; cycles per load at the
// throughput limit = wavefronts per load.  nvcc -arch=sm_100a -O3 lds64.cu -o lds64 && ./lds64
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(const int* __restrict__ pat, int npat, long long* out, double* sink)
{
    __shared__ double s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
    __syncthreads();
    for (int p = 0; p < npat; p++) {
        const int j0 = pat[p*32 + (threadIdx.x & 31)];
        // integer accumulation of the low words: the loads, not the FP64 pipe, must be the limit
        unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        __syncthreads();
        long long t0 = clock64();
        for (int it = 0; it < 256; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) acc[u] ^= (unsigned long long)__double_as_longlong(s[(j0 + u*40 + it*2) & 4095]);
        }
        __syncthreads();
        long long t1 = clock64();
        unsigned long long a = 0;
        for (int u = 0; u < 8; u++) a += acc[u];
        if (threadIdx.x == 0) out[p] = t1 - t0;
        sink[(p*32 + (threadIdx.x & 31))] = (double)a;
    }
}

int main()
{
    const int NP = 7;
    int h[NP*32];
    const char* names[NP] = {"32 distinct consecutive (256 B)", "all lanes same word (broadcast)", "lane/3: 11 distinct, one 128-B window",
                             "lane/2: 16 distinct = 128 B exactly", "half-warps identical: lanes 16-31 repeat 0-15",
                             "stride 2 doubles (2-way conflict)", "lane/3 + unaligned start (straddles two windows)"};
    for (int l = 0; l < 32; l++) {
        h[0*32 + l] = l;
        h[1*32 + l] = 5;
        h[2*32 + l] = l/3;
        h[3*32 + l] = l/2;
        h[4*32 + l] = l & 15;
        h[5*32 + l] = 2*l;
        h[6*32 + l] = 11 + l/3;
    }
    int* d; long long* o; double* sink;
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, NP*sizeof(long long)); cudaMalloc(&sink, NP*32*sizeof(double));
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    k<<<1, 512>>>(d, NP, o, sink);
    k<<<1, 512>>>(d, NP, o, sink);
    long long ho[NP];
    cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    for (int p = 0; p < NP; p++) printf("%-52s %6.2f cycles per warp-wide LDS.64 (16 warps in flight)\n", names[p], ho[p]/(256.0*8*16));
    // 8 warps per CTA: throughput with several warps in flight
    return 0;
}
