// discrete.cu -- SimpsIntegrate / TrapzIntegrate (pyRSD/_gcl/cpp/DiscreteQuad.cpp:3-66) and the discrete-quadrature
// branch of ComputeXiLM (pyRSD/_gcl/cpp/CorrelationFunction.cpp:92-125) on the device, batched over rows.
//
// One CTA per output value: the threads stride over the sample pairs / Simpson triples, a shuffle + shared-memory
// reduction finishes the sum.  The composite Simpson rule is the reference's non-uniform one (three-point
// parabola per pair of intervals); with an even number of samples it averages "Simpson on the first N-1 points +
// trapezoid on the last interval" with "trapezoid on the first interval + Simpson on the last N-1 points".
#include "discrete.cuh"

#include <cmath>

namespace prsd {

// spherical Bessel functions with the reference's small-argument series (SpecialFunctions.cpp:12-66)
__device__ __forceinline__ double sph_bessel(int l, double x)
{
    const double x2 = x*x;
    if (fabs(x) < 1e-4) {
        const double x3 = x2*x, x4 = x2*x2, x6 = x4*x2, x8 = x4*x4;
        switch (l) {
            case 0: return 1.0 - x2/6.0 + x4/120.0 - x6/5040.0;
            case 1: return x/3.0 - x3/30.0 + x4*x/840.0 - x6*x/45360.0;
            case 2: return x2/15.0 - x4/210.0 + x6/7560.0;
            case 3: return x3/105.0 - x4*x/1890.0 + x6*x/83160.0;
            case 4: return x4/945.0 - x6/20790.0;
            case 6: return x6/135135.0 - x8/4054050.0;
            default: return x8/34459425.0 - x8*x2/1309458150.0;       // l = 8
        }
    }
    double s, c;
    sincos(x, &s, &c);
    const double x3 = x2*x, x4 = x2*x2, x6 = x4*x2, x8 = x4*x4;
    switch (l) {
        case 0: return s/x;
        case 1: return (s - x*c)/x2;
        case 2: return ((3.0 - x2)*s - 3.0*x*c)/x3;
        case 3: return ((15.0 - 6.0*x2)*s - (15.0*x - x3)*c)/x4;
        case 4: return 5.0*(2.0*x2 - 21.0)*c/x4 + (x4 - 45.0*x2 + 105.0)*s/(x4*x);
        case 6: return ((-x6 + 210.0*x4 - 4725.0*x2 + 10395.0)*s - (21.0*x*(x4 - 60.0*x2 + 495.0))*c)/(x6*x);
        default: return ((x8 - 630.0*x6 + 51975.0*x4 - 945945.0*x2 + 2027025.0)*s
                         + (9.0*x*(4.0*x6 - 770.0*x4 + 30030.0*x2 - 225225.0))*c)/(x8*x);
    }
}

// one term of basic_simps (DiscreteQuad.cpp:3-28) for the triple starting at sample i
template <class Y>
__device__ __forceinline__ double simps_term(const double* __restrict__ x, const Y& y, int i)
{
    const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
    const double hsum = h0 + h1, hprod = h0*h1, r = h0/h1;
    return (1.0/6.0)*hsum*(y(i)*(2.0 - 1.0/r) + y(i + 1)*hsum*hsum/hprod + y(i + 2)*(2.0 - r));
}

// this thread's share of SimpsIntegrate / TrapzIntegrate over n samples
template <class Y>
__device__ double discrete_partial(int method, int n, const double* __restrict__ x, const Y& y, int tid, int nt)
{
    double acc = 0.0;
    if (method == 2) {                                  // TrapzIntegrate (:30-39)
        for (int i = tid; i < n - 1; i += nt) acc += (x[i + 1] - x[i])*0.5*(y(i + 1) + y(i));
        return acc;
    }
    if (n & 1) {                                        // odd: plain composite Simpson (:62-63)
        for (int i = 2*tid; i < n - 2; i += 2*nt) acc += simps_term(x, y, i);
        return acc;
    }
    // even (:47-60): 0.5 (simps[0, n-3) + simps[1, n-2)) + 0.5 (trapezoid last + trapezoid first)
    for (int i = 2*tid; i < n - 3; i += 2*nt) acc += 0.5*simps_term(x, y, i);
    for (int i = 1 + 2*tid; i < n - 2; i += 2*nt) acc += 0.5*simps_term(x, y, i);
    if (tid == 0) {
        acc += 0.25*(x[n - 1] - x[n - 2])*(y(n - 1) + y(n - 2));
        acc += 0.25*(x[1] - x[0])*(y(1) + y(0));
    }
    return acc;
}

__device__ __forceinline__ double block_sum(double v, double* red)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
    return t;
}

struct RowY {
    const double* y;
    __device__ __forceinline__ double operator()(int i) const { return y[i]; }
};

__global__ void __launch_bounds__(256)
k_discrete_integrate(int method, int n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out)
{
    __shared__ double red[8];
    RowY Y = {y + (size_t)blockIdx.x*n};
    const double t = block_sum(discrete_partial(method, n, x, Y, threadIdx.x, blockDim.x), red);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}

struct XiY {
    const double* k; const double* pk; double r, smoothing; int l, m;
    __device__ __forceinline__ double operator()(int j) const
    {
        const double kk = k[j], ks = kk*smoothing;
        return exp(-ks*ks)*pk[j]*pow(kk, (double)m)/(2.0*M_PI*M_PI)*sph_bessel(l, kk*r);
    }
};

__global__ void __launch_bounds__(256)
k_xilm_discrete(int method, int l, int m, int nk, const double* __restrict__ k, const double* __restrict__ pk,
                int nr, const double* __restrict__ r, double smoothing, double scale, double* __restrict__ out)
{
    __shared__ double red[8];
    const int ir = blockIdx.x, b = blockIdx.y;
    XiY Y = {k, pk + (size_t)b*nk, r[ir], smoothing, l, m};
    const double t = block_sum(discrete_partial(method, nk, k, Y, threadIdx.x, blockDim.x), red);
    if (threadIdx.x == 0) out[(size_t)b*nr + ir] = scale*t;
}

// LinearSpline(X, Y)(x) / .EvaluateDerivative(x)  [Spline.i:24, Spline.cpp:58-130]: piecewise-linear interpolation by
// binary search, 0 outside [X_0, X_{n-1}] like Spline::Evaluate (Spline.cpp:362-367)
__global__ void __launch_bounds__(256)
k_linear_spline(int n, const double* __restrict__ X, const double* __restrict__ Y, int nq, const double* __restrict__ xq,
                int deriv, double* __restrict__ out)
{
    const int i = blockIdx.x*blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= nq) return;
    const double x = xq[i];
    const double* y = Y + (size_t)b*n;
    double v = 0.0;
    if (x >= X[0] && x <= X[n - 1]) {
        int lo = 0, hi = n - 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (X[mid] <= x) lo = mid; else hi = mid;
        }
        const double slope = (y[lo + 1] - y[lo])/(X[lo + 1] - X[lo]);
        v = deriv ? slope : y[lo] + slope*(x - X[lo]);
    }
    out[(size_t)b*nq + i] = v;
}

void linear_spline_dev(Context& ctx, int batch, int n, const double* d_x, const double* d_y, int nq, const double* d_xq, int deriv,
                       double* d_out)
{
    k_linear_spline<<<dim3((nq + 255)/256, batch), 256, 0, ctx.stream>>>(n, d_x, d_y, nq, d_xq, deriv, d_out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

void discrete_integrate_dev(Context& ctx, int method, int batch, int n, const double* d_x, const double* d_y, double* d_out)
{
    PRSD_REQUIRE(method == 1 || method == 2, "discrete integration method must be 1 (Simpson) or 2 (trapezoid)");
    PRSD_REQUIRE(n >= 3, "discrete integration needs at least 3 samples");
    k_discrete_integrate<<<batch, 256, 0, ctx.stream>>>(method, n, d_x, d_y, d_out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

void xilm_discrete_dev(Context& ctx, int method, int batch, int l, int m, int nk, const double* d_k, const double* d_pk,
                       int nr, const double* d_r, double smoothing, double scale, double* d_out)
{
    PRSD_REQUIRE(method == 1 || method == 2, "ComputeXiLM: the integration method must be one of fftlog = 0, simps = 1, trapz = 2");
    PRSD_REQUIRE(l == 0 || l == 1 || l == 2 || l == 3 || l == 4 || l == 6 || l == 8, "ComputeXiLM: l = %d not supported", l);
    PRSD_REQUIRE(nk >= 3, "ComputeXiLM: at least 3 samples of P(k) are needed");
    k_xilm_discrete<<<dim3(nr, batch), 256, 0, ctx.stream>>>(method, l, m, nk, d_k, d_pk, nr, d_r, smoothing, scale, d_out);
    PRSD_CUDA(cudaGetLastError());
    ctx.launches++;
}

} // namespace prsd
