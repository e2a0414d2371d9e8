// transfer.cu -- gridded (k, mu) reductions and the window-function convolution (see transfer.cuh).
#include "transfer.cuh"
#include "fftlog_host.h"

#include <cmath>
#include <complex>

namespace prsd {

// ------------------------------------------------------------------------------------------------ GridOp
// One CTA per (tile of KT k rows, group of GRID_WPC walkers): the tiles of the weights are staged in shared memory once
// and reused for every walker of the group, whose tile of P follows (coalesced loads; row stride nmu + 1 doubles, so
// that threads walk different rows without bank conflicts); every (k, bin) pair is one thread's dot product over mu.
static const int GRID_WPC = 8;

__global__ void __launch_bounds__(128)
k_grid_reduce(int nw, int nk, int nmu, int nbin, int KT, const double* __restrict__ w, const double* __restrict__ power,
              double* __restrict__ out)
{
    extern __shared__ double sm_grid[];
    const int ld = nmu + 1;
    double* Pt = sm_grid;                         // [KT][ld]
    double* Wt = sm_grid + (size_t)KT*ld;         // [nbin][KT][ld]
    const int k0 = blockIdx.x*KT, tid = threadIdx.x;
    const int rows = min(KT, nk - k0);
    const int cells = rows*nmu;
    for (int b = 0; b < nbin; b++) {
        const double* wb = w + ((size_t)b*nk + k0)*nmu;
        for (int i = tid; i < cells; i += blockDim.x) {
            const int r = i/nmu, m = i - r*nmu;
            Wt[((size_t)b*KT + r)*ld + m] = wb[i];
        }
    }
    const int w_end = min(nw, (int)(blockIdx.y + 1)*GRID_WPC);
    for (int wlk = blockIdx.y*GRID_WPC; wlk < w_end; wlk++) {
        const double* P = power + ((size_t)wlk*nk + k0)*nmu;
        __syncthreads();                          // previous walker's dot products are done with Pt
        for (int i = tid; i < cells; i += blockDim.x) {
            const int r = i/nmu, m = i - r*nmu;
            Pt[r*ld + m] = P[i];
        }
        __syncthreads();
        for (int item = tid; item < rows*nbin; item += blockDim.x) {
            const int b = item/rows, r = item - b*rows;
            const double* pr = Pt + r*ld;
            const double* wr = Wt + ((size_t)b*KT + r)*ld;
            double s = 0.0;
            for (int m = 0; m < nmu; m++) {
                const double ww = wr[m];
                if (ww != 0.0) s = fma(ww, pr[m], s);    // weight 0 = null grid point: P may be anything there
            }
            out[((size_t)wlk*nbin + b)*nk + k0 + r] = s;
        }
    }
}

// nmu <= 64: one WARP per (walker, k) row.  The row's nmu values are one coalesced read (two elements per lane), kept in
// registers for every bin; the weights come from L2 (the operator is a few hundred kB); a five-step butterfly finishes
// the dot product.  Measured on the benchmark's 256 x 1000 x 40 grid (82 MB read, inputs > L2): 0.035 ms = 2.3 TB/s (0.047
// with a 64-bit index division per warp -- ncu: 68 % of the issue slots at 22 % of the DRAM bandwidth,
// profiles/r02_ncu_build_and_reduce.json); the staged kernel above, which does the dot products of a 32-row tile on 32 of its
// 128 threads, 0.054 ms.
__global__ void __launch_bounds__(256)
k_grid_reduce_rows(int nw, int nk, int nmu, int nbin, const double* __restrict__ w, const double* __restrict__ power,
                   double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    // blockIdx.y = walker, 8 consecutive k per CTA: no index division (a 64-bit division per warp made the first version
    // issue-bound: ncu 68 % issue slots at 22 % of the DRAM bandwidth)
    const int wlk = blockIdx.y, k = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k < nk) {
        const size_t row = (size_t)wlk*nk + k;
        const double* P = power + (size_t)row*nmu;
        const bool has0 = lane < nmu, has1 = lane + 32 < nmu;
        const double p0 = has0 ? P[lane] : 0.0, p1 = has1 ? P[lane + 32] : 0.0;
        for (int b = 0; b < nbin; b++) {
            const double* wr = w + ((size_t)b*nk + k)*nmu;
            const double w0 = has0 ? __ldg(wr + lane) : 0.0, w1 = has1 ? __ldg(wr + lane + 32) : 0.0;
            // weight 0 = null grid point: P may be anything there
            double s = (w0 != 0.0 ? w0*p0 : 0.0) + (w1 != 0.0 ? w1*p1 : 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                const size_t idx = ((size_t)wlk*nbin + b)*nk + k;
                out[idx] = s;
            }
        }
    }
}

GridOp::GridOp(Context& c, int nbin_, int nk_, int nmu_, const double* h_w) : ctx(&c), nbin(nbin_), nk(nk_), nmu(nmu_)
{
    PRSD_REQUIRE(nbin > 0 && nk > 0 && nmu > 0, "GridOp: empty grid (%d bins, %d k, %d mu)", nbin, nk, nmu);
    for (size_t i = 0; i < (size_t)nbin*nk*nmu; i++)
        PRSD_REQUIRE(std::isfinite(h_w[i]), "GridOp: weight %zu is not finite (use 0 for null grid points)", i);
    w.upload(h_w, (size_t)nbin*nk*nmu, c.stream);
    c.sync();
}

void GridOp::apply(int nw, const double* d_power, double* d_out)
{
    if (nw <= 0) return;
    ProfScope prof(*ctx, PROF_GALAXY);
    if (nmu <= 64) {
        PRSD_REQUIRE(nw <= 65535, "GridOp: at most 65535 walkers per call (%d)", nw);
        k_grid_reduce_rows<<<dim3((nk + 7)/8, nw), 256, 0, ctx->stream>>>(nw, nk, nmu, nbin, w.p, d_power, d_out);
        PRSD_CUDA(cudaGetLastError());
        ctx->launches++;
        return;
    }
    int KT = 32;
    auto need = [&](int kt) { return (size_t)(1 + nbin)*kt*(nmu + 1)*sizeof(double); };
    while (KT > 1 && need(KT) > 96*1024) KT /= 2;
    const size_t smem = need(KT);
    PRSD_REQUIRE(smem <= ctx->smem_optin, "GridOp: %d bins x %d mu need %zu B of shared memory per k row", nbin, nmu, smem);
    if (smem > 48*1024) PRSD_CUDA(cudaFuncSetAttribute(k_grid_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_grid_reduce<<<dim3((nk + KT - 1)/KT, (nw + GRID_WPC - 1)/GRID_WPC), 128, smem, ctx->stream>>>(nw, nk, nmu, nbin, KT, w.p, d_power, d_out);
    PRSD_CUDA(cudaGetLastError());
    ctx->launches++;
}

// ---------------------------------------------------------------------------------------------- WindowOp
// Mellin transform of the spherical Bessel kernel (extern/mcfit/kernels.py:14-17)
static std::complex<double> ln_uk_sph(int l, std::complex<double> z)
{
    return std::log(2.0)*(z - 1.5) + lngamma_complex(0.5*((double)l + z)) - lngamma_complex(0.5*(3.0 + (double)l - z));
}

WindowOp::WindowOp(Context& c, int nk_, const double* k_, int nell_, const int* ells_, double q_, int N_)
    : ctx(&c), nk(nk_), nell(nell_), q(q_)
{
    PRSD_REQUIRE(nell >= 1 && nell <= 8, "window: %d multipoles (1 .. 8 supported)", nell);
    PRSD_REQUIRE(nk >= 2, "window: length of input argument is too short");
    ells.assign(ells_, ells_ + nell);
    for (int e : ells) PRSD_REQUIRE(e >= 0 && e % 2 == 0, "window: multipole %d must be even and >= 0", e);
    k.assign(k_, k_ + nk);
    // mcfit._setup (mcfit.py:104-129)
    delta = std::log(k[nk - 1]/k[0])/(nk - 1);
    for (int i = 1; i < nk; i++)
        PRSD_REQUIRE(std::fabs(k[i]/k[i - 1] - std::exp(delta)) <= 1e-3*std::exp(delta), "window: input argument must be log-evenly spaced");
    if (N_ <= 0) {
        int folds = 0;
        while ((1 << folds) < nk) folds++;
        N = 1 << (folds + 1);
    } else N = N_;
    PRSD_REQUIRE(N >= nk, "window: N is shorter than the length of input argument");
    PRSD_REQUIRE((N & (N - 1)) == 0 && N >= 1024 && N <= 8192, "window: FFT length %d must be a power of two in [1024, 8192]", N);
    const int npad = N - nk;
    off_in = npad/2;                       // mcfit.__call__ (mcfit.py:156-176)
    off_out = npad - npad/2;

    lnxy.resize(nell);
    std::vector<double> h_pre1((size_t)nell*nk), h_post1((size_t)nell*nk), h_pre2(nk), h_post2((size_t)nell*nk);
    const double twopi15 = std::pow(2.0*M_PI, 1.5);
    for (int e = 0; e < nell; e++) {
        const int l = ells[e];
        // low-ringing condition: lnxy = Delta/pi * angle(UK(q + i pi/Delta))
        const std::complex<double> lu = ln_uk_sph(l, std::complex<double>(q, M_PI/delta));
        lnxy[e] = delta/M_PI*std::remainder(lu.imag(), 2.0*M_PI);
        std::vector<double> re(N/2 + 1), im(N/2 + 1);
        for (int m = 0; m <= N/2; m++) {
            const double y = 2.0*M_PI*m/((double)N*delta);
            const std::complex<double> lw = ln_uk_sph(l, std::complex<double>(q, y));
            // u_m exp(2 pi i m / N): the extra phase turns mcfit's hfft into the reversed inverse FFT of the kernel
            const double ph = lw.imag() - lnxy[e]*y + 2.0*M_PI*m/(double)N;
            const double amp = std::exp(lw.real());
            re[m] = amp*std::cos(ph);
            im[m] = amp*std::sin(ph);
        }
        im[0] = 0.0;
        im[N/2] = 0.0;                     // hfft ignores the imaginary part at Nyquist (mcfit.py:126-128)
        plan.push_back(fftlog_plan_custom(c, N, delta, re, im));
    }
    // the reference keeps the separation grid of the last multipole for every column (window.py:98-110)
    rr.resize(nk);
    for (int i = 0; i < nk; i++) rr[i] = std::exp(lnxy[nell - 1] - delta)/k[nk - 1 - i];
    for (int e = 0; e < nell; e++) {
        const double sgn = (ells[e]/2) % 2 ? -1.0 : 1.0;               // real_if_close(1j**l)
        for (int i = 0; i < nk; i++) {
            const double y1 = std::exp(lnxy[e] - delta)/k[nk - 1 - i];       // P2xi output argument of this ell
            const double y2 = std::exp(lnxy[e] - delta)/rr[nk - 1 - i];      // xi2P output argument of this ell
            h_pre1[(size_t)e*nk + i] = sgn/twopi15*std::pow(k[i], 3.0 - q);   // prefac * x^-q (cosmology.py:22)
            h_post1[(size_t)e*nk + i] = std::pow(y1, -q);
            h_post2[(size_t)e*nk + i] = twopi15/sgn*std::pow(y2, -q);         // postfac * y^-q (cosmology.py:34-35)
        }
    }
    for (int i = 0; i < nk; i++) h_pre2[i] = std::pow(rr[i], 3.0 - q);
    pre1.upload(h_pre1, c.stream);
    post1.upload(h_post1, c.stream);
    pre2.upload(h_pre2, c.stream);
    post2.upload(h_post2, c.stream);
    c.sync();
}

void WindowOp::set_kernel(const double* h_kern)
{
    const size_t n = (size_t)nell*nell*nk;
    for (size_t i = 0; i < n; i++) PRSD_REQUIRE(std::isfinite(h_kern[i]), "window: kernel entry %zu is not finite", i);
    kern.upload(h_kern, n, ctx->stream);
    ctx->sync();
    has_kernel = true;
}

// xic[w][i][n] = sum_j kern[i][j][n] xi[w][j][n]   (WindowConvolution.__call__, rsd/window.py:165-213)
__global__ void k_window_mix(int nell, int nk, const double* __restrict__ kern, const double* __restrict__ xi,
                             double* __restrict__ xic)
{
    const int w = blockIdx.y, n = blockIdx.x*blockDim.x + threadIdx.x;
    if (n >= nk) return;
    double x[8];
    for (int j = 0; j < nell; j++) x[j] = xi[((size_t)w*nell + j)*nk + n];
    for (int i = 0; i < nell; i++) {
        double s = 0.0;
        for (int j = 0; j < nell; j++) s += x[j]*kern[((size_t)i*nell + j)*nk + n];      // einsum order of the reference
        xic[((size_t)w*nell + i)*nk + n] = s;
    }
}

void WindowOp::apply(int nw, const double* d_in, double* d_out, int mode)
{
    if (nw <= 0) return;
    PRSD_REQUIRE(mode == 0 || mode == 1, "window: mode must be 0 (convolve) or 1 (dry run)");
    PRSD_REQUIRE(mode == 1 || has_kernel, "window: set the convolution kernel first");
    Context& c = *ctx;
    const size_t n = (size_t)nw*nell*nk;
    if (xi.n < n) { xi.alloc(n); xic.alloc(n); }
    last_nw = nw;
    FFTLogWindow fw;
    fw.n_io = nk;
    fw.off_in = off_in;
    fw.off_out = off_out;
    fw.in_stride = fw.out_stride = (long long)nell*nk;
    for (int e = 0; e < nell; e++) {
        fw.pre = pre1.p + (size_t)e*nk;
        fw.post = post1.p + (size_t)e*nk;
        fftlog_apply(c, *plan[e], nw, d_in + (size_t)e*nk, xi.p + (size_t)e*nk, &fw);
    }
    const double* back = xi.p;
    if (mode == 0) {
        k_window_mix<<<dim3((nk + 127)/128, nw), 128, 0, c.stream>>>(nell, nk, kern.p, xi.p, xic.p);
        PRSD_CUDA(cudaGetLastError());
        c.launches++;
        back = xic.p;
    }
    for (int e = 0; e < nell; e++) {
        fw.pre = pre2.p;
        fw.post = post2.p + (size_t)e*nk;
        fftlog_apply(c, *plan[e], nw, back + (size_t)e*nk, d_out + (size_t)e*nk, &fw);
    }
}

} // namespace prsd
