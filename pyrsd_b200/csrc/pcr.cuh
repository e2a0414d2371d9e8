// pcr.cuh -- tridiagonal solves by parallel cyclic reduction inside one CTA.
//
// Every cubic spline on the path (the netlib spline of T(k) and of the one-loop tables, Spline.cpp:211-257; the
// not-a-knot splines of the PT tables and of the halo moments, rsd/tools.py:331-388) ends in a tridiagonal system.
// The Thomas algorithm is a chain of n dependent divisions on one thread (0.1-0.2 ms for n = 1000, the longest
// serial stretch of a single-cosmology evaluation); cyclic reduction needs ceil(log2 n) fully parallel sweeps.
// The systems here are (weakly) diagonally dominant, for which the reduction is as accurate as the elimination
// (measured: <= 3e-13 pointwise, 4e-16 of the largest unknown, profiles/experiments/r01/pcr_test.py).
#pragma once

namespace prsd {

// Solves `nsys` independent systems of n unknowns stored back to back:
//     a[i] x[i-1] + b[i] x[i] + c[i] x[i+1] = d[i]       (a[0] and c[n-1] of every system are ignored)
// a, b, c, d and the four work arrays live in shared memory and are destroyed; the solution is written to x
// (which may alias d or any of the eight arrays only if it is the LAST array read: use a separate one).
// All threads of the CTA must call this.
__device__ __forceinline__ void pcr_solve(int nsys, int n, double* a, double* b, double* c, double* d,
                                          double* a2, double* b2, double* c2, double* d2, double* x)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int sy = 0; sy < nsys; sy++) {
        if (tid == 0) { a[sy*n] = 0.0; c[sy*n + n - 1] = 0.0; }
    }
    __syncthreads();
    for (int s = 1; s < n; s <<= 1) {
        for (int sy = 0; sy < nsys; sy++) {
            const int off = sy*n;
            for (int i = tid; i < n; i += nt) {
                const int g = off + i;
                double aa = 0.0, cc = 0.0, bb = b[g], dd = d[g];
                if (i - s >= 0) {
                    const double al = -a[g]/b[g - s];
                    aa = al*a[g - s];
                    bb = fma(al, c[g - s], bb);
                    dd = fma(al, d[g - s], dd);
                }
                if (i + s < n) {
                    const double ga = -c[g]/b[g + s];
                    cc = ga*c[g + s];
                    bb = fma(ga, a[g + s], bb);
                    dd = fma(ga, d[g + s], dd);
                }
                a2[g] = aa; b2[g] = bb; c2[g] = cc; d2[g] = dd;
            }
        }
        __syncthreads();
        double* t;
        t = a; a = a2; a2 = t;
        t = b; b = b2; b2 = t;
        t = c; c = c2; c2 = t;
        t = d; d = d2; d2 = t;
    }
    for (int g = tid; g < nsys*n; g += nt) x[g] = d[g]/b[g];
    __syncthreads();
}

} // namespace prsd
