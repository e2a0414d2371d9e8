/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into the product library.
 *
 * Plain-C restatement of the parts of Hamilton's FFTLog that the reference
 * calls through `FortranFFTLog` (pyRSD/_gcl/cpp/FortranFFTLog.cpp:5-33):
 *
 *   fhti_   pyRSD/_gcl/extern/fftlog/fftlog.f:233-470   (set-up, u_m table)
 *   fht_    pyRSD/_gcl/extern/fftlog/fftlog.f:565-641   (power-law bias wrapper)
 *   fhtq    pyRSD/_gcl/extern/fftlog/fftlog.f:644-784   (rfft, x u_m, irfft, reverse, /n)
 *   krgood  pyRSD/_gcl/extern/fftlog/fftlog.f:787-834   (low-ringing kr)
 *   cdgamma pyRSD/_gcl/extern/fftlog/cdgamma.f          (complex ln Gamma; restated as
 *                                                        shift + Stirling series)
 *   drfftf/drfftb (FFTPACK half-complex real FFT)       restated as a radix-2 FFT
 *                                                        (O(n^2) DFT when n is not 2^p)
 *
 * The reference cannot build its Fortran here (no gfortran), so these symbols
 * are exported with the Fortran ABI (`fhti_`, `fht_`) and linked under the
 * UNMODIFIED reference C++ (ZeldovichPS.cpp, CorrelationFunction.cpp,
 * FortranFFTLog.cpp) to form oracle/_ref/libgcl_ref.so.  The restatement is
 * pinned by the reference's shipped Zel'dovich tables (tests/golden), see
 * tests/test_oracle_golden.py.
 *
 * wsave layout follows fftlog.f: first 2n+15 doubles reserved for the FFT
 * (unused here), then q, dlnr, kr, then the u_m table.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

/* ---- complex log-gamma: ln Gamma(x + i y), x > 0 ------------------------- */
static double complex ln_gamma_c(double complex z)
{
    /* shift the argument up until |z| is large, then Stirling */
    double complex shift = 0.0;
    while (creal(z) < 12.0) {
        shift += clog(z);
        z += 1.0;
    }
    static const double B[] = {
        1.0/12.0, -1.0/360.0, 1.0/1260.0, -1.0/1680.0, 1.0/1188.0,
        -691.0/360360.0, 1.0/156.0, -3617.0/122400.0
    };
    double complex zi = 1.0/z, zi2 = zi*zi, term = zi, series = 0.0;
    for (int j = 0; j < 8; j++) {
        series += B[j]*term;
        term *= zi2;
    }
    return (z - 0.5)*clog(z) - z + 0.5*log(2.0*M_PI) + series - shift;
}

/* ---- real FFT in FFTPACK half-complex layout ------------------------------
 * forward : r[0] = sum a ; r[2m-1] = Re c_m ; r[2m] = Im c_m ; r[n-1] = c_{n/2} (n even)
 *           with c_m = sum_j a_j exp(-2 pi i j m / n)
 * backward: the unnormalised inverse of the above                                    */
static int is_pow2(int n) { return n > 0 && (n & (n-1)) == 0; }

static void cfft_pow2(int n, double complex* x, int sign)
{
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double complex t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = sign*2.0*M_PI/len;
        for (int i = 0; i < n; i += len) {
            for (int j = 0; j < len/2; j++) {
                double complex w = cos(ang*j) + I*sin(ang*j);
                double complex u = x[i+j], v = x[i+j+len/2]*w;
                x[i+j] = u + v;
                x[i+j+len/2] = u - v;
            }
        }
    }
}

static void cdft(int n, double complex* x, int sign)
{
    if (is_pow2(n)) { cfft_pow2(n, x, sign); return; }
    double complex* y = (double complex*)malloc(sizeof(double complex)*n);
    for (int m = 0; m < n; m++) {
        double complex s = 0;
        for (int j = 0; j < n; j++) {
            long jm = ((long)j*m) % n;
            double ang = sign*2.0*M_PI*jm/n;
            s += x[j]*(cos(ang) + I*sin(ang));
        }
        y[m] = s;
    }
    memcpy(x, y, sizeof(double complex)*n);
    free(y);
}

static void rfft_forward(int n, double* a)
{
    double complex* x = (double complex*)malloc(sizeof(double complex)*n);
    for (int i = 0; i < n; i++) x[i] = a[i];
    cdft(n, x, -1);
    a[0] = creal(x[0]);
    for (int m = 1; m <= (n-1)/2; m++) {
        a[2*m-1] = creal(x[m]);
        a[2*m]   = cimag(x[m]);
    }
    if (n % 2 == 0) a[n-1] = creal(x[n/2]);
    free(x);
}

static void rfft_backward(int n, double* a)
{
    double complex* x = (double complex*)malloc(sizeof(double complex)*n);
    x[0] = a[0];
    for (int m = 1; m <= (n-1)/2; m++) {
        x[m]   = a[2*m-1] + I*a[2*m];
        x[n-m] = a[2*m-1] - I*a[2*m];
    }
    if (n % 2 == 0) x[n/2] = a[n-1];
    cdft(n, x, +1);
    for (int i = 0; i < n; i++) a[i] = creal(x[i]);
    free(x);
}

/* ---- krgood: fftlog.f:787-834 ------------------------------------------- */
static double krgood(double mu, double q, double dlnr, double kr)
{
    if (dlnr == 0.0) return kr;
    double xp = (mu + 1.0 + q)/2.0;
    double xm = (mu + 1.0 - q)/2.0;
    double y = M_PI/(2.0*dlnr);
    double complex zp = ln_gamma_c(xp + I*y);
    double complex zm = ln_gamma_c(xm + I*y);
    double arg = log(2.0/kr)/dlnr + (cimag(zp) + cimag(zm))/M_PI;
    double iarg = round(arg);
    if (arg != iarg) kr = kr*exp((arg - iarg)*dlnr);
    return kr;
}

/* ---- fhti: fftlog.f:233-470 (kropt 0/1 only; 2/3 are interactive) -------- */
void fhti_(int* n_, double* mu_, double* q_, double* dlnr_, double* kr_, int* kropt_,
           double* wsave, _Bool* ok)
{
    int n = *n_;
    double mu = *mu_, q = *q_, dlnr = *dlnr_;
    if (*kropt_ != 0) *kr_ = krgood(mu, q, dlnr, *kr_);
    double kr = *kr_;
    *ok = 1;
    if (n <= 0) return;

    int l = 2*n + 15;
    wsave[l]   = q;
    wsave[l+1] = dlnr;
    wsave[l+2] = kr;
    l += 3;
    double* w = wsave + l;          /* w[0] == fortran wsave(l+1) */
    double ln2kr = log(2.0/kr);
    double d = M_PI/(n*dlnr);
    if (q == 0.0) {
        double xp = (mu + 1.0)/2.0;
        for (int m = 1; m <= n/2; m++) {
            double y = m*d;
            double complex zp = ln_gamma_c(xp + I*y);
            double arg = 2.0*(ln2kr*y + cimag(zp));
            w[2*m-2] = cos(arg);
            w[2*m-1] = sin(arg);
        }
    } else {
        double ln2 = log(2.0);
        double xp = (mu + 1.0 + q)/2.0;
        double xm = (mu + 1.0 - q)/2.0;
        double amp, arg;
        int xp_neg = (round(xp) == xp && round(xp) <= 0.0);
        int xm_neg = (round(xm) == xm && round(xm) <= 0.0);
        if (xp_neg || xm_neg) {
            if (xp_neg && xm_neg) {
                amp = exp(ln2*q);
                if (xp > xm) {
                    int cnt = (int)lround(xp - xm);
                    for (int m = 1; m <= cnt; m++) amp *= (xm + m - 1.0);
                } else if (xp < xm) {
                    int cnt = (int)lround(xm - xp);
                    for (int m = 1; m <= cnt; m++) amp /= (xp + m - 1.0);
                }
                arg = round(xp + xm)*M_PI;
            } else {
                amp = 0.0;
                arg = 0.0;
            }
        } else {
            double complex zp = ln_gamma_c(xp + I*0.0);
            double complex zm = ln_gamma_c(xm + I*0.0);
            amp = exp(ln2*q + creal(zp) - creal(zm));
            arg = cimag(zp) + cimag(zm);
        }
        w[0] = amp*cos(arg);
        for (int m = 1; m <= n/2; m++) {
            double y = m*d;
            double complex zp = ln_gamma_c(xp + I*y);
            double complex zm = ln_gamma_c(xm + I*y);
            amp = exp(ln2*q + creal(zp) - creal(zm));
            arg = 2.0*ln2kr*y + cimag(zp) + cimag(zm);
            w[3*m-2] = amp;
            w[3*m-1] = cos(arg);
            w[3*m]   = sin(arg);
        }
    }
}

/* ---- fhtq: fftlog.f:644-784 --------------------------------------------- */
static void fhtq(int n, double* a, int dir, const double* wsave)
{
    int l = 2*n + 15;
    double q = wsave[l];
    l += 3;
    const double* w = wsave + l;     /* w[j-1] == fortran wsave(l+j) */

    rfft_forward(n, a);
    /* fortran a(2m), a(2m+1) == C a[2m-1], a[2m] */
    if (q == 0.0) {
        for (int m = 1; m <= (n-1)/2; m++) {
            double ar = a[2*m-1], ai = a[2*m];
            double c = w[2*m-2], s = w[2*m-1];
            a[2*m-1] = ar*c - ai*s;
            a[2*m]   = ar*s + ai*c;
        }
        if (n % 2 == 0) {
            double ar = w[n-2];      /* wsave(l+n-1): cos(arg) of m = n/2 */
            if (dir == 1) a[n-1] *= ar;
            else if (dir == -1) a[n-1] /= ar;
        }
    } else {
        for (int m = 1; m <= (n-1)/2; m++) {
            double ar = a[2*m-1], ai = a[2*m];
            double c = w[3*m-1], s = w[3*m];
            a[2*m-1] = ar*c - ai*s;
            a[2*m]   = ar*s + ai*c;
        }
        if (dir == 1) {
            a[0] *= w[0];
            for (int m = 1; m <= (n-1)/2; m++) {
                a[2*m-1] *= w[3*m-2];
                a[2*m]   *= w[3*m-2];
            }
        } else if (dir == -1) {
            double ar = w[0];
            a[0] = (ar == 0.0) ? 0.0 : a[0]/ar;
            for (int m = 1; m <= (n-1)/2; m++) {
                a[2*m-1] /= w[3*m-2];
                a[2*m]   /= w[3*m-2];
            }
        }
        if (n % 2 == 0) {
            int m = n/2;
            double ar = w[3*m-1]*w[3*m-2];
            if (dir == 1) a[n-1] *= ar;
            else if (dir == -1) a[n-1] /= ar;
        }
    }
    rfft_backward(n, a);
    /* reverse the array and undo the FFTs' factor n */
    for (int m = 0; m < n/2; m++) {
        double ar = a[m];
        a[m] = a[n-1-m]/n;
        a[n-1-m] = ar/n;
    }
    if (n % 2 == 1) a[(n-1)/2] /= n;
}

/* ---- fht: fftlog.f:565-641 ---------------------------------------------- */
void fht_(int* n_, double* a, int* dir_, double* wsave)
{
    int n = *n_, dir = *dir_;
    int l = 2*n + 15;
    double q = wsave[l], dlnr = wsave[l+1], kr = wsave[l+2];
    double jc = (n + 1)/2.0;
    if (q != 0.0)
        for (int j = 1; j <= n; j++) a[j-1] *= exp(-dir*q*(j - jc)*dlnr);
    fhtq(n, a, dir, wsave);
    if (q != 0.0) {
        double lnkr = log(kr);
        for (int j = 1; j <= n; j++) a[j-1] *= exp(-dir*q*((j - jc)*dlnr + lnkr));
    }
}
