// fftlog_host.h -- host-side set-up of the discrete log-space Hankel transform.
//
// The reference calls Hamilton's FFTLog (pyRSD/_gcl/extern/fftlog/fftlog.f) through
// FortranFFTLog (pyRSD/_gcl/cpp/FortranFFTLog.cpp:10-33): fhti builds the Fourier-space
// multipliers u_m = (kr)^{-2 i y_m} U_mu(q + 2 i y_m), y_m = pi m / (N dlnr), from complex
// log-Gamma functions (fftlog.f:358-470), optionally after snapping kr to the nearest
// low-ringing value (krgood, fftlog.f:787-834).  These depend only on (N, mu, q, dlnr, kr),
// never on the data, so they are computed once here (in double-double-free plain FP64, own
// log-Gamma: shift + Stirling) and kept on the device.
#pragma once
#include <complex>
#include <vector>

namespace prsd {

std::complex<double> lngamma_complex(std::complex<double> z);

// nearest low-ringing kr (kropt = 1)
double fftlog_krgood(double mu, double q, double dlnr, double kr);

// Fourier multipliers for m = 0 .. N/2 (complex, including the amplitude for q != 0).
// For even N the m = N/2 entry is REAL (fhtq multiplies the Nyquist mode by the real part only,
// fftlog.f:700-716) -- its imaginary part is returned as 0.
void fftlog_multipliers(int N, double mu, double q, double dlnr, double kr, std::vector<double>& re,
                        std::vector<double>& im);

} // namespace prsd
