// fftlog_host.cpp -- see fftlog_host.h
#include "fftlog_host.h"

#include <cmath>

namespace prsd {

std::complex<double> lngamma_complex(std::complex<double> z)
{
    // ln Gamma(z) = ln Gamma(z + n) - sum_{j<n} ln(z + j), then Stirling for |z| >= 14
    std::complex<double> shift(0.0, 0.0);
    while (z.real() < 14.0) {
        shift += std::log(z);
        z += 1.0;
    }
    static const double c[] = {1.0/12, -1.0/360, 1.0/1260, -1.0/1680, 1.0/1188, -691.0/360360, 1.0/156};
    const std::complex<double> zi = 1.0/z, zi2 = zi*zi;
    std::complex<double> t = zi, s(0.0, 0.0);
    for (double cj : c) { s += cj*t; t *= zi2; }
    return (z - 0.5)*std::log(z) - z + 0.5*std::log(2.0*M_PI) + s - shift;
}

double fftlog_krgood(double mu, double q, double dlnr, double kr)
{
    if (dlnr == 0.0) return kr;
    const double xp = 0.5*(mu + 1.0 + q), xm = 0.5*(mu + 1.0 - q), y = M_PI/(2.0*dlnr);
    const double ip = lngamma_complex({xp, y}).imag(), im = lngamma_complex({xm, y}).imag();
    const double arg = std::log(2.0/kr)/dlnr + (ip + im)/M_PI;
    const double iarg = std::nearbyint(arg);
    return (arg != iarg) ? kr*std::exp((arg - iarg)*dlnr) : kr;
}

void fftlog_multipliers(int N, double mu, double q, double dlnr, double kr, std::vector<double>& re,
                        std::vector<double>& im)
{
    const int nh = N/2;
    re.assign(nh + 1, 0.0);
    im.assign(nh + 1, 0.0);
    const double ln2 = std::log(2.0), ln2kr = std::log(2.0/kr);
    const double xp = 0.5*(mu + 1.0 + q), xm = 0.5*(mu + 1.0 - q);
    const double d = M_PI/(N*dlnr);
    if (q == 0.0) {
        re[0] = 1.0;                 // the m = 0 mode is left untouched (fftlog.f:689-699)
        for (int m = 1; m <= nh; m++) {
            const double y = m*d;
            const double arg = 2.0*(ln2kr*y + lngamma_complex({xp, y}).imag());
            re[m] = std::cos(arg);
            im[m] = std::sin(arg);
        }
    } else {
        // m = 0: U_mu(q) = 2^q Gamma(xp)/Gamma(xm); singular cases as in fftlog.f:383-432
        auto negint = [](double x) { return std::nearbyint(x) == x && x <= 0.0; };
        double amp, arg;
        if (negint(xp) || negint(xm)) {
            if (negint(xp) && negint(xm)) {
                amp = std::exp(ln2*q);
                if (xp > xm) for (int m = 1; m <= (int)std::lround(xp - xm); m++) amp *= (xm + m - 1.0);
                else if (xp < xm) for (int m = 1; m <= (int)std::lround(xm - xp); m++) amp /= (xp + m - 1.0);
                arg = std::nearbyint(xp + xm)*M_PI;
            } else { amp = 0.0; arg = 0.0; }
        } else {
            const std::complex<double> zp = lngamma_complex({xp, 0.0}), zm = lngamma_complex({xm, 0.0});
            amp = std::exp(ln2*q + zp.real() - zm.real());
            arg = zp.imag() + zm.imag();
        }
        re[0] = amp*std::cos(arg);
        for (int m = 1; m <= nh; m++) {
            const double y = m*d;
            const std::complex<double> zp = lngamma_complex({xp, y}), zm = lngamma_complex({xm, y});
            const double a = std::exp(ln2*q + zp.real() - zm.real());
            const double ph = 2.0*ln2kr*y + zp.imag() + zm.imag();
            re[m] = a*std::cos(ph);
            im[m] = a*std::sin(ph);
        }
    }
    if (N % 2 == 0) im[nh] = 0.0;
}

} // namespace prsd
