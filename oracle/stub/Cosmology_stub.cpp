/* TEST INFRASTRUCTURE ONLY.  Implementation of the Cosmology stand-in used to
 * compile the reference C++ without CLASS (see Cosmology_stub.h). */
#include "Quadrature.h"
#include "SpecialFunctions.h"

using namespace Common;

Cosmology::Cosmology(int n, const double* k, const double* T, double sigma8, double n_s,
                     double h, double Omega0_m, double Omega0_b, double Tcmb)
    : n_s_(n_s), h_(h), Omega0_m_(Omega0_m), Omega0_b_(Omega0_b), Tcmb_(Tcmb),
      sigma8_(sigma8), sigma8_z_(sigma8), delta_H_(1.)
{
    ki = parray(n);
    Ti = parray(n);
    for (int i = 0; i < n; i++) { ki[i] = k[i]; Ti[i] = T[i]; }

    /* Eisenstein-Hu no-wiggle parameters: Cosmology.cpp:311-362 */
    double h2 = h_*h_;
    double omh2 = Omega0_m_*h2;
    double obh2 = Omega0_b_*h2;
    double theta_cmb = Tcmb_/2.7;
    double f_baryon = obh2/omh2;
    k_equality = 0.0746*omh2/pow2(theta_cmb);
    s = h_ * 44.5*log(9.83/omh2)/sqrt(1 + 10*pow(obh2, 0.75));
    alpha_gamma = 1 - 0.328*log(431*omh2)*f_baryon + 0.38*log(22.3*omh2)*pow2(f_baryon);

    /* spline set-up: Cosmology.cpp:288-300 */
    Tk = CubicSpline(ki, Ti);
    k0 = ki[0];   k1 = ki[n-1];
    T0 = Ti[0];   T1 = Ti[n-1];
    T0_nw = GetNoWiggleTransfer(k0);
    T1_nw = GetNoWiggleTransfer(k1);

    NormalizeTransferFunction(sigma8_);
}

/* Cosmology.cpp:187-199 */
double Cosmology::EvaluateTransfer(double k) const
{
    if (k <= 0) return 0;
    else if (k <= k0) return T0 * GetNoWiggleTransfer(k)/T0_nw;
    else if (k >= k1) return T1 * GetNoWiggleTransfer(k)/T1_nw;
    else return Tk(k);
}

/* Cosmology.cpp:233-246 */
double Cosmology::GetNoWiggleTransfer(double k) const
{
    k *= h_;
    double omh2 = Omega0_m_*h_*h_;
    double ks = k*s/h_;
    double q = k/(13.41*k_equality);
    double gamma_eff = omh2*(alpha_gamma + (1 - alpha_gamma)/(1 + pow4(0.43*ks)));
    double q_eff = q*omh2/gamma_eff;
    double L0 = log(2*M_E + 1.8*q_eff);
    double C0 = 14.2 + 731.0/(1 + 62.5*q_eff);
    return L0/(L0 + C0*pow2(q_eff));
}

static double sigma8_integrand(const Cosmology& C, double R, double k)
{
    double kr = k*R;
    double j1_kr = (kr < 1e-4) ? 1/3. - kr*kr/30. : SphericalBesselJ1(kr)/kr;
    return k*k/(2*M_PI*M_PI) * pow(k, C.n_s())*pow2(C.EvaluateTransfer(k))*pow2(3*j1_kr);
}

/* Cosmology.cpp:167-184 */
void Cosmology::NormalizeTransferFunction(double sigma8)
{
    using namespace std::placeholders;
    double sigma2 = Integrate<ExpSub>(std::bind(sigma8_integrand, std::cref(*this), 8, _1),
                                      1e-5, 1e2, 1e-5, 1e-12);
    delta_H_ = sigma8/sqrt(sigma2);
}
