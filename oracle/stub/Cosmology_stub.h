/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Stand-in for the reference's `Cosmology` class so that the reference's
 * perturbation-theory C++ (Imn, Jmn, Kmn, ImnOneLoop, OneLoopPS, PowerSpectrum,
 * LinearPS, ZeldovichPS, ZeldovichCF, CorrelationFunction, Spline, Quadrature)
 * can be compiled FROM WHERE IT LIES under /root/reference without CLASS.
 *
 * The real class (pyRSD/_gcl/include/Cosmology.h, cpp/Cosmology.cpp) derives
 * from ClassEngine (CLASS Boltzmann solver, not vendored, not buildable here).
 * The hot path only needs five things from it, restated here:
 *   EvaluateTransfer(k)   Cosmology.cpp:153-199  spline + scaled EH no-wiggle
 *   GetNoWiggleTransfer   Cosmology.cpp:233-246
 *   EH no-wiggle params   Cosmology.cpp:360-362 (s, alpha_gamma), :327 (k_eq)
 *   delta_H normalisation Cosmology.cpp:167-184  (sigma8 at R = 8 Mpc/h)
 *   n_s(), sigma8(), Sigma8_z(z)
 *
 * It is force-included (`g++ -include`) and defines the reference header's
 * include guard, so the reference's own Cosmology.h / ClassEngine.h are skipped.
 */
#ifndef COSMOLOGY_H
#define COSMOLOGY_H

#include <cmath>
#include <functional>
#include "parray.h"
#include "Spline.h"
#include "Common.h"

class Cosmology {
public:
    /* k, T: discrete transfer function (what pygcl.Cosmology.__getstate__ stores) */
    Cosmology(int n, const double* k, const double* T, double sigma8, double n_s,
              double h, double Omega0_m, double Omega0_b, double Tcmb);
    ~Cosmology() {}

    double EvaluateTransfer(double k) const;
    double GetNoWiggleTransfer(double k) const;

    inline double n_s() const { return n_s_; }
    inline double h() const { return h_; }
    inline double Omega0_m() const { return Omega0_m_; }
    inline double Omega0_b() const { return Omega0_b_; }
    inline double Tcmb() const { return Tcmb_; }
    inline double sigma8() const { return sigma8_; }
    inline double delta_H() const { return delta_H_; }

    /* sigma8(z): the growth factor comes from CLASS in the reference; the
     * oracle takes D(z) = sigma8_z / sigma8 as a host-supplied scalar. */
    inline double Sigma8_z(double) const { return sigma8_z_; }
    inline void SetSigma8_z(double s) { sigma8_z_ = s; }

    void NormalizeTransferFunction(double sigma8);

    inline parray GetDiscreteK() const { return ki; }
    inline parray GetDiscreteTk() const { return Ti; }

private:
    double n_s_, h_, Omega0_m_, Omega0_b_, Tcmb_;
    double sigma8_, sigma8_z_, delta_H_;
    parray ki, Ti;
    double k0, T0, T0_nw, k1, T1, T1_nw;
    Spline Tk;
    double k_equality, alpha_gamma, s;
};

#endif /* COSMOLOGY_H */
