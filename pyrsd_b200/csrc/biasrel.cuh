// biasrel.cuh -- bias <-> mass <-> velocity dispersion relations (pyRSD/rsd/tools.py:393-649), see biasrel.cu
#pragma once
#include "spectra.cuh"

namespace prsd {

// Sigma(R) on R = logspace(-3, 4, 1000) for every cosmology of the batch (device [ncosmo][1000])
void sigma_R_table(SpectraBatch& sp, double* d_sigma);
std::vector<double> bias_relation_rgrid();

// mass_out[w] [M_sun/h]: root of bias_Tinker(rescale[w] Sigma_c(R(M)), delta_halo) = b1[w] in [mass_lo, mass_hi], R(M) =
// (3 M / 4 pi rho_bar)^(1/3); cosmology c = cmap[w] (NULL: w, or 0 for a single-cosmology batch).  All pointers on the
// host.  sigma_out (optional, host [ncosmo][1000]) receives the Sigma(R) tables.
void bias_to_mass(SpectraBatch& sp, int nw, const int* cmap, const double* rescale, const double* b1, double rho_bar,
                  double delta_halo, double mass_lo, double mass_hi, double* mass_out, double* sigma_out);

} // namespace prsd
