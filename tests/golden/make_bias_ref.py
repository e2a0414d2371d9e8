"""Generate ``tests/golden/bias_ref.npz``: BiasToMassRelation / BiasToSigmaRelation of the UNMODIFIED reference Python
(``pyRSD.rsd.tools``, imported from /root/reference) on the reference C++ (oracle/_ref through fake_gcl.py), both in the
``compute_fresh`` mode (brentq around the adaptive Sigma(R) integral) and in the interpolation-table mode
(rsd/tools.py:393-649).

Stand-ins (as in make_galaxy_ref.py): ``rho_bar_z`` / ``H0`` come from the reference's own formulae and constants
(ClassEngine.cpp:552-582, Common.h:101-140) since CLASS is absent; PowerSpectrum::Sigma runs at TIGHT tolerance
(epsrel 1e-10 instead of its hard-coded 1e-5: 1e-5 parity is defined against the converged reference).

Run in the build container only:   python tests/golden/make_bias_ref.py   (a few minutes)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import fake_gcl  # noqa: E402
import make_galaxy_ref as G  # noqa: E402


def main():
    t0 = time.time()
    G.install_stubs()
    from pyRSD import pygcl
    from pyRSD.rsd import tools

    def rho_crit_z(self, z):
        Gn, M_sun, au = 6.67384e-8, 1.9891e33, 149597870700.*1e2
        Mpc = 1e6*au/(np.pi/180./60./60.)
        H = 100.*1e5/Mpc
        return 3.*H**2/(8.*np.pi*Gn)/(M_sun/Mpc**3)*(self.H_z(z)/self.H0())**2

    def Omega0_m(self):
        p = self._params
        return p['Omega_b'] + p['Omega_cdm'] + p.get('m_ncdm', 0.0)/93.14/p['h']**2

    fake_gcl.Cosmology.H0 = lambda self: 100.*self._params['h']
    fake_gcl.Cosmology.Omega0_m = Omega0_m
    fake_gcl.Cosmology.rho_crit_z = rho_crit_z
    fake_gcl.Cosmology.rho_bar_z = lambda self, z: self.Omega0_m()*(1. + z)**3/(self.H_z(z)/self.H0())**2*self.rho_crit_z(z)
    # tight Sigma(R)
    fake_gcl._PS.Sigma = lambda self, R: (self._r.Sigma_tol(np.atleast_1d(R)) if np.ndim(R) else float(self._r.Sigma_tol([R])[0]))

    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    params = {k[len('cosmo_'):]: float(g[k]) for k in g.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['_growth'] = {0.0: 1.0, 0.55: 0.7486}
    params['_f'] = {0.0: float(g['par_f']), 0.55: 0.7602}
    params['_H'] = {0.0: 100.*params['h'], 0.55: 100.*params['h']*1.3386}
    cosmo = pygcl.Cosmology(pygcl.ClassParams.from_dict(params), 0, float(g['cosmo_sigma8']), g['cosmo_k'], g['cosmo_T'])

    out = {'z': 0.55, 'rho_bar_0': cosmo.rho_bar_z(0.), 'Hz': cosmo.H_z(0.55), 'H0': cosmo.H0()}
    rng = np.random.default_rng(3)
    s8 = np.concatenate([rng.uniform(0.32, 0.98, 20), [0.25, 1.05, 0.6]])       # the last three leave / touch the table
    b1 = np.concatenate([rng.uniform(1.0, 6.0, 20), [2.0, 3.0, 0.95]])
    out['sigma8_z'], out['b1'] = s8, b1
    fresh = tools.BiasToSigmaRelation(0.55, cosmo, interpolate=False)
    out['mass_fresh'] = np.array([fresh.mass(a, b) for a, b in zip(s8, b1)])
    out['sigmav_fresh'] = np.array([fresh(a, b) for a, b in zip(s8, b1)])
    print('fresh done %.0fs' % (time.time() - t0), flush=True)
    interp = tools.BiasToSigmaRelation(0.55, cosmo, interpolate=True)
    out['mass_interp'] = np.array([np.squeeze(interp.mass(a, b)) for a, b in zip(s8, b1)], dtype=float)
    out['sigmav_interp'] = np.array([np.squeeze(interp(a, b)) for a, b in zip(s8, b1)], dtype=float)
    tab = interp.interpolation_table.values
    out['table_corner_values'] = np.array([tab[0, 0], tab[0, -1], tab[-1, 0], tab[-1, -1], tab[50, 35]])
    print('interpolated done %.0fs' % (time.time() - t0), flush=True)
    norm = tools.BiasToSigmaRelation(0.55, cosmo, interpolate=False, sigmav_0=3.6, M_0=5.4903e13)
    out['sigmav_normed'] = np.array([norm(a, b) for a, b in zip(s8[:5], b1[:5])])
    out['Sigma_R'] = np.array([0.01, 0.5, 8.0, 30.0, 200.0])
    out['Sigma'] = fresh.power_lin.Sigma(out['Sigma_R'])
    path = os.path.join(HERE, 'bias_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f kB' % (os.path.getsize(path)/1e3))


if __name__ == '__main__':
    main()
