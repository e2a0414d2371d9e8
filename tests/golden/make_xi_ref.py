"""Generate ``tests/golden/xi_ref.npz``: the reference's ``ExtrapolatedPowerSpectrum`` / ``SmoothedXiMultipoles``
(``pyRSD/rsd/power_extrapolator.py``, ``pyRSD/rsd/correlation.py``, imported UNMODIFIED from /root/reference) on top of the
reference ``GalaxySpectrum`` of make_galaxy_ref.py (same stand-ins, tight tables, converged Zel'dovich state).

The reference classes call ``model.Pgal(k, mu)``, which the current ``GalaxySpectrum`` no longer has: the model handed to them
is a thin wrapper whose ``Pgal`` is ``GalaxySpectrum.power``.

Run in the build container only:  python tests/golden/make_xi_ref.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_galaxy_ref as G  # noqa: E402

KW = dict(k_lo=2e-3, k_hi=0.4)


class WithPgal(object):
    def __init__(self, model):
        self.m = model
        self.kmin, self.kmax = model.kmin, model.kmax

    def Pgal(self, k, mu):
        return np.asarray(self.m.power(k, mu))


def main():
    t0 = time.time()
    G.ZEL_SOURCE = 'pi'
    cosmo_mod = G.install_stubs()
    G.install_tight()
    print('dense fixture seeded:', G.seed_from_dense(), flush=True)
    from pyRSD import pygcl
    from pyRSD.rsd import ExtrapolatedPowerSpectrum
    from pyRSD.rsd.correlation import SmoothedXiMultipoles

    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    params = {k[len('cosmo_'):]: float(g[k]) for k in g.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['_growth'] = {0.0: 1.0, 0.55: 0.7486}
    params['_f'] = {0.0: float(g['par_f']), 0.55: 0.7602}
    params['_H'] = {0.0: 100.*params['h'], 0.55: 100.*params['h']*1.3386}
    cosmo = pygcl.Cosmology(pygcl.ClassParams.from_dict(params), 0, float(g['cosmo_sigma8']), g['cosmo_k'], g['cosmo_T'])
    model = G.build_model(cosmo_mod, cosmo, 0.55, interpolate=False)
    model.update(sigma8_z=0.61, f=0.77)
    w = WithPgal(model)
    ex = ExtrapolatedPowerSpectrum(w, **KW)
    out = {}
    k = np.array([2e-5, 3e-4, 1.5e-3, 2.5e-3, 0.02, 0.11, 0.35, 0.45, 0.9, 3.0, 9.0])
    mu = np.linspace(0., 1., 6)
    out['k'], out['mu'] = k, mu
    out['extrap'] = np.asarray(ex(k, mu))
    out['alpha_lo'] = np.asarray(ex.low_k_params(mu))
    out['alpha_hi'], out['beta_hi'] = [np.asarray(x) for x in ex.high_k_params(mu)]
    ks = np.logspace(-5, 1, 1000)
    out['poles'] = np.asarray(ex.to_poles(ks, [0, 2, 4]))
    print('poles done %.0fs' % (time.time() - t0), flush=True)
    r = np.logspace(0.3, 2.2, 40)
    out['r'] = r
    out['xi_R0'] = np.asarray(SmoothedXiMultipoles(w, r, [0, 2, 4], R=0., **KW))
    out['xi_R3'] = np.asarray(SmoothedXiMultipoles(w, r, [0, 2], R=3., **KW))
    print('xi done %.0fs' % (time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, 'xi_ref.npz'), **out)
    G.save_cache()
    print('wrote xi_ref.npz')


if __name__ == '__main__':
    main()
