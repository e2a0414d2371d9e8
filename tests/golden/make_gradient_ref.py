"""Generate ``tests/golden/gradient_ref.npz``: the reference's ANALYTIC derivatives of GalaxySpectrum.power
(pyRSD/rsd/power/gal/derivatives/*.py, unmodified, imported from /root/reference) on top of the same tight-tolerance
model as tests/golden/make_galaxy_ref.py, plus the reference's central finite difference
(PkmuGradient.__call__, rsd/power/gradient.py:178-203, epsilon = 1e-4) for parameters without an analytic form.

Run in the build container only (reuses /tmp/galaxy_ref_cache.npz when present):
    python tests/golden/make_gradient_ref.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_galaxy_ref as G  # noqa: E402

ANALYTIC = ['fs', 'fcB', 'fsB', 'sigma_c', 'NcBs', 'NsBsB', 'alpha_par', 'alpha_perp']
ANALYTIC_SO = ANALYTIC + ['f_so', 'sigma_so']
# fsB: the reference's analytic form (derivatives/fsB.py:14-15) has the satellite-satellite part only and omits the
# dependence of the central-satellite term on fsB (compare fcB.py:16-17) -- its finite difference is the pin
NUMERICAL = ['sigma8_z', 'f', 'b1_cA', 'b1_sB', 'sigma_sA', 'sigma_sB', 'fsB']


def main():
    cosmo_mod = G.install_stubs()
    G.install_tight()
    from pyRSD import pygcl
    from pyRSD.rsd.power.gal.derivatives import PgalDerivative
    registry = PgalDerivative.registry()

    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    params = {k[len('cosmo_'):]: float(g[k]) for k in g.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['_growth'] = {0.0: 1.0, 0.55: 0.7486}
    params['_f'] = {0.0: float(g['par_f']), 0.55: 0.7602}
    params['_H'] = {0.0: 100.*params['h'], 0.55: 100.*params['h']*1.3386}
    cosmo = pygcl.Cosmology(pygcl.ClassParams.from_dict(params), 0, float(g['cosmo_sigma8']), g['cosmo_k'], g['cosmo_T'])

    k = np.logspace(np.log10(0.012), np.log10(0.38), 16)
    mu = np.linspace(0.05, 0.95, 7)
    kk, mm = [x.ravel() for x in np.meshgrid(k, mu, indexing='ij')]
    out = {'k': kk, 'mu': mm}
    cases = {
        'ap': (dict(interpolate=False),
               dict(alpha_par=1.03, alpha_perp=0.98, b1_cA=1.9, b1_cB=2.7, b1_sA=2.4, b1_sB=3.3, fs=0.12, fcB=0.09, fsB=0.35,
                    sigma_c=1.1, sigma_sA=3.9, sigma_sB=5.5, NcBs=4.2e4, NsBsB=8.1e4, sigma8_z=0.61, f=0.78), ANALYTIC),
        'so': (dict(interpolate=False, use_so_correction=True, fog_model='gaussian', max_mu=6, use_tidal_bias=True),
               dict(f_so=0.04, sigma_so=4.0, alpha_par=0.97, alpha_perp=1.02), ANALYTIC_SO),
    }
    for tag, (kw, upd, names) in cases.items():
        model = G.build_model(cosmo_mod, cosmo, 0.55, **kw)
        model.update(**upd)
        out[tag + '_P'] = np.asarray(model.power(kk, mm))
        for name in names:
            out['%s_d_%s' % (tag, name)] = np.asarray(registry[name].eval(model, None, kk, mm))*np.ones(len(kk))
            print(tag, name, 'analytic', flush=True)
        if tag == 'ap':
            eps = 1e-4
            for name in NUMERICAL:
                x0 = float(getattr(model, name))
                model.update(**{name: x0 + eps})
                Pp = np.asarray(model.power(kk, mm))
                model.update(**{name: x0 - eps})
                Pm = np.asarray(model.power(kk, mm))
                model.update(**{name: x0})
                out['%s_n_%s' % (tag, name)] = (Pp - Pm)/(2*eps)
                print(tag, name, 'numerical', flush=True)
        G.save_cache()
    path = os.path.join(HERE, 'gradient_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f kB' % (os.path.getsize(path)/1e3))


if __name__ == '__main__':
    main()
