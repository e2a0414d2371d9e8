"""Generate ``tests/golden/qso_ref.npz``: the reference's ``QuasarSpectrum`` (``pyRSD/rsd/power/qso/power_qso.py``, imported
UNMODIFIED from /root/reference) over the reference C++ ``LinearPS`` (oracle/_ref): Kaiser terms with the scale-dependent
bias of primordial non-Gaussianity, Finger-of-God damping, Alcock-Paczynski distortion.

Run in the build container only:  python tests/golden/make_qso_ref.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_galaxy_ref as G  # noqa: E402

CASES = {
    'gauss_fnl0': (dict(), dict(b1=2.3, sigma_fog=3.5, alpha_par=1.02, alpha_perp=0.98, sigma8_z=0.61, f=0.78)),
    'gauss_fnl': (dict(), dict(b1=2.3, f_nl=25., p=1.6, sigma_fog=3.5, N=120., sigma8_z=0.61, f=0.78)),
    'lorentz_fnl': (dict(fog_model='modified_lorentzian', max_mu=2), dict(b1=1.9, f_nl=-40., p=1.0, sigma_fog=5.0, alpha_par=0.97,
                                                                           alpha_perp=1.03, sigma8_z=0.55, f=0.7)),
}


def main():
    cosmo_mod = G.install_stubs()
    from pyRSD import pygcl
    from pyRSD.rsd import QuasarSpectrum

    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    params = {k[len('cosmo_'):]: float(g[k]) for k in g.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['_growth'] = {0.0: 1.0, 0.55: 0.7486}
    params['_f'] = {0.0: float(g['par_f']), 0.55: 0.7602}
    params['_H'] = {0.0: 100.*params['h'], 0.55: 100.*params['h']*1.3386}
    cosmo = pygcl.Cosmology(pygcl.ClassParams.from_dict(params), 0, float(g['cosmo_sigma8']), g['cosmo_k'], g['cosmo_T'])
    out = {'k': np.logspace(np.log10(1.2e-3), np.log10(0.4), 60), 'mu': np.linspace(0., 1., 9)}
    for name, (kw, upd) in CASES.items():
        m = QuasarSpectrum(params=cosmo_mod.Cosmology(cosmo), z=0.55, **kw)
        m.update(**upd)
        k, mu = out['k'], out['mu']
        out[name + '_power'] = np.asarray(m.power(k, mu))
        out[name + '_moments'] = np.array([m.P_mu0(k), m.P_mu2(k), m.P_mu4(k), m.P_mu6(k)])
        out[name + '_btot'] = np.asarray(m.btot(k))
        out[name + '_alpha_png'] = np.asarray(m.alpha_png(k))
        print(name, 'done', out[name + '_power'].shape, float(out[name + '_btot'][0]))
    out['Omega0_m'] = cosmo['Omega0_m']
    np.savez_compressed(os.path.join(HERE, 'qso_ref.npz'), **out)
    print('wrote qso_ref.npz')


if __name__ == '__main__':
    main()
