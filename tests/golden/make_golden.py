"""Generate ``tests/golden/golden_pt.npz`` from the reference's own shipped model.

Source: ``/root/reference/docs/data/galaxy_power.npy`` -- a pickled, fully
initialised ``pyRSD.rsd.GalaxySpectrum`` (model version 0.3.1, Planck15, z=0)
that the reference's documentation loads.  Everything stored there was computed
by the reference's C++ (``pyRSD/_gcl``) at its *default* ``epsrel``.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The output is committed; tests never touch /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from unpickle_ref import load_reference_model  # noqa: E402

REF = '/root/reference/docs/data/galaxy_power.npy'


def main():
    m = load_reference_model(REF).item()
    st = m._state
    c = st['_cache']
    out = {}

    # --- cosmology: pygcl.Cosmology.__getstate__ (pyRSD/pygcl.py:60-63)
    params, tf, sigma8, k_i, T_i = c['cosmo']._state['args']
    out['cosmo_k'] = np.asarray(k_i, dtype='f8')
    out['cosmo_T'] = np.asarray(T_i, dtype='f8')
    out['cosmo_sigma8'] = float(sigma8)
    out['cosmo_tf'] = int(tf)
    for key in ('Omega_b', 'Omega_cdm', 'h', 'n_s', 'm_ncdm', 'N_ur', 'N_ncdm', 'Omega_g'):
        out['cosmo_' + key] = float(params[key])

    # --- one-loop tables (OneLoopPS.GetOneLoopPower, pygcl.py:198-201)
    for name in ('_Pdd_0', '_Pdv_0', '_Pvv_0', '_P22bar_0'):
        args = c[name]._state['args']
        out['oneloop' + name] = np.asarray(args[2], dtype='f8')
        out['oneloop' + name + '_epsrel'] = float(args[1])

    # --- unnormalised PT tables on k_interp (rsd/pt_integrals.py)
    for name in sorted(c.keys()):
        if not name.startswith('_unnormalized_'):
            continue
        sp = c[name]._state['spline']
        short = name[len('_unnormalized_'):]
        if isinstance(sp, list):
            out['pt_' + short] = np.array([np.asarray(s._state['y'], 'f8') for s in sp])
            x = sp[0]._state['x']
        else:
            out['pt_' + short] = np.asarray(sp._state['y'], 'f8')
            x = sp._state['x']
        if 'k_interp' not in out:
            out['k_interp'] = np.asarray(x, 'f8')
        else:
            assert np.array_equal(out['k_interp'], np.asarray(x, 'f8'))

    out['sigma_lin_unnormed'] = float(c['_sigma_lin_unnormed'])
    out['velocity_kurtosis_unnormed'] = float(c['_unnormed_velocity_kurtosis'])

    # --- Zel'dovich base state (pygcl.py:251-253)
    zargs = c['hzpt']._state['_base_zeldovich']._state['args']
    out['zel_approx_lowk'] = bool(zargs[1])
    out['zel_sigma8_z'] = float(zargs[2])
    out['zel_k0_low'] = float(zargs[3])
    out['zel_sigmasq'] = float(zargs[4])
    out['zel_X0'] = np.asarray(zargs[5], 'f8')
    out['zel_XX'] = np.asarray(zargs[6], 'f8')
    out['zel_YY'] = np.asarray(zargs[7], 'f8')

    # --- HZPT (sigma8, k) tables (rsd/hzpt/interpolated.py:26-72); keep a subset
    #     of sigma8 rows to stay small (every 9th row + the ends => 13 rows)
    table = c['hzpt']._state['table']
    rows = sorted(set(list(range(0, 100, 9)) + [99]))
    out['hzpt_rows'] = np.array(rows)
    for name in ('P00', 'P01', 'P11', 'Phm'):
        t = dict.__getitem__(table, name)._state
        s8, kk = t['grid']
        if 'hzpt_sigma8' not in out:
            out['hzpt_sigma8'] = np.asarray(s8, 'f8')
            out['hzpt_k'] = np.asarray(kk, 'f8')
        out['hzpt_' + name] = np.asarray(t['values'], 'f8')[rows]
    for name in ('_P00_driver', '_P01_driver', '_P11_driver', '_Phm_driver'):
        d = c['hzpt']._state[name]._state
        for k, v in d['_cache'].items():
            out['hzpt%s_cache_%s' % (name, k)] = float(v)

    # --- model-level 200-pt splines stored in the cache
    out['model_k'] = np.asarray(c['k'], 'f8')
    for name in ('Pdd', 'Pdv', 'Pvv'):
        sp = c[name]._state['spline']._state
        assert np.array_equal(sp['x'], out['model_k'])
        out['model_' + name] = np.asarray(sp['y'], 'f8')

    # --- scalar model parameters
    for key, val in st.items():
        if key.startswith('__') and isinstance(val, (int, float, bool, np.floating)) \
                and not isinstance(val, str):
            out['par' + key] = float(val)
    out['par_f'] = float(st['__f'])
    out['par_sigma8_z'] = float(st['__sigma8_z'])

    path = os.path.join(HERE, 'golden_pt.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f kB' % (os.path.getsize(path) / 1e3), len(out), 'arrays')


if __name__ == '__main__':
    main()
