"""shared helpers for the GPU parity tests"""
import json
import os

import numpy as np


def record_parity(section, tag, value):
    """worst errors of a GPU parity test -> gpurun_out/r02_parity.json[section][tag] (rendered by profiles/make_parity_md.py)"""
    try:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'r02_parity.json')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur.setdefault(section, {})[tag] = value
        json.dump(cur, open(path, 'w'), indent=1, sort_keys=True)
    except OSError:
        pass


def golden_pygcl_cosmology(golden, tight=None):
    """pyrsd_b200.pygcl.Cosmology for the reference's shipped Planck15 transfer function"""
    from pyrsd_b200 import pygcl
    params = {k[len('cosmo_'):]: float(golden[k]) for k in golden.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['T_cmb'] = 2.7255
    params['growth'] = {0.0: 1.0, 0.55: 0.7486}
    params['f'] = {0.0: float(golden['par_f']), 0.55: 0.7602}
    params['H'] = {0.0: 100.*params['h'], 0.55: 100.*params['h']*1.3386}
    return pygcl.Cosmology(params, 0, float(golden['cosmo_sigma8']), golden['cosmo_k'], golden['cosmo_T'])


def relerr_floor(mine, ref, floor):
    """|mine - ref| / max(|ref|, floor): the parity metric of SURVEY 8c for I / K integrals"""
    return np.abs(mine - ref)/np.maximum(np.abs(ref), floor)


def ml_scale_floor(ref_parts, plin):
    """scale of one ImnOneLoop integral as the model reads it.  ``ref_parts`` [3][nk] = its linear / cross / one-loop
    parts on the tabulation grid.  The reference only ever uses norm^2 lin + norm^3 cross + norm^4 oneloop, splined in
    k (rsd/pt_integrals.py:31-40, ``interpolated_function``), and the f32 / f33 parts change sign between tabulated
    points (k ~ 0.5, 0.86-0.9) where |value| drops to 1/15 - 1/500 of its neighbours and of the other parts: the
    floor of the parity metric is therefore max(P_L(k), |part| over the three parts at k_{i-2} .. k_{i+2})."""
    a = np.max(np.abs(np.asarray(ref_parts)), axis=0)
    n = len(a)
    loc = np.array([a[max(i - 2, 0):i + 3].max() for i in range(n)])
    return np.maximum(loc, plin)


# synthetic stand-ins for the reference's Gaussian-process fits -- identical to tests/golden/make_galaxy_ref.py
# (both sides of the galaxy parity test use them); they live with the benchmark's workload generator
from pyrsd_b200.synthetic import b2_00, b2_01, stochasticity  # noqa: E402,F401
