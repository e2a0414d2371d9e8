"""shared helpers for the GPU parity tests"""
import numpy as np


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


# synthetic stand-ins for the reference's Gaussian-process fits -- identical to
# tests/golden/make_galaxy_ref.py (both sides of the galaxy parity test use them)
def b2_00(b1):
    return 0.77 - 2.43*b1 + b1**2


def b2_01(b1):
    return 0.9 - 1.4*b1 + 0.45*b1**2


def stochasticity(k, sigma8_z, b1a, b1b):
    return (sigma8_z/0.8)*(120.*b1a*b1b - 400.)*(1.0 + 0.3*np.log(k/0.1) - 0.05*np.log(k/0.1)**2)
