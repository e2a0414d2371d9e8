"""Synthetic workload of the benchmark and the smooth stand-ins for the reference's Gaussian-process
fits that the parity fixtures use on BOTH sides (tests/golden/make_galaxy_ref.py has the same three functions).

Nothing here is physics of the reference: it only manufactures distinct, plausible inputs so that every table of
every cosmology genuinely has to be rebuilt."""
import os

import numpy as np

SEED = 20261019
NK_OUT, NMU_OUT = 1000, 40


def b2_00(b1):
    return 0.77 - 2.43*b1 + b1**2


def b2_01(b1):
    return 0.9 - 1.4*b1 + 0.45*b1**2


def stochasticity(k, sigma8_z, b1a, b1b):
    """smooth stand-in for the GP stochasticity Lambda(k; sigma8, b1, b1_bar)"""
    return (sigma8_z/0.8)*(120.*b1a*b1b - 400.)*(1.0 + 0.3*np.log(k/0.1) - 0.05*np.log(k/0.1)**2)


def vel_disp(b1, sigma8_z):
    """smooth stand-in for VelocityDispersionFits (halo sigma_v in Mpc/h)"""
    return 3.2 + 0.9*(b1 - 2.0) + 2.0*(sigma8_z - 0.6)


def Pmu2_corr(f, sigma8_z, b1, k):
    """smooth stand-in for Pmu2ResidualCorrection"""
    return 900.*b1*(f/0.75)*(sigma8_z/0.6)**2*(k/0.1)/(1.0 + (k/0.15)**2)


def Pmu4_corr(f, sigma8_z, b1, k):
    """smooth stand-in for Pmu4ResidualCorrection"""
    return -350.*b1*(f/0.75)**2*(sigma8_z/0.6)**2*(k/0.1)**2/(1.0 + (k/0.2)**3)


def loaded_terms(k):
    """smooth stand-ins for user-loaded dark-matter terms (SimLoaderMixin._load): name -> P(k)"""
    shape = 2.2e4*(k/0.02)/(1.0 + (k/0.02)**2.3)
    return {'Pdv': -0.78*shape*(1.0 - 0.8*k), 'P00_mu0': shape*(1.0 + 1.5*k), 'P01_mu2': 1.5*shape*(1.0 + 0.6*k),
            'P11_mu4': 0.55*shape*(1.0 - 1.2*k)}


def base_cosmology():
    """the reference's shipped Planck15 transfer function T(k) (500 samples, from docs/data/galaxy_power.npy) and the
    constants of its Eisenstein-Hu extrapolation; delta_H is the sigma8 = 0.8159 normalisation (data/planck15_transfer.npz,
    written from tests/golden/golden_pt.npz)"""
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'planck15_transfer.npz'))
    out = {k: float(d[k]) for k in d.files if k not in ('k', 'T')}
    out['k'], out['T'] = np.asarray(d['k'], dtype=float), np.asarray(d['T'], dtype=float)
    return out


def tilts(n, rank=0):
    """the shape tilts delta ~ N(0, 0.02) of rank `rank`'s batch (first draw of the batch's generator)"""
    return np.random.default_rng(SEED + 1000*rank).normal(0, 0.02, n)


def synthetic_batch(n, rank=0, nk_out=NK_OUT, nmu_out=NMU_OUT):
    """n synthetic cosmologies + walker vectors: base = the reference's shipped Planck15 transfer function; variant
    i: sigma8_z in U[0.45, 0.95], f in U[0.4, 1.0], b1_cA in U[1.2, 2.5] and a shape tilt T_i(k) (k/0.05)^(delta/2),
    delta ~ N(0, 0.02), so that every table genuinely has to be rebuilt."""
    from . import rsd
    c = base_cosmology()
    rng = np.random.default_rng(SEED + 1000*rank)
    k_i, T_i = c['k'], c['T']
    delta = rng.normal(0, 0.02, n)
    T = T_i[None, :]*(k_i[None, :]/0.05)**(delta[:, None]/2)
    k = np.repeat(k_i[None, :], n, axis=0)
    pars = np.zeros((n, 8))
    pars[:, 0] = c['n_s']
    pars[:, 1] = c['delta_H']**2        # re-normalisation to sigma8 is per-cosmology algebra done at assembly
    pars[:, 2], pars[:, 3], pars[:, 4], pars[:, 5] = c['h'], c['Omega0_m'], c['Omega0_b'], c['Tcmb']
    s8z = rng.uniform(0.45, 0.95, n)
    f = rng.uniform(0.4, 1.0, n)
    b1cA = rng.uniform(1.2, 2.5, n)
    kmodel = np.logspace(np.log10(1e-3), np.log10(0.5), 200)
    ks = rsd.stochasticity_grid(kmodel)
    wpar = np.zeros((n, rsd.NPAR))
    stoch = np.zeros((n, rsd.NPAIR, rsd.NSTOCH))
    for i in range(n):
        b1 = [b1cA[i], 1.5*b1cA[i], 1.4*b1cA[i], 1.9*b1cA[i]]
        b00, b01 = [b2_00(b) for b in b1], [b2_01(b) for b in b1]
        wpar[i] = rsd.pack_walker(s8z[i], f[i], 1.0, 1.0, 1.0, b1, 0.10, 0.08, 0.40, 1.0, 4.2, 6.0, 3e4, 9e4, 0., 0., None,
                                  0.7486, c['sigma8'], [b00, b00, b00, b00, b01, b01])
        for p, (a, b) in enumerate(rsd.PAIRS):
            lo, hi = sorted([b1[a], b1[b]])
            stoch[i, p] = stochasticity(ks, s8z[i], lo, hi)
    kout = np.logspace(-2.95, np.log10(0.49), nk_out)
    mu = np.linspace(0., 1., nmu_out)
    kk, mm = [a.ravel() for a in np.meshgrid(kout, mu, indexing='ij')]
    return dict(k=k, T=T, pars=pars, wpar=wpar, stoch=stoch, kmodel=kmodel, kk=np.ascontiguousarray(kk),
                mm=np.ascontiguousarray(mm), delta=delta, kout=kout, mu=mu)
