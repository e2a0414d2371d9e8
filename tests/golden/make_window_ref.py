"""Generate ``tests/golden/window_ref.npz``: gridded multipoles / wedges and window-convolved multipoles computed
with the reference's OWN Python, imported from /root/reference:

  * ``pyRSD.extern.mcfit`` (P2xi, xi2P)                      -- unmodified
  * ``pyRSD.rsd.window.WindowConvolution``                   -- unmodified
  * ``pyRSD.rsd.transfers.grid.GriddedWedgeTransfer.average`` / ``GriddedMultipoleTransfer.legendre_weights``
    -- unmodified methods; their ``__call__`` wrappers need xarray (absent here), so the three lines around them
    (grid.py:300-311) and the zero-padding / FFT / spline glue of ``WindowFunctionTransfer.__call__``
    (transfers/window.py:78-124) are restated below, line references given.

The same inputs go through oracle/transfer_restate.py and must agree to rounding (checked here and again in
tests/test_transfers_oracle.py).  Run in the build container only:  python tests/golden/make_window_ref.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_galaxy_ref import install_stubs  # noqa: E402
from oracle import transfer_restate as T  # noqa: E402
from window_inputs import synthetic_power, synthetic_window  # noqa: E402


def main():
    install_stubs()
    from pyRSD.extern import mcfit
    from pyRSD.rsd.window import WindowConvolution
    from pyRSD.rsd.transfers import PkmuGrid
    from pyRSD.rsd.transfers.grid import GriddedMultipoleTransfer, GriddedWedgeTransfer
    from scipy.interpolate import InterpolatedUnivariateSpline as spline

    out = {}
    window = synthetic_window()
    out['window'] = window
    ells = [0, 2, 4]
    for tag, (kmin, kmax, Nk, Nmu) in {'a': (1e-4, 0.7, 1024, 40), 'b': (1e-3, 0.5, 300, 40)}.items():
        # WindowFunctionTransfer.__init__ (transfers/window.py:33-49)
        k = np.logspace(np.log10(kmin), np.log10(kmax), Nk)
        mu_edges = np.linspace(0., 1., Nmu + 1)
        mu = 0.5*(mu_edges[1:] + mu_edges[:-1])
        grid_k, grid_mu = np.meshgrid(k, mu, indexing='ij')
        grid = PkmuGrid([k, mu], grid_k, grid_mu, np.ones_like(grid_k))
        tr = GriddedMultipoleTransfer(grid, ells, kmin=kmin, kmax=kmax)
        power = synthetic_power(grid_k, grid_mu)
        # GriddedMultipoleTransfer.__call__ (grid.py:306-311)
        tobin = tr.legendre_weights*power
        Pell0 = np.asarray([np.squeeze(tr.average(d, tr.grid.modes)) for d in tobin]).T
        convolver = WindowConvolution(window[:, 0], window[:, 1:], max_ellprime=4, max_ell=max(ells))
        # WindowFunctionTransfer.__call__ (transfers/window.py:84-124), extrap=False
        dk = np.diff(np.log10(k))[0]
        newk = 10**(np.arange(np.log10(k.max()) + dk, 2 + 0.5*dk, dk))
        newk = np.concatenate([k, newk])
        Pell = np.zeros((len(newk), len(ells)))
        Pell[:Nk] = Pell0
        xi = np.empty((len(newk), len(ells)), order='F')
        for i, ell in enumerate(ells):
            P2xi = mcfit.P2xi(newk, l=ell)
            rr, xi[:, i] = P2xi(Pell[:, i], extrap=False)
        xi_conv = convolver(ells, rr, xi, order='F')
        Pconv = np.empty((len(newk), len(ells)), order='F')
        Pdry = np.empty((len(newk), len(ells)), order='F')
        for i, ell in enumerate(ells):
            xi2P = mcfit.xi2P(rr, l=ell)
            kk, Pconv[:, i] = xi2P(xi_conv[:, i], extrap=False)
            kk, Pdry[:, i] = xi2P(xi[:, i], extrap=False)
        k_out = np.linspace(0.01, 0.4, 40) if tag == 'a' else np.linspace(0.02, 0.3, 15)
        Pout = np.array([spline(newk, Pconv[:, i])(k_out) for i in range(len(ells))]).T

        # the restatement must agree with the reference's own code
        R = T.window_transfer(power, k, mu, window, ells, 4, k_out)
        for name, ref in (('Pell0', Pell0), ('xi', xi), ('xi_conv', xi_conv), ('Pconv', Pconv), ('Pout', Pout)):
            err = np.max(np.abs(R[name] - ref))/np.max(np.abs(ref))
            print(tag, name, 'restatement vs reference: %.2e' % err)
            assert err < 2e-11, (name, err)      # 1-ulp differences of Pell0, amplified by the oscillatory transforms

        # k, mu and power are reproducible from (kmin, kmax, Nk, Nmu) and window_inputs.py: not stored
        out[tag + '_grid'] = np.array([kmin, kmax, Nk, Nmu])
        for name, val in (('Pell0', Pell0), ('newk', newk), ('rr', rr), ('xi', xi),
                          ('xi_conv', xi_conv), ('Pconv', Pconv), ('Pdry', Pdry), ('k_out', k_out), ('Pout', Pout),
                          ('N', P2xi.N)):
            out[tag + '_' + name] = np.ascontiguousarray(val)

    # a gridded measurement with uneven mode counts and empty cells: wedges and multipoles (grid.py:165-187, 300-311)
    rng = np.random.default_rng(20261020)
    Nk, Nmu = 60, 50
    k_edges = np.linspace(0.005, 0.305, Nk + 1)
    k_cen = 0.5*(k_edges[1:] + k_edges[:-1])
    mu_edges = np.linspace(0., 1., Nmu + 1)
    mu_cen = 0.5*(mu_edges[1:] + mu_edges[:-1])
    gk = k_cen[:, None] + 0.3*np.diff(k_edges)[0]*(rng.random((Nk, Nmu)) - 0.5)
    gmu = mu_cen[None, :] + 0.3*np.diff(mu_edges)[0]*(rng.random((Nk, Nmu)) - 0.5)
    modes = np.floor(4e3*(gk/0.1)**2*rng.random((Nk, Nmu)))
    modes[:6][rng.random((6, Nmu)) < 0.5] = 0.                   # low-k cells without modes
    modes[3, 7] = np.nan
    grid = PkmuGrid([k_cen, mu_cen], gk, gmu, modes)
    power = synthetic_power(gk, gmu)
    bounds = [(0., 0.2), (0.2, 0.4), (0.4, 0.6), (0.6, 0.8), (0.8, 1.0)]
    wt = GriddedWedgeTransfer(grid, list(bounds), kmin=0.01, kmax=0.3)
    wedges = wt.average(power, grid.modes)
    mt = GriddedMultipoleTransfer(grid, [0, 2, 4], kmin=0.01, kmax=0.3)
    poles = np.asarray([np.squeeze(mt.average(d, grid.modes)) for d in mt.legendre_weights*power]).T
    Rw = T.gridded_wedges(power, gmu, modes, bounds)
    Rp = T.gridded_poles(power, gmu, modes, [0, 2, 4])
    assert np.array_equal(np.isnan(Rw), np.isnan(wedges)) and np.array_equal(np.isnan(Rp), np.isnan(poles))
    for name, a, b in (('wedges', Rw, wedges), ('poles', Rp, poles)):
        err = np.nanmax(np.abs(a - b))/np.nanmax(np.abs(b))
        print('g', name, 'restatement vs reference: %.2e' % err)
        assert err < 1e-13, (name, err)
    for name, val in (('k_cen', k_cen), ('mu_cen', mu_cen), ('k', gk), ('mu', gmu), ('modes', modes), ('power', power),
                      ('wedges', wedges), ('poles', poles), ('bounds', np.array(bounds))):
        out['g_' + name] = val

    # the older route, rsd/window.py:214-307 (convolve_multipoles), UNMODIFIED, both branches; its transforms go through the
    # reference C++ (oracle/_ref: ComputeXiLM TRAPZ, ComputeXiLM_fftlog)
    from pyRSD.rsd.window import convolve_multipoles
    convolver = WindowConvolution(window[:, 0], window[:, 1:], max_ellprime=4, max_ell=4)
    kc = np.logspace(-3, np.log10(0.6), 120)                    # < 500 points: the FITPACK refinement branch
    mu_g = np.linspace(0.0125, 0.9875, 40)
    Pg = synthetic_power(kc[:, None]*np.ones((1, 40)), np.ones((120, 1))*mu_g[None, :])
    from scipy.special import legendre
    Pell_c = np.array([(2*ell + 1)*np.mean(Pg*legendre(ell)(mu_g)[None, :], axis=1) for ell in ells]).T
    out['c_k'], out['c_Pell'] = kc, Pell_c
    out['c_legacy'] = np.asarray(convolve_multipoles(kc, Pell_c, ells, convolver, legacy=True)[1])
    out['c_legacy_dry'] = np.asarray(convolve_multipoles(kc, Pell_c, ells, convolver, dry_run=True, legacy=True)[1])
    kf = np.logspace(-4, 1, 1024)
    Pgf = synthetic_power(kf[:, None]*np.ones((1, 40)), np.ones((1024, 1))*mu_g[None, :])
    Pell_f = np.array([(2*ell + 1)*np.mean(Pgf*legendre(ell)(mu_g)[None, :], axis=1) for ell in ells]).T
    kk, Pc = convolve_multipoles(kf, Pell_f, ells, convolver, qbias=0.7, legacy=False)
    out['f_k'], out['f_Pell'], out['f_kout'], out['f_conv'] = kf, Pell_f, np.asarray(kk), np.asarray(Pc)
    print('convolve_multipoles: legacy', out['c_legacy'].shape, 'fftlog', out['f_conv'].shape)

    path = os.path.join(HERE, 'window_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f kB' % (os.path.getsize(path)/1e3))


if __name__ == '__main__':
    main()
