"""CPU: the NumPy restatement of the reference's transfer functions (oracle/transfer_restate.py) against the golden
vectors produced by the reference's own mcfit / WindowConvolution / grid averaging
(tests/golden/window_ref.npz, tests/golden/make_window_ref.py), and the host-side set-up of pyrsd_b200.transfers."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from window_inputs import synthetic_power, synthetic_window  # noqa: E402
from oracle import transfer_restate as T  # noqa: E402

REF = np.load(os.path.join(HERE, 'golden', 'window_ref.npz'))


def grid_of(tag):
    kmin, kmax, Nk, Nmu = REF[tag + '_grid']
    k = np.logspace(np.log10(kmin), np.log10(kmax), int(Nk))
    e = np.linspace(0., 1., int(Nmu) + 1)
    return k, 0.5*(e[1:] + e[:-1])


def relmax(a, b):
    return np.nanmax(np.abs(a - b))/np.nanmax(np.abs(b))


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_window_restatement_matches_reference(tag):
    k, mu = grid_of(tag)
    power = synthetic_power(k[:, None], mu[None, :])
    R = T.window_transfer(power, k, mu, synthetic_window(), [0, 2, 4], 4, REF[tag + '_k_out'])
    assert np.array_equal(R['newk'], REF[tag + '_newk'])
    assert relmax(R['rr'], REF[tag + '_rr']) < 1e-15
    # tolerance: 1-ulp differences of the multipoles are amplified ~1e3 by the oscillatory transforms
    for name in ('Pell0', 'xi', 'xi_conv', 'Pconv', 'Pout'):
        assert relmax(R[name], REF[tag + '_' + name]) < 2e-11, name
    D = T.window_transfer(power, k, mu, synthetic_window(), [0, 2, 4], 4, dry_run=True)
    assert relmax(D['Pconv'], REF[tag + '_Pdry']) < 2e-11
    # P -> xi -> P returns the input where it is well inside the padded range
    sel = (k > 0.03) & (k < 0.3)
    assert np.max(np.abs(D['Pconv'][:len(k)][sel]/R['Pell0'][sel] - 1.)[:, 0]) < 1e-2


def test_gridded_restatement_matches_reference():
    bounds = [tuple(b) for b in REF['g_bounds']]
    w = T.gridded_wedges(REF['g_power'], REF['g_mu'], REF['g_modes'], bounds)
    p = T.gridded_poles(REF['g_power'], REF['g_mu'], REF['g_modes'], [0, 2, 4])
    assert np.array_equal(np.isnan(w), np.isnan(REF['g_wedges']))
    assert relmax(w, REF['g_wedges']) < 1e-14
    assert relmax(p, REF['g_poles']) < 1e-13


def test_coefficients_match_restatement():
    from pyrsd_b200.transfers import get_coefficients
    for ell in (0, 2, 4, 6):
        for ellp in (0, 2, 4, 6):
            q1, c1 = get_coefficients(ell, ellp)
            q2, c2 = T.get_coefficients(ell, ellp)
            assert q1 == q2
            assert np.allclose(c1, c2, rtol=1e-14, atol=0)
    # known values (Wilson et al. 2017, eq. 19-21): monopole kernel = Q_0, Q_2/5, Q_4/9
    assert get_coefficients(0, 2) == ([2], [0.2])
    assert np.allclose(get_coefficients(0, 4)[1], [1./9])
    q, c = get_coefficients(2, 2)
    assert q == [0, 2, 4] and np.allclose(c, [1., 2./7, 2./7])


def test_host_weights_reproduce_reference_average():
    """the folded weight array of pyrsd_b200.transfers, contracted on the host, equals the reference's average"""
    from pyrsd_b200 import transfers as TR
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        grid = TR.PkmuGrid([REF['g_k_cen'], REF['g_mu_cen']], REF['g_k'], REF['g_mu'], REF['g_modes'])
    bounds = [tuple(b) for b in REF['g_bounds']]
    wt = TR.GriddedWedgeTransfer.__new__(TR.GriddedWedgeTransfer)
    TR.GriddedWedgeTransfer.__init__(wt, grid, list(bounds), kmin=0.01, kmax=0.3, context=object())
    w = wt._weights()
    got = np.einsum('bkm,km->kb', w, np.where(grid.notnull, REF['g_power'], 0.))
    ref = REF['g_wedges']
    ok = ~np.isnan(ref)
    assert np.array_equal(wt._empty.T, ~ok)
    assert np.max(np.abs(got[ok] - ref[ok]))/np.max(np.abs(ref[ok])) < 1e-14
    mt = TR.GriddedMultipoleTransfer.__new__(TR.GriddedMultipoleTransfer)
    TR.GriddedMultipoleTransfer.__init__(mt, grid, [0, 2, 4], kmin=0.01, kmax=0.3, context=object())
    got = np.einsum('bkm,km->kb', mt._weights(), np.where(grid.notnull, REF['g_power'], 0.))
    ok = ~np.isnan(REF['g_poles'])
    assert np.max(np.abs(got[ok] - REF['g_poles'][ok]))/np.max(np.abs(REF['g_poles'][ok])) < 1e-13


def test_window_kernel_matrix_matches_restatement():
    from pyrsd_b200.transfers import WindowConvolution
    win = synthetic_window()
    conv = WindowConvolution(win[:, 0], win[:, 1:], max_ellprime=4, max_ell=4)
    r = REF['a_rr']
    K = conv.kernel_matrix([0, 2, 4], r)
    ref = T.window_kernels(win[:, 0], win[:, 1:], [0, 2, 4], 4, r)
    for i, ell in enumerate([0, 2, 4]):
        assert np.max(np.abs(K[i].T - ref[ell])) < 1e-14
    xi = REF['a_xi']
    assert relmax(conv([0, 2, 4], r, xi), REF['a_xi_conv']) < 1e-14
    with pytest.raises(ValueError):
        conv([0, 2, 4], r, xi[:, :2])


def test_simpson_weights_match_reference_SimpsIntegrate(oracle):
    """the folded Simpson weights of MultipoleTransfer against the reference C++ SimpsIntegrate (DiscreteQuad.cpp:40-66)"""
    from pyrsd_b200.transfers import simpson_weights
    rng = np.random.default_rng(3)
    for n in (3, 11, 41):
        x = np.linspace(0., 1., n)
        y = rng.normal(size=n)
        assert abs(simpson_weights(x) @ y - oracle.SimpsIntegrate(x, y)) < 1e-14
        xs = np.sort(rng.random(n))              # uneven spacing
        assert abs(simpson_weights(xs) @ y - oracle.SimpsIntegrate(xs, y)) < 1e-12
    with pytest.raises(ValueError):
        simpson_weights(np.linspace(0., 1., 40))


def test_continuous_transfer_weights():
    from pyrsd_b200 import transfers as TR
    from scipy.special import legendre
    k = np.logspace(-2, -0.4, 12)
    mt = TR.MultipoleTransfer.__new__(TR.MultipoleTransfer)
    TR.MultipoleTransfer.__init__(mt, k, [0, 2, 4], Nmu=40, context=object())
    assert mt.grid.Nmu == 41
    w = mt._weights()
    mu = mt.grid.mu_cen
    P = (1. + 0.5*mu[None, :]**2)**2*(1. + k[:, None])
    got = np.einsum('bkm,km->kb', w, P)
    # Kaiser-like input: P_0 = (1 + b/3 + b^2/20 ...) known in closed form: int (1 + mu^2/2)^2 = 1 + 1/3 + 1/20
    assert np.allclose(got[:, 0], (1. + 1./3 + 1./20)*(1. + k), rtol=1e-6)
    ref2 = np.array([np.dot(TR.simpson_weights(mu), 5.*legendre(2)(mu)*row) for row in P])
    assert np.allclose(got[:, 1], ref2, rtol=1e-13)
    wt = TR.WedgeTransfer.__new__(TR.WedgeTransfer)
    bounds = [(0., 0.2), (0.2, 0.4), (0.4, 0.6), (0.6, 0.8), (0.8, 1.0)]
    TR.WedgeTransfer.__init__(wt, k, list(bounds), Nmu=40, context=object())
    ww = wt._weights()
    assert ww.shape == (5, 12, 41) and np.allclose(ww.sum(axis=-1), 1.)
    # wedges.py:84-95: plain mean of the samples with lo <= mu < hi (the last bin includes mu = 1)
    mug = wt.grid.mu_cen
    for b, (lo, hi) in enumerate(bounds):
        sel = (mug >= lo) & (mug < (1.0005 if hi == 1.0 else hi))
        assert np.allclose(ww[b, 0], sel/sel.sum())
