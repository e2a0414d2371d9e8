"""GPU parity of the transfer functions (gridded wedges / multipoles, window-function convolution) against the
golden vectors of the reference's own Python (tests/golden/window_ref.npz) and the NumPy oracle, through the
C-ABI (prsd_gridop_*, prsd_window_*, prsd_spline_resample_device)."""
import os
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from window_inputs import synthetic_power, synthetic_window  # noqa: E402

REF = np.load(os.path.join(HERE, 'golden', 'window_ref.npz'))
# deterministic transforms: the budget is rounding (1 ulp of the input moves xi by ~3e-13 of its maximum, see
# tests/golden/make_window_ref.py), far inside the 1e-5 of the north star
TOL = 1e-10


def relmax(a, b):
    return np.nanmax(np.abs(a - b))/np.nanmax(np.abs(b))


def grid_of(tag):
    kmin, kmax, Nk, Nmu = REF[tag + '_grid']
    return float(kmin), float(kmax), int(Nk), int(Nmu)


def test_gridded_wedges_and_poles_vs_reference():
    from pyrsd_b200 import transfers as TR
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        grid = TR.PkmuGrid([REF['g_k_cen'], REF['g_mu_cen']], REF['g_k'], REF['g_mu'], REF['g_modes'])
    bounds = [tuple(b) for b in REF['g_bounds']]
    power = np.where(grid.notnull, REF['g_power'], np.nan)          # null cells hold NaN: they must not leak
    in_range = (REF['g_k_cen'] >= 0.01) & (REF['g_k_cen'] <= 0.3)
    # the low-k rows of this grid have empty mu bins: the reference raises there, so start above them
    kmin = REF['g_k_cen'][6] - 1e-9
    wt = TR.GriddedWedgeTransfer(grid, list(bounds), kmin=kmin, kmax=0.3)
    got = wt(power)
    sel = (REF['g_k_cen'] >= kmin) & (REF['g_k_cen'] <= 0.3)
    assert np.isnan(got[~sel]).all() and np.isfinite(got[sel]).all()
    assert relmax(got[sel], REF['g_wedges'][sel]) < 1e-13
    mt = TR.GriddedMultipoleTransfer(grid, [0, 2, 4], kmin=0.01, kmax=0.3)
    got = mt(power)
    assert np.isnan(got[~in_range]).all()
    assert relmax(got[in_range], REF['g_poles'][in_range]) < 1e-13
    # batch of walkers, and device tensors in / out
    import torch
    batch = np.stack([power, 2.*power, power + 1.])
    gb = mt(batch)
    assert relmax(gb[0][in_range], got[in_range]) == 0.
    dev = mt(torch.from_numpy(np.nan_to_num(batch)).cuda())
    assert dev.is_cuda and tuple(dev.shape) == (3, 3, grid.Nk)
    assert relmax(dev.cpu().numpy()[1].T[in_range], gb[1][in_range]) < 1e-15
    modes = REF['g_modes'].copy()
    modes[10, REF['g_mu'][10] < 0.2] = 0.                            # an empty mu bin inside the k range
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        holed = TR.PkmuGrid([REF['g_k_cen'], REF['g_mu_cen']], REF['g_k'], REF['g_mu'], modes)
    with pytest.raises(ValueError):
        TR.GriddedWedgeTransfer(holed, list(bounds), kmin=0.01, kmax=0.3)(power)
    got = TR.GriddedWedgeTransfer(holed, list(bounds), kmin=0.06, kmax=0.3)(power)   # row 10 (k = 0.0575) excluded
    assert np.isfinite(got[(REF['g_k_cen'] >= 0.06) & (REF['g_k_cen'] <= 0.3)]).all()
    with pytest.raises(ValueError):
        mt(power[:, :10])


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_window_transfer_vs_reference(tag):
    from pyrsd_b200 import transfers as TR
    kmin, kmax, Nk, Nmu = grid_of(tag)
    tr = TR.WindowFunctionTransfer(synthetic_window(), [0, 2, 4], kmin=kmin, kmax=kmax, Nk=Nk, Nmu=Nmu, max_ellprime=4)
    k, mu = tr.grid.k_cen, tr.grid.mu_cen
    power = synthetic_power(k[:, None], mu[None, :])
    full, xi, xic = tr(power, return_xi=True)
    assert tr.N_fft == int(REF[tag + '_N'].ravel()[0])
    assert np.array_equal(tr.k_pad, REF[tag + '_newk'])
    assert relmax(tr.rr, REF[tag + '_rr']) < 1e-14
    assert relmax(xi[0].T, REF[tag + '_xi']) < TOL
    assert relmax(xic[0].T, REF[tag + '_xi_conv']) < TOL
    assert relmax(full, REF[tag + '_Pconv']) < TOL
    out = tr(power, k_out=REF[tag + '_k_out'])
    assert relmax(out, REF[tag + '_Pout']) < TOL
    dry = tr(power, dry_run=True)
    assert relmax(dry, REF[tag + '_Pdry']) < TOL
    noconv = tr(power, no_convolution=True)
    assert relmax(noconv[:Nk], REF[tag + '_Pell0']) < 1e-14 and np.all(noconv[Nk:] == 0.)
    with pytest.raises(NotImplementedError):
        tr(power, extrap=True)


def test_window_transfer_batch_linearity_and_device_tensors():
    """at the reference's default size (Nk = 1024, N = 4096): a batch of walkers, odd batch size (the FFT kernel
    packs two sequences per CTA), linearity of the whole chain, and device-resident input / output"""
    import torch
    from pyrsd_b200 import transfers as TR
    tr = TR.WindowFunctionTransfer(synthetic_window(), [0, 2, 4])
    k, mu = tr.grid.k_cen, tr.grid.mu_cen
    p1 = synthetic_power(k[:, None], mu[None, :])
    p2 = synthetic_power(k[:, None], mu[None, :], b=1.3, f=0.5, sigma=5.0)
    batch = np.stack([p1, p2, 0.3*p1 - 1.7*p2, p1, p2])
    k_out = np.linspace(0.01, 0.4, 40)
    out = tr(batch, k_out=k_out)
    assert out.shape == (5, 40, 3)
    assert relmax(out[0], REF['a_Pout']) < TOL
    # two sequences share one complex FFT: the same walker in another slot agrees to rounding, not bit for bit
    assert relmax(out[3], out[0]) < 1e-12 and relmax(out[4], out[1]) < 1e-12
    assert relmax(out[2], 0.3*out[0] - 1.7*out[1]) < 1e-11
    dev = tr(torch.from_numpy(batch).cuda(), k_out=k_out)
    assert dev.is_cuda and tuple(dev.shape) == (5, 3, 40)
    assert relmax(np.swapaxes(dev.cpu().numpy(), 1, 2), out) < 1e-15


def test_window_errors():
    from pyrsd_b200 import transfers as TR
    with pytest.raises(ValueError):
        TR.WindowFunctionTransfer(synthetic_window(), [0, 2])               # needs the first 3 even multipoles
    with pytest.raises(ValueError):
        TR.WindowFunctionTransfer(synthetic_window()[:, :4], [0, 2, 4])     # window multipoles up to ell = 8
    from pyrsd_b200._native import NativeError
    with pytest.raises(NativeError):
        TR.WindowFunctionTransfer(synthetic_window(), [0, 2, 4], Nk=100)._window()     # FFT length below 1024


def test_survey_likelihood_chain_on_device(golden):
    """PT tables -> galaxy P(k, mu) on the window grid -> multipoles -> window convolution -> spline -> chi^2, all
    on the device (PTEngine.galaxy_window_chi2), against the same chain assembled on the host from this
    library's P(k, mu) with the NumPy oracle of the reference's transfer functions"""
    from pyrsd_b200 import rsd, transfers as TR
    from oracle import transfer_restate as T
    from tests.gpu_common import golden_pygcl_cosmology, b2_00, b2_01, stochasticity
    cosmo = golden_pygcl_cosmology(golden)
    engine = rsd.PTEngine()
    sp = engine.spectra(cosmo.GetDiscreteK(), cosmo.GetDiscreteTk(), cosmo.n_s(), cosmo.delta_H()**2, cosmo.h(),
                        cosmo.Omega0_m(), cosmo.Omega0_b(), cosmo.Tcmb())
    res = engine.run(sp, result=rsd._new_result())
    kmin, kmax = 1e-4, 0.7
    kmodel = np.logspace(np.log10(kmin), np.log10(kmax), 200)
    gal = engine.galaxy(kmodel, rsd.galaxy_config())
    ks = rsd.stochasticity_grid(kmodel)
    nw = 5
    wpar = np.zeros((nw, rsd.NPAR)); stoch = np.zeros((nw, rsd.NPAIR, rsd.NSTOCH))
    for i in range(nw):
        b1 = [1.8 + 0.05*i, 2.7, 2.5, 3.4]
        b00, b01 = [b2_00(b) for b in b1], [b2_01(b) for b in b1]
        wpar[i] = rsd.pack_walker(0.6 + 0.01*i, 0.75, 1.0, 1.0, 1.0, b1, 0.10, 0.08, 0.40, 1.0, 4.2, 6.0, 3e4, 9e4, 0., 0., None,
                                  0.7486, cosmo.sigma8(), [b00, b00, b00, b00, b01, b01])
        for p, (a, b) in enumerate(rsd.PAIRS):
            lo, hi = sorted([b1[a], b1[b]])
            stoch[i, p] = stochasticity(ks, wpar[i, 0], lo, hi)
    cmap = np.zeros(nw, dtype=np.int32)
    window = synthetic_window()
    tr = TR.WindowFunctionTransfer(window, [0, 2, 4], kmin=kmin, kmax=kmax, Nk=1024, Nmu=40)
    k_out = np.linspace(0.02, 0.38, 19)
    # host chain: this library's P(k, mu), the reference's transfer arithmetic (NumPy restatement)
    P = engine.galaxy_power(gal, sp, res, wpar, stoch, tr.grid.k.ravel(), tr.grid.mu.ravel(), cmap=cmap)
    P = P.reshape(nw, 1024, 40)
    ref = np.array([T.window_transfer(P[i], tr.grid.k_cen, tr.grid.mu_cen, window, [0, 2, 4], 4, k_out)['Pout'].T
                    for i in range(nw)])                                              # [nw][ell][k]
    rng = np.random.default_rng(5)
    n = 3*len(k_out)
    data = ref[2].ravel()*(1 + 0.01*rng.normal(size=n))
    A = rng.normal(size=(n, n))
    cinv = (A @ A.T/n + np.eye(n))/(np.outer(np.abs(data), np.abs(data))*1e-4)
    d = ref.reshape(nw, n) - data[None, :]
    chi2_ref = np.einsum('wi,ij,wj->w', d, cinv, d)
    chi2, poles = engine.galaxy_window_chi2(gal, sp, res, wpar, stoch, tr, k_out, data, cinv, cmap=cmap, return_poles=True)
    assert poles.is_cuda and tuple(poles.shape) == (nw, 3, len(k_out))
    got = poles.cpu().numpy()
    assert np.max(np.abs(got - ref))/np.max(np.abs(ref)) < TOL
    # chi^2 amplifies the 1e-10 of the multipoles by |d|/sigma ~ 1e2
    assert np.allclose(chi2, chi2_ref, rtol=1e-6)
    assert np.argmin(chi2) == 2
    assert np.allclose(engine.chi2(ref, data, cinv), chi2_ref, rtol=1e-12)


def test_continuous_multipole_and_wedge_transfers(oracle):
    """MultipoleTransfer (Simpson over mu, rsd/transfers/poles.py) and WedgeTransfer (rsd/transfers/wedges.py) on the
    device against the reference C++ SimpsIntegrate / a plain NumPy mean"""
    from pyrsd_b200 import transfers as TR
    from scipy.special import legendre
    k = np.logspace(-2, np.log10(0.4), 37)
    mt = TR.MultipoleTransfer(k, [0, 2, 4], Nmu=40)
    mu = mt.grid.mu_cen
    assert len(mu) == 41
    P = synthetic_power(k[:, None], mu[None, :])
    got = mt(P)
    assert got.shape == (37, 3)
    for i, ell in enumerate([0, 2, 4]):
        kern = (2*ell + 1.)*legendre(ell)(mu)
        ref = np.array([oracle.SimpsIntegrate(mu, kern*row) for row in P])
        assert np.max(np.abs(got[:, i] - ref)) < 1e-12*np.max(np.abs(P))
    bounds = [(0., 0.2), (0.2, 0.4), (0.4, 0.6), (0.6, 0.8), (0.8, 1.0)]
    wt = TR.WedgeTransfer(k, list(bounds), Nmu=40)
    mug = wt.grid.mu_cen
    Pw = synthetic_power(k[:, None], mug[None, :])
    got = wt(Pw)
    for b, (lo, hi) in enumerate(bounds):
        sel = (mug >= lo) & (mug < (1.0005 if hi == 1.0 else hi))
        assert np.max(np.abs(got[:, b] - Pw[:, sel].mean(axis=1))) < 1e-12*np.max(np.abs(Pw))


def test_convolve_multipoles_vs_reference():
    """the older window route (rsd/window.py:214-307) vs the UNMODIFIED reference function run on the reference C++
    transforms (tests/golden/make_window_ref.py): trapezoid branch with the FITPACK refinement, and biased FFTLog branch"""
    from pyrsd_b200.transfers import WindowConvolution, convolve_multipoles
    window = REF['window']
    conv = WindowConvolution(window[:, 0], window[:, 1:], max_ellprime=4, max_ell=4)
    ells = [0, 2, 4]
    k, out = convolve_multipoles(REF['c_k'], REF['c_Pell'], ells, conv, legacy=True)
    assert np.array_equal(k, REF['c_k']) and out.shape == REF['c_legacy'].shape
    assert np.max(np.abs(out - REF['c_legacy'])) < 1e-9*np.max(np.abs(REF['c_legacy']))
    _, dry = convolve_multipoles(REF['c_k'], REF['c_Pell'], ells, conv, dry_run=True, legacy=True)
    assert np.max(np.abs(dry - REF['c_legacy_dry'])) < 1e-9*np.max(np.abs(REF['c_legacy_dry']))
    kk, Pc = convolve_multipoles(REF['f_k'], REF['f_Pell'], ells, conv, qbias=0.7, legacy=False)
    assert np.max(np.abs(kk/REF['f_kout'] - 1)) < 1e-12
    for i in range(3):
        assert np.max(np.abs(Pc[:, i] - REF['f_conv'][:, i])) < 1e-9*np.max(np.abs(REF['f_conv'][:, i])), i
    with pytest.raises(ValueError):
        convolve_multipoles(REF['c_k'], REF['c_Pell'], [0, 2], conv)
    with pytest.raises(ValueError):
        convolve_multipoles(REF['c_k'], REF['c_Pell'], [0, 2, 3], conv)


def test_peer_exchange_single_rank_matches_gridop():
    """the exchange path of the multi-GPU step (csrc/exchange.cu) with one rank: the gridded transfer written through the
    peer table + flag kernel equals ``prsd_gridop_apply``, step after step, alternating buffer halves"""
    import ctypes as C
    import torch
    from pyrsd_b200 import _native as N, transfers as T, parallel
    L = T._L()
    ctx = N.default_context()
    nw, nk, nmu = 5, 37, 12
    rng = np.random.default_rng(5)
    w = rng.random((2, nk, nmu))
    gop = C.c_void_p()
    N.check(L.prsd_gridop_create(ctx.ptr, 2, nk, nmu, w, C.byref(gop)))
    ex = parallel.PeerExchange(ctx, nw*2*nk)
    dev = torch.device('cuda', ctx.device)
    for step in range(3):
        P = rng.random((nw, nk, nmu))
        d_P = torch.from_numpy(P).to(dev)
        d_ref = torch.empty((nw, 2, nk), dtype=torch.float64, device=dev)
        N.check(L.prsd_gridop_apply(gop, nw, C.c_void_p(d_P.data_ptr()), C.c_void_p(d_ref.data_ptr()), 1))
        ex.apply(gop, nw, C.c_void_p(d_P.data_ptr()))
        got = ex.received(check=True)
        ctx.sync()
        assert got.shape == (1, nw*2*nk)
        assert torch.equal(got.reshape(nw, 2, nk), d_ref)
        assert np.allclose(d_ref.cpu().numpy(), np.einsum('wkm,bkm->wbk', P, w), rtol=1e-13)
    with pytest.raises(N.NativeError):
        ex.apply(gop, nw + 1, C.c_void_p(d_P.data_ptr()))
    ex.close()
    N.check(L.prsd_gridop_destroy(gop))
