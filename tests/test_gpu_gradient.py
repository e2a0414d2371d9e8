"""GPU parity of GalaxySpectrum.gradient (batched central differences on the device) against the reference's
ANALYTIC derivatives (pyRSD/rsd/power/gal/derivatives/*.py, run unmodified on the tight-tolerance reference model,
tests/golden/make_gradient_ref.py) and against the reference's own finite differences for the rest."""
import os

import numpy as np
import pytest

from tests.gpu_common import golden_pygcl_cosmology, b2_00, b2_01, stochasticity

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, 'golden', 'gradient_ref.npz')
# the model itself matches the reference to 1e-5 of P; a derivative inherits that, relative to its own scale
TOL = 2e-5

CASES = {
    'ap': (dict(),
           dict(alpha_par=1.03, alpha_perp=0.98, b1_cA=1.9, b1_cB=2.7, b1_sA=2.4, b1_sB=3.3, fs=0.12, fcB=0.09, fsB=0.35,
                sigma_c=1.1, sigma_sA=3.9, sigma_sB=5.5, NcBs=4.2e4, NsBsB=8.1e4, sigma8_z=0.61, f=0.78)),
    'so': (dict(use_so_correction=True, fog_model='gaussian', max_mu=6, use_tidal_bias=True),
           dict(f_so=0.04, sigma_so=4.0, alpha_par=0.97, alpha_perp=1.02)),
}


@pytest.fixture(scope='module')
def ref():
    if not os.path.exists(REF):
        pytest.skip("tests/golden/gradient_ref.npz not generated")
    return np.load(REF)


@pytest.fixture(scope='module')
def models(golden):
    from pyrsd_b200 import rsd
    cosmo = golden_pygcl_cosmology(golden)
    engine = rsd.PTEngine()
    out = {}
    for tag, (kw, upd) in CASES.items():
        m = rsd.GalaxySpectrum(params=cosmo, z=0.55, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity, engine=engine, interpolate=False, **kw)
        m.update(**upd)
        out[tag] = m
    return out


@pytest.mark.parametrize('tag', ['ap', 'so'])
def test_gradient_vs_reference_analytic(ref, models, tag):
    m = models[tag]
    k, mu = ref['k'], ref['mu']
    P = m.power(k, mu)
    assert np.max(np.abs(P/ref[tag + '_P'] - 1.)) < 1e-5
    # fsB: the reference's analytic form keeps only the satellite-satellite part (derivatives/fsB.py:14-15; compare
    # fcB.py:16-17, which has the cross term) -- it is pinned by the reference's finite difference below instead
    names = [n[len(tag) + 3:] for n in ref.files if n.startswith(tag + '_d_') and not n.endswith('_fsB')]
    assert len(names) >= 7
    # P is linear in the one-halo amplitudes: any step is exact, a unit step avoids cancellation
    eps = [1.0 if n in ('NcBs', 'NsBsB') else 1e-4 for n in names]
    g = m.gradient(k, mu, names, epsilon=eps)
    assert g.shape == (len(names), len(k))
    for n, row in zip(names, g):
        want = ref['%s_d_%s' % (tag, n)]
        err = np.max(np.abs(row - want))/np.max(np.abs(want))
        assert err < TOL, (tag, n, err)
    # the model is left untouched
    assert np.array_equal(m.power(k, mu), P)


def test_gradient_vs_reference_finite_differences(ref, models):
    m = models['ap']
    k, mu = ref['k'], ref['mu']
    names = [n[len('ap_n_'):] for n in ref.files if n.startswith('ap_n_')]
    g = m.gradient(k, mu, names)                  # epsilon = 1e-4, the reference's default
    for n, row in zip(names, g):
        want = ref['ap_n_' + n]
        err = np.max(np.abs(row - want))/np.max(np.abs(want))
        # both sides difference P at +-1e-4: the 1e-5-level mismatch of P is smooth in the parameter and cancels, what
        # is left is the rounding of the reference's own difference quotient
        assert err < 5e-5, (n, err)


def test_gradient_properties(models):
    m = models['ap']
    k = np.logspace(-2, np.log10(0.3), 12)
    mu = np.linspace(0., 1., 5)
    g = m.gradient(k, mu, 'fs')                   # broadcast grid, scalar name
    assert g.shape == (1, 60)
    # step-size independence where the model is smooth: O(eps^2) truncation
    g2 = m.gradient(k, mu, ['fs', 'sigma8_z', 'f'], epsilon=[2e-4, 2e-4, 2e-4])
    g1 = m.gradient(k, mu, ['fs', 'sigma8_z', 'f'])
    assert np.max(np.abs(g1 - g2))/np.max(np.abs(g1)) < 1e-6
    # f_so does nothing without the SO correction
    assert np.all(m.gradient(k, mu, 'f_so') == 0.)
    with pytest.raises(ValueError):
        m.gradient(k, mu, ['not_a_parameter'])
    with pytest.raises(ValueError):
        m.gradient(k, mu, ['fs'], epsilon=0.)


def test_component_spectra(ref, models):
    """Pgal_cc / Pgal_cs / Pgal_ss (power_gal.py:316-399): they recompose power(), and the reference's analytic
    dP/dfs = -2 (1 - fs) Pcc + 2 (1 - 2 fs) Pcs + 2 fs Pss (derivatives/fs.py:11-15) evaluated with them matches the
    reference's value"""
    for tag in ('ap', 'so'):
        m = models[tag]
        k, mu = ref['k'], ref['mu']
        Pcc, Pcs, Pss = m.components(k, mu)
        fs = m.fs
        P = m.power(k, mu)
        assert np.max(np.abs((1 - fs)**2*Pcc + 2*fs*(1 - fs)*Pcs + fs**2*Pss - P))/np.max(np.abs(P)) < 1e-13
        dfs = -2*(1 - fs)*Pcc + 2*(1 - 2*fs)*Pcs + 2*fs*Pss
        want = ref[tag + '_d_fs']
        assert np.max(np.abs(dfs - want))/np.max(np.abs(want)) < TOL
        assert np.array_equal(m.Pgal_cc(k, mu), Pcc) and np.array_equal(m.Pgal_ss(k, mu), Pss)
        assert m.fs == fs
    grid = models['ap'].Pgal_cs(np.logspace(-2, -1, 4), np.linspace(0, 1, 3))
    assert grid.shape == (4, 3)
    assert models['ap'].Pgal_cs(np.logspace(-2, -1, 4), np.linspace(0, 1, 3), flatten=True).shape == (12,)


def test_derivative_k_mu_chain_rule_vs_reference_alpha_derivatives(ref, models):
    """derivative_k / derivative_mu (power_gal.py:427-441) through the chain rule the reference's analytic alpha_par and
    alpha_perp derivatives use (rsd/power/gal/derivatives/alpha_par.py, alpha_perp.py): dP_obs/d alpha = D_k dk'/d alpha +
    D_mu dmu'/d alpha - (volume factor derivative) P_obs, compared with the reference's values"""
    for tag in ('ap', 'so'):
        m = models[tag]
        k, mu = ref['k'], ref['mu']
        P = m.power(k, mu)
        Dk, Dmu = m.derivative_k(k, mu), m.derivative_mu(k, mu)
        ap, ae = m.alpha_par, m.alpha_perp
        F = ap/ae
        A = 1. + mu**2*(1./F**2 - 1.)
        kp, mup = k/ae*np.sqrt(A), mu/F/np.sqrt(A)
        # d/d alpha_par at fixed alpha_perp
        dA = mu**2*(-2./F**3)/ae
        dk_par = k/(2.*ae)*dA/np.sqrt(A)
        dmu_par = mu*(-1./F**2/ae/np.sqrt(A) - 0.5/F*dA/A**1.5)
        d_par = Dk*dk_par + Dmu*dmu_par - P/ap
        want = ref[tag + '_d_alpha_par']
        assert np.max(np.abs(d_par - want))/np.max(np.abs(want)) < TOL, tag
        # d/d alpha_perp at fixed alpha_par: F = ap / ae, dF/dae = -F / ae
        dA = mu**2*(-2./F**3)*(-F/ae)
        dk_perp = -kp/ae + k/(2.*ae)*dA/np.sqrt(A)
        dmu_perp = mu*(1./(F*ae)/np.sqrt(A) - 0.5/F*dA/A**1.5)
        d_perp = Dk*dk_perp + Dmu*dmu_perp - 2.*P/ae
        want = ref[tag + '_d_alpha_perp']
        assert np.max(np.abs(d_perp - want))/np.max(np.abs(want)) < TOL, tag
        assert np.array_equal(m.power(k, mu), P)          # the model is left untouched


def test_multipole_shortcuts_apply_transfer_and_bias_functions(models, golden):
    from pyrsd_b200 import pygcl, tools, transfers
    m = models['ap']
    k = np.logspace(-2, np.log10(0.3), 12)
    mus = np.linspace(0., 1., 41)
    P = m.power(k, mus)
    w = transfers.simpson_weights(mus)
    for ell, fn, kern in ((0, m.monopole, np.ones(41)), (2, m.quadrupole, 2.5*(3*mus**2 - 1.)),
                          (4, m.hexadecapole, 9./8.*(35*mus**4 - 30.*mus**2 + 3.))):
        assert np.max(np.abs(fn(k)/(P*kern*w).sum(-1) - 1)) < 1e-12, ell      # tools.monopole ... (rsd/tools.py:268-310)
    t = transfers.MultipoleTransfer(k, [0, 2], Nmu=40)
    assert np.max(np.abs(m.apply_transfer(t)/np.array(m.poles(k, [0, 2], Nmu=40)).T - 1)) < 1e-12
    # mass_from_bias / sigma_from_bias (rsd/tools.py:655-720): the Tinker bias of the mass that comes back is the bias that
    # went in, at the growth of the requested redshift
    cosmo = golden_pygcl_cosmology(golden)
    PL = pygcl.LinearPS(cosmo, 0.)
    b = np.array([1.4, 2.2, 3.1])
    M = tools.mass_from_bias(b, 0.55, PL)
    R = (3.*M/(4.*np.pi*cosmo.rho_bar_z(0.)))**(1./3.)
    assert np.max(np.abs(tools.bias_Tinker(cosmo.D_z(0.55)*PL.Sigma(R)) - b)) < 1e-6
    assert isinstance(tools.mass_from_bias(2.2, 0.55, PL), float) and abs(tools.mass_from_bias(2.2, 0.55, PL)/M[1] - 1) < 1e-12
    assert np.allclose(tools.sigma_from_bias(b, 0.55, PL), 3.6*(M/5.4903e13)**(1./3), rtol=1e-14)
