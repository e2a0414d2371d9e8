"""GPU parity of ``ExtrapolatedPowerSpectrum`` / ``SmoothedXiMultipoles`` (pyrsd_b200/correlation.py) against the reference's
own classes run on the reference ``GalaxySpectrum`` (tests/golden/xi_ref.npz, tests/golden/make_xi_ref.py)."""
import os

import numpy as np
import pytest

from tests.gpu_common import golden_pygcl_cosmology, b2_00, b2_01, stochasticity, record_parity

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, 'golden', 'xi_ref.npz')
KW = dict(k_lo=2e-3, k_hi=0.4)


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/xi_ref.npz not generated")
def test_extrapolated_power_and_xi_multipoles(golden):
    from pyrsd_b200 import rsd
    from pyrsd_b200.correlation import ExtrapolatedPowerSpectrum, SmoothedXiMultipoles
    ref = np.load(REF)
    cosmo = golden_pygcl_cosmology(golden)
    model = rsd.GalaxySpectrum(params=cosmo, z=0.55, interpolate=False, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity)
    model.update(sigma8_z=0.61, f=0.77)
    ex = ExtrapolatedPowerSpectrum(model, **KW)
    mu = ref['mu']
    # fitted exponents: least squares on 1e-5-accurate inputs
    assert np.max(np.abs(ex.low_k_params(mu) - ref['alpha_lo'])) < 1e-5
    a, b = ex.high_k_params(mu)
    assert np.max(np.abs(a - ref['alpha_hi'])) < 2e-4 and np.max(np.abs(b - ref['beta_hi'])) < 2e-4
    P = ex(ref['k'], mu)
    assert P.shape == ref['extrap'].shape
    err = np.abs(P/ref['extrap'] - 1)
    inside = (ref['k'] >= KW['k_lo']) & (ref['k'] <= KW['k_hi'])
    assert err[inside].max() < 1e-5
    # extrapolated: amplitude x (k / k_hi)^(alpha + beta log10) up to k = 9 = 22 k_hi amplifies the exponents' differences
    assert err[~inside].max() < 5e-4
    poles = ex.to_poles(np.logspace(-5, 1, 1000), [0, 2, 4])
    assert poles.shape == ref['poles'].shape
    # scale: the multipole or, where it is small (the hexadecapole is a few per cent of the monopole: an integral over mu that
    # cancels), a tenth of the monopole
    scale = np.maximum(np.abs(ref['poles']), 0.1*np.abs(ref['poles'][:, :1]))
    assert np.max(np.abs(poles - ref['poles'])/scale) < 5e-4
    sel = (np.logspace(-5, 1, 1000) >= KW['k_lo']) & (np.logspace(-5, 1, 1000) <= KW['k_hi'])
    assert np.max((np.abs(poles - ref['poles'])/scale)[sel]) < 1e-5
    r = ref['r']
    xi0 = SmoothedXiMultipoles(model, r, [0, 2, 4], R=0., **KW)
    xi3 = SmoothedXiMultipoles(model, r, [0, 2], R=3., **KW)
    for name, mine, want in (('xi_ell(s), R = 0', xi0, ref['xi_R0']), ('xi_ell(s), R = 3', xi3, ref['xi_R3'])):
        assert mine.shape == want.shape
        # scale: the monopole at that separation (the higher multipoles cross zero)
        scale = np.maximum(np.abs(want), np.abs(want[:, :1]))
        record_parity('correlation', name, float(np.max(np.abs(mine - want)/scale)))
        assert np.max(np.abs(mine - want)/scale) < 2e-4
    record_parity('correlation', 'P(k, mu) inside [k_lo, k_hi]', float(err[inside].max()))
    record_parity('correlation', 'P(k, mu) extrapolated (to k = 9)', float(err[~inside].max()))
    # the default k_lo = kmin / k_hi = kmax lies outside the reference's own fit grid: same error as there
    with pytest.raises(ValueError):
        ExtrapolatedPowerSpectrum(model)(np.array([1e-4]), mu)
