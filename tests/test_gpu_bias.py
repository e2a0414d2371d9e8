"""BiasToMassRelation / BiasToSigmaRelation on the device (pyrsd_b200/tools.py, csrc/biasrel.cu) against the UNMODIFIED
reference classes (rsd/tools.py:393-649) run on the reference C++ (tests/golden/make_bias_ref.py -> bias_ref.npz)."""
import os

import numpy as np
import pytest

from tests.gpu_common import golden_pygcl_cosmology

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, 'golden', 'bias_ref.npz')
TOL = 1e-5


@pytest.fixture(scope='module')
def cosmo(golden):
    return golden_pygcl_cosmology(golden)


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/bias_ref.npz not generated")
def test_bias_to_sigma_vs_reference(cosmo):
    from pyrsd_b200 import tools, pygcl
    r = np.load(REF)
    assert abs(cosmo.rho_bar_z(0.)/float(r['rho_bar_0']) - 1) < 1e-12
    assert abs(cosmo.H_z(0.55)/float(r['Hz']) - 1) < 1e-12
    assert np.max(np.abs(pygcl.LinearPS(cosmo, 0.).Sigma(r['Sigma_R'])/r['Sigma'] - 1)) < 1e-6
    s8, b1 = r['sigma8_z'], r['b1']
    fresh = tools.BiasToSigmaRelation(0.55, cosmo, interpolate=False)
    assert np.max(np.abs(fresh.mass(s8, b1)/r['mass_fresh'] - 1)) < TOL
    assert np.max(np.abs(fresh(s8, b1)/r['sigmav_fresh'] - 1)) < TOL
    assert isinstance(fresh(0.61, 2.3), float)
    interp = tools.BiasToSigmaRelation(0.55, cosmo, interpolate=True)
    tab = interp.interpolation_table
    assert tab.shape == (100, 70)
    mine = np.array([tab[0, 0], tab[0, -1], tab[-1, 0], tab[-1, -1], tab[50, 35]])
    assert np.max(np.abs(mine/r['table_corner_values'] - 1)) < TOL
    assert np.max(np.abs(interp.mass(s8, b1)/r['mass_interp'] - 1)) < TOL
    assert np.max(np.abs(interp(s8, b1)/r['sigmav_interp'] - 1)) < TOL
    norm = tools.BiasToSigmaRelation(0.55, cosmo, interpolate=False, sigmav_0=3.6, M_0=5.4903e13)
    assert np.max(np.abs(norm(s8[:5], b1[:5])/r['sigmav_normed'] - 1)) < TOL
    # the Tinker bias of the mass that comes back is the bias that went in
    M, sig = tools.bias_to_mass(fresh.power_lin._sp, s8/cosmo.sigma8(), b1, cosmo.rho_bar_z(0.), return_sigma=True,
                                cmap=np.zeros(len(b1), dtype=np.int32))
    R = (3.*M/(4.*np.pi*cosmo.rho_bar_z(0.)))**(1./3.)
    sigma_at = pygcl.LinearPS(cosmo, 0.).Sigma(R)*s8/cosmo.sigma8()
    assert np.max(np.abs(tools.bias_Tinker(sigma_at) - b1)) < 1e-6
    assert sig.shape == (1, 1000) and np.all(np.diff(sig[0]) < 0)


def test_bias_to_mass_batched_and_errors(cosmo):
    """many walkers on many cosmologies in one launch; an unreachable bias raises like brentq"""
    from pyrsd_b200 import tools, pygcl, _native
    P = pygcl.LinearPS(cosmo, 0.)
    n = 4096
    rng = np.random.default_rng(9)
    res, b1 = rng.uniform(0.4, 1.2, n), rng.uniform(1.0, 5.0, n)
    M = tools.bias_to_mass(P._sp, res, b1, cosmo.rho_bar_z(0.), cmap=np.zeros(n, dtype=np.int32))
    assert np.all(np.isfinite(M)) and np.all((M > 1e5) & (M < 1e16))
    one = tools.bias_to_mass(P._sp, res[:3], b1[:3], cosmo.rho_bar_z(0.), cmap=np.zeros(3, dtype=np.int32))
    assert np.array_equal(one, M[:3])
    # higher bias at fixed sigma8 -> higher mass
    Ms = tools.bias_to_mass(P._sp, np.full(5, 0.8), np.linspace(1.2, 4.0, 5), cosmo.rho_bar_z(0.), cmap=np.zeros(5, dtype=np.int32))
    assert np.all(np.diff(Ms) > 0)
    with pytest.raises(_native.NativeError):
        tools.bias_to_mass(P._sp, [0.8], [0.2], cosmo.rho_bar_z(0.), cmap=[0])       # below the Tinker floor: no sign change
