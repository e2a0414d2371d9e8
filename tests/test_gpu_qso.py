"""GPU parity of ``QuasarSpectrum`` against the UNMODIFIED reference class (tests/golden/qso_ref.npz,
tests/golden/make_qso_ref.py): Kaiser terms with the primordial non-Gaussianity bias, FOG, Alcock-Paczynski."""
import os
import pickle

import numpy as np
import pytest

from tests.gpu_common import golden_pygcl_cosmology

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, 'golden', 'qso_ref.npz')
TOL = 1e-5
CASES = {
    'gauss_fnl0': (dict(), dict(b1=2.3, sigma_fog=3.5, alpha_par=1.02, alpha_perp=0.98, sigma8_z=0.61, f=0.78)),
    'gauss_fnl': (dict(), dict(b1=2.3, f_nl=25., p=1.6, sigma_fog=3.5, N=120., sigma8_z=0.61, f=0.78)),
    'lorentz_fnl': (dict(fog_model='modified_lorentzian', max_mu=2), dict(b1=1.9, f_nl=-40., p=1.0, sigma_fog=5.0, alpha_par=0.97,
                                                                           alpha_perp=1.03, sigma8_z=0.55, f=0.7)),
}


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/qso_ref.npz not generated")
@pytest.mark.parametrize('case', sorted(CASES))
def test_quasar_spectrum_vs_reference_python(golden, case):
    from pyrsd_b200.qso import QuasarSpectrum
    ref = np.load(REF)
    cosmo = golden_pygcl_cosmology(golden)
    assert abs(cosmo['Omega0_m'] - float(ref['Omega0_m'])) < 1e-14
    kw, upd = CASES[case]
    m = QuasarSpectrum(params=cosmo, z=0.55, **kw)
    m.update(**upd)
    k, mu = ref['k'], ref['mu']
    P = m.power(k, mu)
    assert P.shape == ref[case + '_power'].shape
    assert np.max(np.abs(P/ref[case + '_power'] - 1)) < TOL
    mom = np.array([m.P_mu0(k), m.P_mu2(k), m.P_mu4(k), m.P_mu6(k)])
    rm = ref[case + '_moments']
    # (b_tot changes sign with f_nl < 0: a moment is measured against itself or 1e-3 of the largest moment at that k)
    scale = np.maximum(np.abs(rm[:3]), 1e-3*np.abs(rm[:3]).max(axis=0)[None, :])
    assert np.max(np.abs(mom[:3] - rm[:3])/scale) < TOL
    assert np.all(mom[3] == 0)
    assert np.max(np.abs(m.btot(k) - ref[case + '_btot'])) < TOL*np.max(np.abs(ref[case + '_btot']))
    assert np.max(np.abs(m.alpha_png(k)/ref[case + '_alpha_png'] - 1)) < TOL
    # multipoles by Simpson's rule of the power grid; pickling; parameter checks
    from pyrsd_b200 import pygcl
    mus = np.linspace(0., 1., 41)
    assert np.allclose(m.monopole(k), pygcl.SimpsIntegrate(mus, m.power(k, mus)), rtol=1e-12)
    clone = pickle.loads(pickle.dumps(m))
    assert np.array_equal(clone.power(k[:5], 0.3), m.power(k[:5], 0.3))
    with pytest.raises(ValueError):
        m.p = 2.0
    with pytest.raises(ValueError):
        m.fog_model = 'box'
