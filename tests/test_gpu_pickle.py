"""The pickling contract of the pygcl classes (pyRSD/pygcl.py:28-34, 60-63, 198-201, 251-253: the state is the
constructor arguments, plus the tables that are expensive to rebuild): a pickled object evaluates identically.
The reference relies on it for ``np.save(model)`` caches and MPI pools."""
import pickle

import numpy as np
import pytest

from tests.gpu_common import golden_pygcl_cosmology

pytestmark = pytest.mark.gpu


def roundtrip(obj):
    return pickle.loads(pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL))


def test_pickle_roundtrip_of_the_pt_objects(golden):
    from pyrsd_b200 import pygcl
    cosmo = golden_pygcl_cosmology(golden)
    k = np.logspace(-2, np.log10(0.4), 9)

    c2 = roundtrip(cosmo)
    assert c2.sigma8() == cosmo.sigma8() and c2.n_s() == cosmo.n_s()
    assert np.array_equal(c2.GetDiscreteTk(), cosmo.GetDiscreteTk())
    assert c2.delta_H() == cosmo.delta_H()

    P = pygcl.LinearPS(cosmo, 0.55)
    P.SetSigma8AtZ(0.61)
    P2 = roundtrip(P)
    assert P2.GetRedshift() == 0.55 and P2.GetSigma8AtZ() == 0.61         # the rescaling survives (pygcl.py:118-124)
    assert np.array_equal(P2(k), P(k))

    P0 = pygcl.LinearPS(cosmo, 0.)
    for obj, args in ((pygcl.Imn(P0), (k, 2, 2)), (pygcl.Jmn(P0), (k, 0, 2)), (pygcl.Kmn(P0), (k, 1, 1, True, 0))):
        twin = roundtrip(obj)
        assert np.array_equal(twin(*args), obj(*args)), type(obj).__name__

    # one-loop spectra carry (epsrel, table): the table is NOT recomputed on unpickling (pygcl.py:198-201)
    table = golden['oneloop_Pdd_0']
    O = pygcl.OneLoopPdd(P0, 1e-4, table)
    O2 = roundtrip(O)
    assert O2.GetEpsrel() == 1e-4 and np.array_equal(O2.GetOneLoopPower(), table)
    assert np.array_equal(O2(k), O(k))
    Ovv = pygcl.OneLoopPvv(P0, 1e-4, golden['oneloop_Pvv_0'])
    M = pygcl.ImnOneLoop(Ovv, O, 1e-4)
    M2 = roundtrip(M)
    assert np.array_equal(M2.EvaluateCross(k, 0, 1), M.EvaluateCross(k, 0, 1))

    # Zel'dovich state = (cosmo, approx_lowk, sigma8_z, k0_low, sigma^2, X0, XX, YY) (pygcl.py:251-253)
    Z = pygcl.ZeldovichP00(cosmo, 0.55, True)
    Z.SetSigma8AtZ(0.7)
    Z2 = roundtrip(Z)
    assert Z2.GetSigma8AtZ() == 0.7 and Z2.GetApproxLowKFlag() and Z2.GetSigmaSq() == Z.GetSigmaSq()
    assert np.array_equal(Z2.GetXZel(), Z.GetXZel())
    kz = np.array([2e-3, 0.01, 0.1, 0.3])
    assert np.array_equal(Z2(kz), Z(kz))
    zcf = pygcl.ZeldovichCF(cosmo, 0.55)
    r = np.array([20., 80., 105.])
    assert np.array_equal(roundtrip(zcf)(r), zcf(r))


def test_galaxy_spectrum_pickle_and_npy(golden, tmp_path):
    import pickle as pk
    from pyrsd_b200 import rsd
    from tests.gpu_common import b2_00, b2_01, stochasticity
    cosmo = golden_pygcl_cosmology(golden)
    m = rsd.GalaxySpectrum(params=cosmo, z=0.55, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity,
                           use_so_correction=True, fog_model='gaussian', max_mu=6)
    m.update(f=0.78, sigma8_z=0.61, f_so=0.04, sigma_so=4.0, alpha_par=0.97, fs=0.12)
    k = np.logspace(-2, np.log10(0.35), 11)
    mu = np.linspace(0., 1., 5)
    P = m.power(k, mu)
    m2 = pk.loads(pk.dumps(m))
    assert m2._result is None and m2.f == 0.78 and m2.use_so_correction and m2.fog_model == 'gaussian'
    assert np.array_equal(m2.power(k, mu), P)
    path = str(tmp_path / 'model.npy')
    m.to_npy(path)
    m3 = rsd.GalaxySpectrum.from_npy(path)
    assert np.array_equal(m3.power(k, mu), P)
    assert np.array_equal(np.array(m3.poles(k, [0, 2])), np.array(m.poles(k, [0, 2])))
