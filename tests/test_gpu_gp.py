"""Gaussian-process simulation fits on the device (pyrsd_b200/simulation.py, csrc/gp.cu) against the NumPy restatement of
the reference's george pipeline (oracle/gp_restate.py) on the shipped training data."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_predict(fit, pts):
    from oracle import gp_restate as G
    X = (fit.x - fit.x_mean)/fit.x_scale
    Xp = (np.asarray(pts, dtype=float).reshape(-1, X.shape[1]) - fit.x_mean)/fit.x_scale
    return fit.y_mean + fit.y_scale*(G.kernel(Xp, X, fit.amp, fit.metric)@fit.alpha)


def test_every_process_matches_the_restatement():
    from pyrsd_b200 import simulation as S
    rng = np.random.default_rng(12)
    for cls in (S.VelocityDispersionFits, S.AutoStochasticityFits, S.CrossStochasticityFits, S.Pmu2ResidualCorrection,
                S.Pmu4ResidualCorrection):
        fit = cls()
        lo, hi = fit.x.min(axis=0), fit.x.max(axis=0)
        pts = rng.uniform(lo, hi, (64, fit.x.shape[1]))
        mine = fit.predict(pts)
        ref = _numpy_predict(fit, pts)
        assert np.max(np.abs(mine - ref)) < 1e-10*np.max(np.abs(ref)), cls.__name__
        # at a training point with a tiny error bar the process reproduces the measurement scale
        assert np.all(np.isfinite(fit.predict(fit.x[:5])))
    nb = S.NonlinearBiasFits()
    b1 = np.linspace(1.0, 5.0, 17)
    out = nb(b1=b1)
    assert len(out) == 6
    for name, v in zip(nb.dependents, out):
        # (amplitudes theta_0 ~ 100 make K ill-conditioned: the weights are +-1e5 and cancel, summation order shows at 1e-9)
        assert np.max(np.abs(v - _numpy_predict(nb._data[name], b1[:, None]))) < 1e-7*np.max(np.abs(v))
    assert isinstance(nb(b1=2.0, select='b2_00_a')[0], float)


def test_call_signature_and_broadcasting():
    from pyrsd_b200 import simulation as S
    auto = S.AutoStochasticityFits()
    k = np.logspace(-3, np.log10(0.5), 20)
    lam = auto(sigma8_z=0.62, b1=2.1, k=k)
    assert lam.shape == (20,)
    assert np.allclose(lam, [auto(sigma8_z=0.62, b1=2.1, k=kk) for kk in k], rtol=1e-13)
    with pytest.raises(ValueError):
        auto(sigma8_z=0.6, b1=2.0)                 # `k` missing
    vd = S.VelocityDispersionFits()
    assert vd(sigma8_z=0.6, b1=2.0) > 0
    # the selection rule of HaloSpectrum.stochasticity: auto for equal biases, cross (sorted) otherwise
    assert np.array_equal(S.stochasticity(k, 0.6, 2.0, 2.0), auto(sigma8_z=0.6, b1=2.0, k=k))
    cross = S.CrossStochasticityFits()
    assert np.array_equal(S.stochasticity(k, 0.6, 2.0, 3.0), cross(sigma8_z=0.6, b1_1=2.0, b1_2=3.0, k=k))


def test_default_galaxy_model_uses_the_simulation_fits(golden):
    """GalaxySpectrum() with no stand-ins: b2_00 / b2_01 / stochasticity come from the Gaussian processes, as in the reference"""
    from pyrsd_b200 import rsd, simulation as S
    from tests.gpu_common import golden_pygcl_cosmology
    cosmo = golden_pygcl_cosmology(golden)
    m = rsd.GalaxySpectrum(params=cosmo, z=0.55)
    assert m.b2_00 is S.b2_00 and m.stochasticity is S.stochasticity
    k, mu = np.logspace(-2, np.log10(0.4), 25), np.linspace(0, 1, 5)
    P = m.power(k, mu)
    assert P.shape == (25, 5) and np.all(np.isfinite(P)) and np.all(P > 0)
    # a different nonlinear bias changes the answer: the fits really are read
    m2 = rsd.GalaxySpectrum(params=cosmo, z=0.55, b2_00=lambda b1: S.b2_00(b1) + 0.5, engine=m._engine)
    assert np.max(np.abs(m2.power(k, mu)/P - 1)) > 1e-4
