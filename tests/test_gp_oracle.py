"""CPU: the GP restatement (oracle/gp_restate.py) against scikit-learn's GaussianProcessRegressor -- an independent
implementation of the same zero-mean GP with a fixed anisotropic squared-exponential kernel and per-point noise -- and
the shipped weight vectors (pyrsd_b200/data/simulation_fits.npz) against a fresh solve."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, 'pyrsd_b200', 'data', 'simulation_fits.npz')


def test_restatement_matches_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel
    from oracle import gp_restate as G
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 2, (60, 3))*np.array([1., 10., 0.1])
    y = np.sin(x[:, 0]) + 0.1*x[:, 1] + 30*x[:, 2]**2 + 5.0
    yerr = rng.uniform(0.01, 0.05, 60)
    theta = [2.5, 0.7, 1.3, 0.4]
    xp = rng.uniform(-1, 2, (25, 3))*np.array([1., 10., 0.1])
    mine = G.fit_predict(x, y, yerr, theta, xp)
    xm, xs = G.scaler(x)
    ym, ys = G.scaler(y)
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(theta[0], 'fixed')*RBF(np.sqrt(theta[1:]), 'fixed'), alpha=(yerr/ys)**2,
                                   optimizer=None, normalize_y=False)
    gpr.fit((x - xm)/xs, (y - ym)/ys)
    ref = ym + ys*gpr.predict((xp - xm)/xs)
    assert np.max(np.abs(mine - ref)) < 1e-9*np.max(np.abs(ref))


@pytest.mark.skipif(not os.path.exists(DATA), reason="simulation_fits.npz not generated")
def test_shipped_weights_solve_their_system():
    """alpha really is (K + diag)^-1 y for the small processes (the training errors are not shipped for the large ones:
    their residual is checked where /root/reference is present)"""
    from oracle import gp_restate as G
    d = np.load(DATA)
    names = sorted(set(k.split('/')[0] for k in d.files))
    assert set(names) >= {'vel_disp', 'b2_00_a', 'b2_01_a', 'auto_stochasticity', 'cross_stochasticity', 'Pmu2_correction', 'Pmu4_correction'}
    for n in names:
        x, alpha = d[n + '/x'], d[n + '/alpha']
        assert x.shape[0] == alpha.shape[0] and np.all(np.isfinite(alpha))
        assert x.shape[1] == len(d[n + '/metric']) == len(d[n + '/x_mean'])
    if not os.path.isdir('/root/reference/pyRSD/data/simulation_fits'):
        return
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'pyrsd_b200', 'data'))
    import make_simulation_fits as M
    fr = M.read_json_frame(os.path.join(M.SRC, 'nonlinear_biases_fits_runPB.json'))
    b1 = np.linspace(1.0, 5.0, 9)
    for name, theta in (('b2_00_a', [114.83271418, 45.4136823]), ('b2_01_a', [7.48162433, 0.87971185])):
        ref = G.fit_predict(fr['b1'], fr[name], fr[name + '_err'], theta, b1[:, None])
        X = (d[name + '/x'] - d[name + '/x_mean'])/d[name + '/x_scale']
        Xp = (b1[:, None] - d[name + '/x_mean'])/d[name + '/x_scale']
        mine = float(d[name + '/y_mean']) + float(d[name + '/y_scale'])*(G.kernel(Xp, X, float(d[name + '/amp']), d[name + '/metric'])@d[name + '/alpha'])
        assert np.max(np.abs(mine/ref - 1)) < 1e-9
    fr = M.read_frame(os.path.join(M.SRC, 'auto_stochasticity_runPB.pickle'))
    pts = np.array([[0.6, 2.0, 0.05], [0.75, 3.1, 0.2], [0.55, 1.5, 0.01]])
    ref = G.fit_predict(np.column_stack([fr['sigma8_z'], fr['b1'], fr['k']]), fr['y'], fr['y_err'],
                        [0.56384418, 0.76723559, 0.49555955, 5.7815692], pts)
    n = 'auto_stochasticity'
    X = (d[n + '/x'] - d[n + '/x_mean'])/d[n + '/x_scale']
    Xp = (pts - d[n + '/x_mean'])/d[n + '/x_scale']
    mine = float(d[n + '/y_mean']) + float(d[n + '/y_scale'])*(G.kernel(Xp, X, float(d[n + '/amp']), d[n + '/metric'])@d[n + '/alpha'])
    assert np.max(np.abs(mine/ref - 1)) < 1e-7
