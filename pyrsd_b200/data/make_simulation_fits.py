"""Convert the reference's simulation training data (pyRSD/data/simulation_fits/*.pickle|json) into
``pyrsd_b200/data/simulation_fits.npz`` and precompute the Gaussian-process weight vectors.

The reference interpolates simulation measurements with ``george`` Gaussian processes (pyRSD/rsd/simulation.py:16-256):
zero-mean GP, ExpSquaredKernel with a diagonal metric theta (x an amplitude when theta has one entry more than the
inputs), inputs and output standardised with sklearn's StandardScaler, the measurement errors on the diagonal, mean
prediction only:   y(x*) = y_mean + y_scale * k(x*, X) (K(X, X) + diag(yerr_scaled^2))^-1 y_scaled.
The weight vector alpha = (K + diag)^-1 y_scaled depends on the training data only; it is solved here once (Cholesky,
scipy) and shipped, so a prediction is a single kernel-vector product on the device (csrc/gp.cu).

The pickles are pandas DataFrames written by pandas 0.1x; they are read with a permissive unpickler that keeps the
block-manager state (axes, block values, block items) and never imports the old pandas classes.

Run in the build container only:   python pyrsd_b200/data/make_simulation_fits.py   (about a minute)
"""
import json
import os
import pickle
import sys
import time

import numpy as np
from scipy.linalg import cho_factor, cho_solve

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/pyRSD/data/simulation_fits'


class _Generic(object):
    def __new__(cls, *a, **k):
        o = object.__new__(cls)
        o.args, o.kw, o.state = a, k, None
        return o

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.state = s


class _Unpickler(pickle.Unpickler):
    _cache = {}

    def find_class(self, module, name):
        if module.startswith('pandas'):
            key = module + '.' + name
            if key not in self._cache:
                if name[0].islower() or name.startswith('_'):
                    self._cache[key] = lambda *a, **k: _Generic(*a, **k)
                else:
                    self._cache[key] = type(key.replace('.', '_'), (_Generic,), {})
            return self._cache[key]
        return pickle.Unpickler.find_class(self, module, name)


def _index_values(ix):
    """labels of a pickled Index (the `data` entry of its constructor keywords)"""
    for a in ix.args:
        if isinstance(a, dict) and 'data' in a:
            return np.asarray(a['data'])
    raise ValueError("index without data")


def read_frame(path):
    """old pandas DataFrame pickle -> dict column -> 1-D array"""
    o = _Unpickler(open(path, 'rb'), encoding='latin1').load()
    axes, blocks, items = o.state.state[0], o.state.state[1], o.state.state[2]
    out = {}
    for vals, names in zip(blocks, items):
        names = [n.decode() if isinstance(n, bytes) else str(n) for n in _index_values(names)]
        for j, n in enumerate(names):
            col = np.asarray(vals[j])
            if col.dtype == object:
                col = np.array([float(x.decode() if isinstance(x, bytes) else x) for x in col])
            out[n] = col
    return out


def read_json_frame(path):
    d = json.load(open(path))
    n = len(next(iter(d.values())))
    return {k: np.array([float(v[str(i)]) for i in range(n)]) for k, v in d.items()}


def standard_scaler(a):
    """sklearn.preprocessing.StandardScaler: mean, population std (0 -> 1)"""
    a = np.asarray(a, dtype=float)
    m, s = a.mean(axis=0), a.std(axis=0)
    return m, np.where(s == 0, 1.0, s)


def fit_gp(data, independent, dependent, theta, use_errors=True):
    x = np.column_stack([data[c] for c in independent])
    y = np.asarray(data[dependent], dtype=float)
    xm, xs = standard_scaler(x)
    ym, ys = standard_scaler(y)
    X = (x - xm)/xs
    Y = (y - ym)/ys
    theta = np.asarray(theta, dtype=float)
    nd = x.shape[1]
    if len(theta) == nd:
        amp, metric = 1.0, theta
    elif len(theta) == nd + 1:
        amp, metric = theta[0], theta[1:]
    else:
        raise ValueError("size mismatch between supplied `x` variables and `theta` length")
    t0 = time.time()
    # K = amp exp(-1/2 sum_d dx_d^2 / metric_d)
    Z = X/np.sqrt(metric)
    sq = np.sum(Z*Z, axis=1)
    K = sq[:, None] + sq[None, :] - 2.0*(Z@Z.T)
    np.maximum(K, 0.0, out=K)
    np.exp(-0.5*K, out=K)
    K *= amp
    if use_errors:
        K[np.diag_indices_from(K)] += (np.asarray(data[dependent + '_err'], dtype=float)/ys)**2
    alpha = cho_solve(cho_factor(K, lower=True, overwrite_a=True, check_finite=False), Y, check_finite=False)
    print('  %-10s n = %5d, %d inputs: solved in %.1f s' % (dependent, len(y), nd, time.time() - t0), flush=True)
    return dict(x=x, x_mean=xm, x_scale=xs, y_mean=ym, y_scale=ys, amp=amp, metric=metric, alpha=alpha,
                independent=np.array(independent))


def main():
    out = {}

    def put(name, gp):
        for k, v in gp.items():
            out[name + '/' + k] = v

    # VelocityDispersionFits (simulation.py:313-329)
    d = read_frame(os.path.join(SRC, 'runPB_vel_disp.pickle'))
    d['sigma_v'] = d['sigma_v']/d['f']
    d['sigma_v_err'] = 1e-5*d['sigma_v']
    put('vel_disp', fit_gp(d, ['sigma8_z', 'b1'], 'sigma_v', [0.48087061, 0.21521814, 0.45073149]))
    # NonlinearBiasFits (simulation.py:331-353)
    d = read_json_frame(os.path.join(SRC, 'nonlinear_biases_fits_runPB.json'))
    thetas = {'b2_00_a': [114.83271418, 45.4136823], 'b2_00_b': [78.74803514, 13.7686103], 'b2_00_c': [519.43798005, 73.94048349],
              'b2_00_d': [31.32395946, 11.08334321], 'b2_01_a': [7.48162433, 0.87971185], 'b2_01_b': [14.06494712, 12.08894434]}
    for name, th in thetas.items():
        put(name, fit_gp(d, ['b1'], name, th))
    # AutoStochasticityFits / CrossStochasticityFits (simulation.py:355-379)
    d = read_frame(os.path.join(SRC, 'auto_stochasticity_runPB.pickle'))
    put('auto_stochasticity', fit_gp(d, ['sigma8_z', 'b1', 'k'], 'y', [0.56384418, 0.76723559, 0.49555955, 5.7815692]))
    d = read_frame(os.path.join(SRC, 'cross_stochasticity_runPB.pickle'))
    put('cross_stochasticity', fit_gp(d, ['sigma8_z', 'b1_1', 'b1_2', 'k'], 'y', [2.42796735, 0.59461745, 3.75844384, 1.5391186, 11.53257307]))
    # Pmu2 / Pmu4 residual corrections (simulation.py:285-311)
    d = read_frame(os.path.join(SRC, 'Pmu2_residual_data.pickle'))
    put('Pmu2_correction', fit_gp(d, ['f', 'sigma8_z', 'b1', 'k'], 'y', [11.84812224, 4.15569036, 1.26297742, 1.03950439]))
    d = read_frame(os.path.join(SRC, 'Pmu4_residual_data.pickle'))
    put('Pmu4_correction', fit_gp(d, ['f', 'sigma8_z', 'b1', 'k'], 'y', [33.66747949, 3.95336447, 1.74027224, 0.62058417]))
    path = os.path.join(HERE, 'simulation_fits.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f kB' % (os.path.getsize(path)/1e3))


if __name__ == '__main__':
    main()
