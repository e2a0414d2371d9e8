"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the Gaussian-process prediction the reference obtains from ``george``
(pyRSD/rsd/simulation.py:16-262): StandardScaler on inputs and output, ExpSquaredKernel with a diagonal metric
(k = exp(-r^2 / 2), r^2 = sum_d dx_d^2 / metric_d; george/kernels.py), measurement errors on the diagonal
(GP.compute(x, yerr)), zero-mean prediction K(x*, X) (K + diag(yerr^2))^-1 y (GP.predict, mean only).

``george`` is absent from this image and its version is not pinned by the reference (setup.py asks for ``george`` only),
so this stage is "parity unpinned" against george itself; it IS pinned against an independent implementation of the
same algorithm, scikit-learn's GaussianProcessRegressor with a fixed anisotropic RBF kernel (tests/test_gp_oracle.py).
Only tests/ may import this."""
import numpy as np


def scaler(a):
    a = np.asarray(a, dtype=float)
    m, s = a.mean(axis=0), a.std(axis=0)
    return m, np.where(s == 0, 1.0, s)


def kernel(A, B, amp, metric):
    Za, Zb = A/np.sqrt(metric), B/np.sqrt(metric)
    r2 = np.sum(Za*Za, axis=1)[:, None] + np.sum(Zb*Zb, axis=1)[None, :] - 2.0*Za@Zb.T
    return amp*np.exp(-0.5*np.maximum(r2, 0.0))


def fit_predict(x, y, yerr, theta, xpred):
    """mean prediction of the reference's GeorgeSimulationData at raw points xpred [npred][ndim]"""
    x = np.asarray(x, dtype=float).reshape(len(y), -1)
    nd = x.shape[1]
    theta = np.asarray(theta, dtype=float)
    amp, metric = (1.0, theta) if len(theta) == nd else (theta[0], theta[1:])
    xm, xs = scaler(x)
    ym, ys = scaler(np.asarray(y, dtype=float))
    X, Y = (x - xm)/xs, (np.asarray(y, dtype=float) - ym)/ys
    K = kernel(X, X, amp, metric) + np.diag((np.asarray(yerr, dtype=float)/ys)**2)
    alpha = np.linalg.solve(K, Y)
    Xp = (np.asarray(xpred, dtype=float).reshape(-1, nd) - xm)/xs
    return ym + ys*(kernel(Xp, X, amp, metric)@alpha)
