"""``ExtrapolatedPowerSpectrum`` and ``SmoothedXiMultipoles`` (``pyRSD/rsd/power_extrapolator.py:12-316``,
``pyRSD/rsd/correlation.py:14-68``): the correlation-function multipoles of a P(k, mu) model, obtained by extending the
model with power laws below ``k_lo`` and above ``k_hi``, integrating it to multipoles on 1000 log-spaced k in
[1e-5, 10] and Fourier transforming each with a Gaussian smoothing kernel.

The model evaluations go through ``model.power`` (one launch of the assembly kernel per call: the 200 x 100 (k, mu) grid
of the fit, the 1000 x 41 grid of the multipoles) and the transforms through the device FFTLog (``pygcl.pk_to_xi`` on a
batch of multipoles); the extrapolation parameters are two small least-squares fits per mu (``scipy.optimize.curve_fit``,
as the reference) on the host.

Reference quirk: ``ExtrapolatedPowerSpectrum`` takes a ``model_func`` name but then calls ``self.model.Pgal(k, mu)``
(power_extrapolator.py:173, 270-281), a method the current ``GalaxySpectrum`` no longer has (it is ``power``); here
``model_func`` (default ``'power'``) is what is called."""
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline as _spline
import scipy.optimize as _opt

from . import pygcl

KMAX = 10.          # correlation.py:12


def power_law_extrapolation(x, alpha, beta=0.):
    """x^(alpha + beta log10 x) (power_extrapolator.py:12-29)"""
    return x**(alpha + beta*np.log10(x))


def _bilinear(ks, mus, values, k, mu):
    """RegularGridInterpolator(linear) of rsd/_interpolate.py:120-203 on a (k, mu) grid at paired points"""
    k, mu = np.atleast_1d(np.asarray(k, dtype=float)), np.atleast_1d(np.asarray(mu, dtype=float))
    if (k < ks[0]).any() or (k > ks[-1]).any() or (mu < mus[0]).any() or (mu > mus[-1]).any():
        raise ValueError("One of the requested xi is out of bounds")
    i = np.clip(np.searchsorted(ks, k) - 1, 0, len(ks) - 2)
    j = np.clip(np.searchsorted(mus, mu) - 1, 0, len(mus) - 2)
    tk = (k - ks[i])/(ks[i + 1] - ks[i])
    tm = (mu - mus[j])/(mus[j + 1] - mus[j])
    return (values[i, j]*(1 - tk)*(1 - tm) + values[i + 1, j]*tk*(1 - tm) + values[i, j + 1]*(1 - tk)*tm
            + values[i + 1, j + 1]*tk*tm)


class ExtrapolatedPowerSpectrum(object):
    """A P(k, mu) model extended by a constant-slope power law below ``k_lo`` and a running-slope power law above
    ``k_hi``, both fitted per mu (power_extrapolator.py:32-316)"""

    def __init__(self, model, model_func='power', **kwargs):
        self.model = model
        self.model_func = model_func
        self.model_kmin, self.model_kmax = model.kmin, model.kmax
        self.k_hi = kwargs.get('k_hi', model.kmax)
        self.k_lo = kwargs.get('k_lo', model.kmin)
        self.kcut_hi = kwargs.get('kcut_hi', 0.5*self.k_hi)
        self.kcut_lo = kwargs.get('kcut_lo', 5.*self.k_lo)
        self._fit = None

    def _P(self, k, mu):
        return np.asarray(getattr(self.model, self.model_func)(k, mu))

    @property
    def _ks(self):
        return np.logspace(np.log10(self.model_kmin*1.05), np.log10(self.model_kmax*0.95), 200)

    @property
    def _mus(self):
        return np.linspace(0, 1., 100)

    def _fits(self):
        """the power on the 200 x 100 fit grid (one launch) and the splines of the fitted exponents vs mu"""
        if self._fit is None:
            ks, mus = self._ks, self._mus
            power = self._P(ks, mus)                                          # [200][100]
            amp_hi = _bilinear(ks, mus, power, np.full(len(mus), self.k_hi), mus) if ks[0] <= self.k_hi <= ks[-1] else None
            amp_lo = _bilinear(ks, mus, power, np.full(len(mus), self.k_lo), mus) if ks[0] <= self.k_lo <= ks[-1] else None
            if amp_hi is None or amp_lo is None:
                # the reference reads the amplitudes off its (k, mu) interpolation grid, which spans [1.05 kmin, 0.95 kmax]:
                # with the default k_lo = kmin / k_hi = kmax it raises the interpolator's out-of-bounds ValueError
                raise ValueError("One of the requested xi is out of bounds: k_lo / k_hi must lie inside "
                                 "[1.05 model.kmin, 0.95 model.kmax] (power_extrapolator.py:146-151, 187, 213)")
            hi_mask, lo_mask = ks >= self.kcut_hi, ks <= self.kcut_lo
            hi, lo = [], []
            for i in range(len(mus)):
                a = amp_hi[i]
                popt, _ = _opt.curve_fit(lambda x, alpha, beta: a*power_law_extrapolation(x, alpha, beta), ks[hi_mask]/self.k_hi,
                                         power[hi_mask, i])
                hi.append(popt)
                a = amp_lo[i]
                popt, _ = _opt.curve_fit(lambda x, alpha: a*power_law_extrapolation(x, alpha), ks[lo_mask]/self.k_lo, power[lo_mask, i])
                lo.append(popt)
            hi, lo = np.asarray(hi), np.asarray(lo)
            self._fit = dict(hi=[_spline(mus, p) for p in hi.T], lo=_spline(mus, lo[:, 0]))
        return self._fit

    def high_k_params(self, mu):
        return tuple(f(mu) for f in self._fits()['hi'])

    def low_k_params(self, mu):
        return self._fits()['lo'](mu)

    def __call__(self, k, mu):
        k = np.atleast_1d(np.asarray(k, dtype=float))
        mu = np.atleast_1d(np.asarray(mu, dtype=float))
        if k.ndim == 1 and mu.ndim == 1 and len(k) != len(mu):
            k, mu = k[:, None], mu[None, :]
        k, mu = np.broadcast_arrays(k, mu)
        out = np.empty(k.shape)
        lo, hi = k < self.k_lo, k > self.k_hi
        mid = ~(lo | hi)
        if lo.any():
            out[lo] = self._P(np.full(lo.sum(), self.k_lo), mu[lo])*power_law_extrapolation(k[lo]/self.k_lo, self.low_k_params(mu[lo]))
        if mid.any():
            out[mid] = self._P(k[mid], mu[mid])
        if hi.any():
            alpha, beta = self.high_k_params(mu[hi])
            out[hi] = self._P(np.full(hi.sum(), self.k_hi), mu[hi])*power_law_extrapolation(k[hi]/self.k_hi, alpha, beta)
        return np.squeeze(out)

    def to_poles(self, k, ells, Nmu=41, flatten=False):
        """multipoles by Simpson's rule over Nmu mu in [0, 1] (power_extrapolator.py:289-316) -> [len(k)][len(ells)]"""
        from scipy.special import legendre
        scalar = np.isscalar(ells)
        ells = [ells] if scalar else list(ells)
        mus = np.linspace(0., 1., Nmu)
        pkmu = np.atleast_2d(self(k, mus))                                    # [nk][Nmu]
        kern = np.array([(2*ell + 1.)*legendre(ell)(mus) for ell in ells])    # [nell][Nmu]
        vals = pygcl.SimpsIntegrate(mus, (kern[:, None, :]*pkmu[None, :, :]).reshape(-1, Nmu)).reshape(len(ells), -1)
        if scalar:
            return vals[0]
        out = vals.T
        return out if not flatten else np.ravel(out, order='F')


def SmoothedXiMultipoles(power_model, r, ells, R=0., **kwargs):
    """Xi_ell(s) = i^ell int dq / (2 pi^2) q^2 W(q R) P_ell(q) j_ell(q s) (correlation.py:14-68) -> [len(r)][len(ells)]"""
    ells = [ells] if np.isscalar(ells) else list(ells)
    extrap = ExtrapolatedPowerSpectrum(power_model, **kwargs)
    k_spline = np.logspace(-5, np.log10(KMAX), 1000)
    poles = extrap.to_poles(k_spline, ells)                                   # [1000][nell]
    xi = [pygcl.pk_to_xi(int(ell), k_spline, poles[:, i], r, smoothing=R) for i, ell in enumerate(ells)]
    return np.squeeze(np.vstack(xi).T)
