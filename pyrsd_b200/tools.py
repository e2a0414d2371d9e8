"""Bias <-> mass <-> velocity-dispersion relations of ``pyRSD/rsd/tools.py`` (lines 393-711) on the device.

``BiasToSigmaRelation(z, cosmo, interpolate)(sigma8_z, b1)`` is the constraint that sets ``sigma_sB`` in the
reference's default galaxy fit (``rsdfit/theory/gal_defaults.py:174``): halo bias -> halo mass by inverting the
Tinker et al. (2010) bias with Sigma(R) of the linear spectrum, then mass -> line-of-sight velocity dispersion
(Evrard et al. 2008).  The reference runs ``scipy.optimize.brentq`` around an adaptive Sigma(R) integral per call;
here any number of (sigma8_z, b1) pairs is one kernel launch (``prsd_bias_to_mass``, csrc/biasrel.cu)."""
import ctypes as C

import numpy as np

from . import _native as N
from . import pygcl

_vp = C.c_void_p
MASS_NORM = 1e13


def bias_Tinker(sigmas, delta_c=1.686, delta_halo=200):
    """Tinker et al. 2010 halo bias (rsd/tools.py:693-711)"""
    y = np.log10(delta_halo)
    A = 1. + 0.24*y*np.exp(-(4./y)**4)
    a = 0.44*y - 0.88
    B, b = 0.183, 1.5
    Cc = 0.019 + 0.107*y + 0.19*np.exp(-(4./y)**4)
    c = 2.4
    nu = delta_c/np.asarray(sigmas, dtype=float)
    return 1. - A*(nu**a)/(nu**a + delta_c**a) + B*nu**b + Cc*nu**c


def _lib():
    L = pygcl._L()
    if not getattr(L, '_tools_sigs', False):
        L.prsd_bias_to_mass.argtypes = [_vp, C.c_int, _vp, N._dp, N._dp, C.c_double, C.c_double, C.c_double, C.c_double, N._dp, _vp]
        L.prsd_bias_to_mass.restype = C.c_int
        L._tools_sigs = True
    return L


def bias_to_mass(spectra, rescale, b1, rho_bar, delta_halo=200., mass_lo=1e5, mass_hi=1e16, cmap=None, return_sigma=False):
    """batched root solve: halo masses [M_sun/h] of the pairs (rescale = sigma8_z / sigma8, b1); ``spectra`` is a
    ``pygcl._Spectra`` batch (walker w -> cosmology cmap[w])"""
    rescale = N.as_f64(rescale)
    b1 = N.as_f64(b1)
    rescale, b1 = [np.ascontiguousarray(a) for a in np.broadcast_arrays(rescale, b1)]
    nw = len(b1)
    out = np.empty(nw)
    cm = None if cmap is None else np.ascontiguousarray(cmap, dtype=np.int32)
    sig = np.empty((spectra.ncosmo, 1000)) if return_sigma else None
    N.check(_lib().prsd_bias_to_mass(spectra.ptr, nw, None if cm is None else cm.ctypes.data_as(_vp), rescale, b1, float(rho_bar),
                                     float(delta_halo), float(mass_lo), float(mass_hi), out,
                                     None if sig is None else sig.ctypes.data_as(_vp)))
    return (out, sig) if return_sigma else out


def _bilinear(grid0, grid1, values, x0, x1):
    """RegularGridInterpolator (rsd/_interpolate.py:120-203, linear); returns None outside the grid, where the
    reference raises InterpolationDomainError"""
    if not (grid0[0] <= x0 <= grid0[-1] and grid1[0] <= x1 <= grid1[-1]):
        return None
    idx, y = [], []
    for g, x in ((grid0, x0), (grid1, x1)):
        i = min(max(int(np.searchsorted(g, x)) - 1, 0), len(g) - 2)
        idx.append(i)
        y.append((x - g[i])/(g[i + 1] - g[i]))
    (i, j), (ys, yb) = idx, y
    v = values
    return (1 - ys)*(1 - yb)*v[i, j] + (1 - ys)*yb*v[i, j + 1] + ys*(1 - yb)*v[i + 1, j] + ys*yb*v[i + 1, j + 1]


class BiasToMassRelation(object):
    """mass [M_sun/h] <-> linear bias through the Tinker bias (rsd/tools.py:393-545)"""
    interpolation_grid = {'sigma8_z': np.linspace(0.3, 1.0, 100), 'b1': np.linspace(0.9, 8., 70)}

    def __init__(self, z, cosmo, interpolate=False):
        if not isinstance(cosmo, pygcl.Cosmology):
            raise TypeError("`cosmo` must be a pyrsd_b200.pygcl.Cosmology")
        self.z, self.cosmo, self.interpolate = z, cosmo, interpolate
        self.delta_halo = 200
        self._table = None
        self._plin = None

    @property
    def power_lin(self):
        """the ``pygcl.LinearPS`` at z = 0 whose Sigma(R) defines the relation"""
        if self._plin is None:
            self._plin = pygcl.LinearPS(self.cosmo, 0.)
        return self._plin

    def _solve(self, sigma8_z, b1, lo, hi):
        return bias_to_mass(self.power_lin._sp, np.asarray(sigma8_z, dtype=float)/self.cosmo.sigma8(), b1,
                            self.cosmo.rho_bar_z(0.), self.delta_halo, lo*MASS_NORM, hi*MASS_NORM,
                            cmap=np.zeros(np.broadcast(sigma8_z, b1).size, dtype=np.int32))

    @property
    def interpolation_table(self):
        """the 100 x 70 grid of masses (tools.py:463-507): all 7000 roots in one launch, bracket [1e-10, 1e4] x 1e13"""
        if self._table is None:
            s8, b1 = np.meshgrid(self.interpolation_grid['sigma8_z'], self.interpolation_grid['b1'], indexing='ij')
            self._table = self._solve(s8.ravel(), b1.ravel(), 1e-10, 1e4).reshape(s8.shape)
        return self._table

    def __call__(self, sigma8_z, b1):
        s8a, b1a = np.broadcast_arrays(np.asarray(sigma8_z, dtype=float), np.asarray(b1, dtype=float))
        scalar = s8a.ndim == 0
        s8f, b1f = s8a.ravel(), b1a.ravel()
        out = np.empty(len(b1f))
        fresh = np.ones(len(b1f), dtype=bool)
        if self.interpolate:
            g = self.interpolation_grid
            tab = self.interpolation_table
            for i in range(len(b1f)):
                v = _bilinear(g['sigma8_z'], g['b1'], tab, s8f[i], b1f[i])
                if v is not None:
                    out[i], fresh[i] = v, False
        if fresh.any():
            out[fresh] = self._solve(s8f[fresh], b1f[fresh], 1e-8, 1e3)         # compute_fresh (tools.py:519-531)
        return float(out[0]) if scalar else out.reshape(s8a.shape)


class BiasToSigmaRelation(BiasToMassRelation):
    """velocity dispersion [Mpc/h] <-> halo bias (rsd/tools.py:550-649)"""

    def __init__(self, z, cosmo, interpolate=False, sigmav_0=None, M_0=None):
        BiasToMassRelation.__init__(self, z, cosmo, interpolate=interpolate)
        self.sigmav_0, self.M_0 = sigmav_0, M_0

    @property
    def Hz(self):
        return self.cosmo.H_z(self.z)

    @property
    def Ez(self):
        return self.cosmo.H_z(self.z)/self.cosmo.H0()

    def evrard_sigmav(self, mass):
        """Evrard et al. 2008: sigma_DM(M, z) = 1082.9 km/s (E(z) M / 1e15)^0.336, line of sight, in Mpc/h"""
        sigmav = 1082.9*(self.Ez*np.asarray(mass, dtype=float)/1e15)**0.336
        sigmav = sigmav/3**0.5
        return sigmav*(1 + self.z)/self.Hz

    def mass(self, sigma8_z, b1):
        return BiasToMassRelation.__call__(self, sigma8_z, b1)

    def __call__(self, sigma8_z, b1):
        M = self.mass(sigma8_z, b1)
        if self.sigmav_0 is None or self.M_0 is None:
            return self.evrard_sigmav(M)
        return self.sigmav_0*(M/self.M_0)**(1./3.)


def mass_from_bias(bias, z, linearPS):
    """halo mass [M_sun / h] of a linear bias at redshift ``z`` from the Tinker et al. bias and Sigma(R) of ``linearPS``
    (a ``pygcl.LinearPS`` at any redshift), rsd/tools.py:655-691: the mean density is the z = 0 one, the growth between the
    spectrum's redshift and ``z`` rescales sigma, the root is bracketed by [1e-5, 1e5] x 1e13.  ``bias`` may be an array
    (one launch for all of them)."""
    cosmo = linearPS.GetCosmology()
    z_P = linearPS.GetRedshift()
    Dz = 1. if z_P == z else cosmo.D_z(z)/cosmo.D_z(z_P)
    b = np.asarray(bias, dtype=float)
    out = bias_to_mass(linearPS._sp, np.full(b.size, Dz), b.ravel(), cosmo.rho_bar_z(0.), 200., 1e-5*MASS_NORM, 1e5*MASS_NORM,
                       cmap=np.zeros(b.size, dtype=np.int32))
    return float(out[0]) if b.ndim == 0 else out.reshape(b.shape)


def sigma_from_bias(bias, z, linearPS):
    """sigma [Mpc/h] = 3.6 (M(bias) / 5.4903e13)^(1/3) (rsd/tools.py:713-720)"""
    return 3.6*(np.asarray(mass_from_bias(bias, z, linearPS))/5.4903e13)**(1./3)
