"""Simulation-calibrated model ingredients of ``pyRSD/rsd/simulation.py`` (lines 16-379): Gaussian-process interpolation of
runPB simulation measurements -- halo velocity dispersion, the nonlinear biases b2_00 / b2_01, the auto / cross
stochasticity Lambda(k) and the P[mu^2] / P[mu^4] residual corrections.

The reference evaluates them with ``george`` (GP.compute + GP.predict, mean only); here the weight vector of every
process is precomputed (``data/make_simulation_fits.py`` -> ``data/simulation_fits.npz``) and a prediction is one
kernel-vector product on the device (``prsd_gp_predict``, csrc/gp.cu).  Same class names and call signatures
(keyword independent variables, arrays broadcast against each other)."""
import ctypes as C
import os

import numpy as np

from . import _native as N
from . import pygcl

_vp = C.c_void_p
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'simulation_fits.npz')
_cache = {}


def _lib():
    L = pygcl._L()
    if not getattr(L, '_gp_sigs', False):
        d = C.c_double
        L.prsd_gp_create.argtypes = [_vp, C.c_int, C.c_int, N._dp, N._dp, N._dp, N._dp, d, N._dp, d, d, C.POINTER(_vp)]
        L.prsd_gp_predict.argtypes = [_vp, C.c_int, _vp, _vp, C.c_int]
        L.prsd_gp_destroy.argtypes = [_vp]
        for n in ('prsd_gp_create', 'prsd_gp_predict', 'prsd_gp_destroy'):
            getattr(L, n).restype = C.c_int
        L._gp_sigs = True
    return L


def training_set(name):
    """the arrays of one process: x, x_mean, x_scale, y_mean, y_scale, amp, metric, alpha, independent"""
    if not os.path.exists(_DATA):
        raise ImportError("%s missing: run pyrsd_b200/data/make_simulation_fits.py in a container that holds the reference" % _DATA)
    d = np.load(_DATA)
    keys = [k for k in d.files if k.startswith(name + '/')]
    if not keys:
        raise KeyError("no Gaussian process called '%s'" % name)
    return {k.split('/', 1)[1]: d[k] for k in keys}


class _GP(N.Handle):
    _destroy = 'prsd_gp_destroy'


class GaussianProcessFit(object):
    """``GeorgeSimulationData`` (simulation.py:16-262): predicts ``dependent`` at the keyword independent variables"""
    name = None

    def __init__(self, name=None, context=None):
        self.name = name or self.name
        t = training_set(self.name)
        self.independent = [str(s) for s in t['independent']]
        self.x, self.alpha = np.ascontiguousarray(t['x'], dtype=float), np.ascontiguousarray(t['alpha'], dtype=float)
        self.x_mean, self.x_scale = N.as_f64(t['x_mean']), N.as_f64(t['x_scale'])
        self.metric = N.as_f64(t['metric'])
        self.amp, self.y_mean, self.y_scale = float(t['amp']), float(t['y_mean']), float(t['y_scale'])
        self._ctx = context
        self._h = None

    def _handle(self):
        if self._h is None:
            ctx = self._ctx if self._ctx is not None else N.default_context()
            p = _vp()
            n, nd = self.x.shape
            N.check(_lib().prsd_gp_create(ctx.ptr, n, nd, self.x, self.x_mean, self.x_scale, self.metric, self.amp, self.alpha,
                                          self.y_mean, self.y_scale, C.byref(p)))
            self._h = _GP(p)
        return self._h

    def __getstate__(self):
        d = dict(self.__dict__)
        d['_h'] = d['_ctx'] = None
        return d

    def predict(self, pts):
        """pts [npred][ndim] in the order of ``self.independent`` -> [npred]"""
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, len(self.independent))
        out = np.empty(len(pts))
        N.check(_lib().prsd_gp_predict(self._handle().ptr, len(pts), pts.ctypes.data_as(_vp), out.ctypes.data_as(_vp), 0))
        return out

    def __call__(self, *args, **indep_vars):
        if len(args):
            if len(args) == 1 and len(self.independent) == 1:
                indep_vars[self.independent[0]] = args[0]
            else:
                raise ValueError("please pass variables as keywords")
        for p in self.independent:
            if p not in indep_vars:
                raise ValueError("please specify the `%s` independent variable" % p)
        arrs = np.broadcast_arrays(*[np.asarray(indep_vars[k], dtype=float) for k in self.independent])
        scalar = arrs[0].ndim == 0
        out = self.predict(np.column_stack([a.ravel() for a in arrs]))
        return float(out[0]) if scalar else out.reshape(arrs[0].shape)


def _shared(cls):
    """one device copy of every training set per process (the reference builds its GP in a cached_property)"""
    if cls not in _cache:
        _cache[cls] = cls()
    return _cache[cls]


class VelocityDispersionFits(GaussianProcessFit):
    """halo velocity dispersion sigma_v / f in Mpc/h as a function of sigma8_z and b1 (simulation.py:313-329)"""
    name = 'vel_disp'


class AutoStochasticityFits(GaussianProcessFit):
    """auto stochasticity Lambda(sigma8_z, b1, k) (simulation.py:355-366)"""
    name = 'auto_stochasticity'


class CrossStochasticityFits(GaussianProcessFit):
    """cross stochasticity Lambda(sigma8_z, b1_1 <= b1_2, k) (simulation.py:368-379)"""
    name = 'cross_stochasticity'


class Pmu2ResidualCorrection(GaussianProcessFit):
    """P[mu^2] model residual (f, sigma8_z, b1, k) (simulation.py:299-311)"""
    name = 'Pmu2_correction'


class Pmu4ResidualCorrection(GaussianProcessFit):
    """P[mu^4] model residual (f, sigma8_z, b1, k) (simulation.py:285-297)"""
    name = 'Pmu4_correction'


class NonlinearBiasFits(object):
    """b2_00_{a,b,c,d}, b2_01_{a,b} as functions of b1 (``GeorgeSimulationDataSet``, simulation.py:264-283, 331-353):
    ``fits(b1=..., select='b2_00_a')`` returns the list of the selected predictions, as the reference does"""
    dependents = ['b2_00_a', 'b2_00_b', 'b2_00_c', 'b2_00_d', 'b2_01_a', 'b2_01_b']

    def __init__(self, context=None):
        self._data = {d: GaussianProcessFit(d, context=context) for d in self.dependents}

    def __call__(self, *args, **indep_vars):
        select = indep_vars.pop('select', None)
        if select is None:
            select = self.dependents
        elif isinstance(select, str):
            select = [select]
        else:
            raise ValueError("do not understand `select` keyword")
        return [self._data[par](*args, **indep_vars) for par in select]


def stochasticity(k, sigma8_z, b1_lo, b1_hi):
    """Lambda(k) of a (b1, b1_bar) pair the way ``HaloSpectrum.stochasticity`` picks its process (power_biased.py:357-364):
    the auto fit for equal biases, the cross fit (b1_1 <= b1_2) otherwise"""
    k = np.asarray(k, dtype=float)
    if b1_lo == b1_hi:
        return _shared(AutoStochasticityFits)(sigma8_z=sigma8_z, b1=b1_lo, k=k)
    return _shared(CrossStochasticityFits)(sigma8_z=sigma8_z, b1_1=b1_lo, b1_2=b1_hi, k=k)


def b2_00(b1):
    """b2_00_a(b1), the single b2_00 of ``use_vlah_biasing`` (nonlinear_biasing.py:10-23)"""
    return float(np.squeeze(_shared(NonlinearBiasFits)(b1=b1, select='b2_00_a')[0]))


def b2_01(b1):
    return float(np.squeeze(_shared(NonlinearBiasFits)(b1=b1, select='b2_01_a')[0]))


# ------------------------------------------------------------------------- dark-matter simulation measurements
_DM_SIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'dm_sims.npz')


def dm_sims():
    """the measurements of ``pyRSD/data/dark_matter`` (see data/make_dm_sims.py): k, z, Pdv_mu0 / P00_mu0 / P01_mu2 / P11_mu4 [3][500]"""
    if 'dm_sims' not in _cache:
        _cache['dm_sims'] = dict(np.load(_DM_SIMS))
    return _cache['dm_sims']


def _lcdm_growth(Om, z):
    """D(z) / D(0) and f(z) of flat LCDM (growth integral; matter + Lambda)"""
    from scipy.integrate import quad
    E = lambda a: np.sqrt(Om/a**3 + 1. - Om)
    I = lambda a: quad(lambda x: 1./(x*E(x))**3, 0., a)[0]
    a = 1./(1. + z)
    D = E(a)*I(a)/(E(1.)*I(1.))
    f = -1.5*Om/a**3/E(a)**2 + a/((a*E(a))**3*I(a))
    return D, f


class SimulationPdv(object):
    """``SimulationPdv`` (rsd/simulation.py:550-660): the density -- velocity-divergence spectrum from the simulation
    measurements at z = 0, 0.509, 0.989, each divided by D(z)^2 f(z) P_lin(k) of the simulation cosmology, interpolated
    bilinearly in (z, k) (scipy RegularGridInterpolator, as the reference) and rescaled by f sigma8_z^2 P_lin(k) / sigma8^2
    of the model cosmology; outside the redshift range the nearest snapshot's shape is rescaled by ``z`` itself, as the
    reference's ``_extrapolate`` does.

    ``sim_cosmo``: a ``pygcl.Cosmology`` of data/params/teppei_sims.ini with ``growth`` / ``f`` at the three redshifts (the
    reference obtains it from CLASS).  None: the z = 0 linear spectrum the reference ships beside the measurements
    (data/dark_matter/Plin_z_0.000.dat, log-log spline) with the flat-LCDM growth integral -- CLASS is not part of this
    library, so this default normalisation is not pinned against the reference's."""

    def __init__(self, cosmo, z, sigma8_z, f, sim_cosmo=None):
        from scipy.interpolate import RegularGridInterpolator, InterpolatedUnivariateSpline
        self.cosmo, self.z, self.sigma8_z, self.f = cosmo, z, sigma8_z, f
        d = dm_sims()
        self.zs, self.ks = d['z'], d['k']
        if sim_cosmo is not None:
            Plin = pygcl.LinearPS(sim_cosmo, 0.)(self.ks)
            growth = [(sim_cosmo.D_z(zz), sim_cosmo.f_z(zz)) for zz in self.zs]
        else:
            spl = InterpolatedUnivariateSpline(np.log(d['plin_k']), np.log(d['plin_P']), k=3)
            k0, ns = d['plin_k'][0], float(d['sim_n_s'])
            # below the first tabulated k the spectrum is the primordial power law
            Plin = np.where(self.ks >= k0, np.exp(spl(np.log(np.maximum(self.ks, k0)))), d['plin_P'][0]*(self.ks/k0)**ns)
            Om = float(d['sim_Omega_b']) + float(d['sim_Omega_cdm'])
            growth = [_lcdm_growth(Om, zz) for zz in self.zs]
        vals = np.array([d['Pdv_mu0'][i]/(D**2*Plin*ff) for i, (D, ff) in enumerate(growth)])
        self.interpolation_table = RegularGridInterpolator((self.zs, self.ks), vals)
        self._power_lin = pygcl.LinearPS(cosmo, 0.)

    def __call__(self, k):
        k = np.atleast_1d(np.asarray(k, dtype=float))
        normed = self._power_lin(k)/self.cosmo.sigma8()**2
        if self.z < self.zs.min() or self.z > self.zs.max():
            zz, factor = (self.zs.min() if self.z < self.zs.min() else self.zs.max()), self.z*normed      # simulation.py:620-641
        else:
            zz, factor = self.z, self.f*self.sigma8_z**2*normed
        return self.interpolation_table(np.column_stack([np.full(len(k), zz), k]))*factor
