"""The rest of the reference's model-class surface on ``GalaxySpectrum`` / ``BiasedSpectrum`` / ``DarkMatterSpectrum``:

* the nine halo-halo terms ``P00_ss`` ... ``P04_ss`` (``biased/P00.py`` ... ``P22.py``, ``power_biased.py:368-448``) with their
  ``.mu0`` ... ``.mu6`` pieces, ``P00_ss_no_stoch``, ``Phm`` / ``Phm_bar``, the bias-dependent scalars (``bs``, ``sigmav_halo``,
  ``sigmav_from_bias``, ``mu2_model_correction`` ...) for the pair (``b1``, ``b1_bar``) of any of the three classes;
* the dark-matter term accessors on the biased classes (the reference's ``BiasedSpectrum`` IS a ``DarkMatterSpectrum``);
* the ten galaxy sub-spectra ``Pgal_cAcA`` ... ``Pgal_sBsB`` (``power_gal.py:292-399``), ``FOG``;
* the generic methods that only need ``power``: ``poles``, ``monopole`` ..., ``derivative_k`` / ``derivative_mu``,
  ``apply_transfer`` for ``DarkMatterSpectrum`` / ``BiasedSpectrum``;
* the state helpers ``preserve``, ``use_spt``, ``use_cache``, ``cache_override``, ``config`` and the plain attributes
  (``power_lin``, ``k_interp``, ``conformalH`` ...).

Every integral still comes from the device (PT tables of the plan, the blocks and moments of the assembly kernel); the
term accessors are the reference's Python algebra over those tables.  The sum of the terms reproduces the moments the
assembly kernel builds itself (``k_gal_moments``) -- tests/test_gpu_dm.py holds both to each other and to the reference's
Python.

Not mirrored (tied to ``pyRSD.rsdfit`` parameter sets or to the reference's cache machinery): ``default_params``,
``default_config``, ``get_gradient`` (``GalaxySpectrum.gradient`` is the device route), ``P11_sim_model``."""
import contextlib

import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline as _spline

from . import pygcl
from . import rsd as _rsd
from . import dm as _dm

HALO_TERMS = ('P00_ss', 'P00_ss_no_stoch', 'P01_ss', 'P11_ss', 'P02_ss', 'P12_ss', 'P22_ss', 'P03_ss', 'P13_ss', 'P04_ss')
DM_NAMES = (('P00', 'P01', 'P11', 'P02', 'P12', 'P22', 'P03', 'P13', 'P04', 'Pdd', 'Pdv', 'Pvv', 'sigmasq_k')
            + tuple('I%d%d' % (m, n) for m in range(4) for n in range(4)) + tuple('J%d%d' % mn for mn in _dm._J_ORDER)
            + tuple(_dm._K_ORDER) + tuple(_dm._ML_ORDER))


def dm_view(g):
    """a ``DarkMatterSpectrum`` over the tables of the assembly model ``g`` (no 2-loop terms: power_biased.py:47)"""
    v = object.__new__(_dm.DarkMatterSpectrum)
    d = v.__dict__
    d['_g'] = g
    d['include_2loop'] = False
    d['transfer_fit'] = g.transfer_fit
    d['max_mu'] = g.max_mu
    d['sigma_v2'], d['sigma_bv2'], d['sigma_bv4'] = g.sigma_v2, g.sigma_bv2, g.sigma_bv4
    d['_tabs'] = g.__dict__.get('_dm_tabs')
    return v


class HaloTerms(object):
    """the halo-halo terms of a pair of linear biases on the assembly model ``g``"""

    def __init__(self, g, b1, b1_bar):
        self.g, self.d = g, dm_view(g)
        if g.use_mean_bias:                                              # power_biased.py:219-237
            b1 = b1_bar = (b1*b1_bar)**0.5
        self.b1, self.b1_bar = float(b1), float(b1_bar)
        self.bs = -2./7*(self.b1 - 1.) if g.use_tidal_bias else 0.     # :239-257
        self.bs_bar = -2./7*(self.b1_bar - 1.) if g.use_tidal_bias else 0.
        self._base = None

    # ---- scalars
    def _sigv(self, b):
        s = self.g._sigmav_halo(b)                                       # :268-286
        if s > 0:
            return s
        return self.g.sigma_v if self.g.sigma_v is not None else self.g.sigma_lin

    @property
    def sigsq(self): return self._sigv(self.b1)**2

    @property
    def sigsq_bar(self): return self._sigv(self.b1_bar)**2

    def base(self):
        if self._base is None:
            g = self.g
            km = g.k
            eng, gal, _ = g._evaluate([g._state(km)], 'power', km[len(km)//2:len(km)//2 + 1], np.zeros(1))      # one point inside the grid whatever the AP parameters
            self._base = eng.galaxy_base(gal, 1)[0]
        return self._base

    def phm(self, b1, k):
        return _dm.phm_from_base(self.g, self.base(), b1, k)

    def stochasticity(self, k):
        g = self.g
        ks = _rsd.stochasticity_grid(g.k)
        lo, hi = sorted((self.b1, self.b1_bar))
        return _spline(ks, g.stochasticity(ks, g.sigma8_z, lo, hi))(np.asarray(k, dtype=float))

    # ---- building blocks shared by several terms
    def Phh_mu0(self, k):
        """biased/P00.py:3-18 (one b2_00 for every term: use_vlah_biasing, nonlinear_biasing.py:10-23)"""
        g, d = self.g, self.d
        b1, bb = self.b1, self.b1_bar
        b2, b2b = g.b2_00(b1), g.b2_00(bb)
        return b1*bb*d.P00.mu0(k) + (b1*b2b + bb*b2)*d.K00(k) + (self.bs*b2b + self.bs_bar*b2)*d.K00s(k)

    def P01_mu2(self, k):
        """biased/P01.py:3-26"""
        g, d = self.g, self.d
        b1, bb = self.b1, self.b1_bar
        b2, b2b = g.b2_01(b1), g.b2_01(bb)
        return (b1*bb*d.P01.mu2(k) - d.Pdv(k)*(b1*(1. - bb) + bb*(1. - b1))
                + g.f*((b2 + b2b)*d.K10(k) + (self.bs + self.bs_bar)*d.K10s(k))
                + g.f*((bb*b2 + b1*b2b)*d.K11(k) + (bb*self.bs + b1*self.bs_bar)*d.K11s(k)))

    # ---- the terms
    def term(self, name):
        g, d = self.g, self.d
        f = g.f
        b1, bb = self.b1, self.b1_bar
        A, T = _dm._Angular, _dm._Term
        arr = lambda k: np.asarray(k, dtype=float)
        if name == 'P00_ss_no_stoch':
            def mu0(k):
                P00 = d.P00.mu0(k)                                       # biased/P00.py:30-40
                return (self.phm(b1, k)/P00)*(self.phm(bb, k)/P00)*P00
            return T(mu0=A(total=mu0))
        if name == 'P00_ss':
            return T(mu0=A(total=lambda k: self.term('P00_ss_no_stoch').mu0(k) + self.stochasticity(k)))
        if name == 'P01_ss':
            return T(mu2=A(total=self.P01_mu2))
        if name == 'P11_ss':
            def mu2(k):
                return b1*bb*f**2*(d.Ivvdd_h01(k) + d.Idvdv_h03(k))    # biased/P11.py:3-16
            def mu4(k):
                return (0.5*(b1 + bb)*d.P11.mu4(k) - 0.5*((b1 - 1.) + (bb - 1.))*d.Pvv(k)
                        + f**2*(d.Ivvdd_h02(k) + d.Idvdv_h04(k))*(b1*bb - 0.5*(b1 + bb)))
            return T(mu2=A(total=mu2), mu4=A(total=mu4))
        if name == 'P02_ss':
            b2, b2b = g.b2_00(b1), g.b2_00(bb)
            def mu2(k):
                k = arr(k)                                               # biased/P02.py:4-32
                return (0.5*(b1 + bb)*d.P02.mu2.no_velocity(k) - 0.5*(f*k)**2*(self.sigsq + self.sigsq_bar)*self.Phh_mu0(k)
                        + 0.5*f**2*((b2 + b2b)*d.K20_a(k) + (self.bs + self.bs_bar)*d.K20s_a(k)))
            def mu4(k):
                return (0.5*(b1 + bb)*d.P02.mu4.no_velocity(k)
                        + 0.5*f**2*((b2 + b2b)*d.K20_b(k) + (self.bs + self.bs_bar)*d.K20s_b(k)))
            return T(mu2=A(total=mu2), mu4=A(total=mu4))
        if name == 'P12_ss':
            def mu4(k):
                k = arr(k)                                               # biased/P12.py:4-22
                P01 = d.P01.mu2(k)
                s = self.sigsq + self.sigsq_bar
                return (d.P12.mu4.no_velocity(k) - 0.25*(f*k)**2*s*P01 - 0.5*((b1 - 1.) + (bb - 1.))*f**3*d.I03(k)
                        - 0.25*(f*k)**2*s*(self.P01_mu2(k) - P01))
            def mu6(k):
                k = arr(k)
                return f**3*(d.I21(k) - 0.5*(b1 + bb)*d.I30(k) + 2*k**2*d.J20(k)*d.normed_power_lin(k))
            return T(mu4=A(total=mu4), mu6=A(total=mu6))
        if name == 'P22_ss':
            def mu4(k):
                k = arr(k)                                               # biased/P22.py:4-32
                s, sb = self.sigsq, self.sigsq_bar
                return (d.P22.mu4.no_velocity(k) + 0.5*(f*k)**4*(b1*bb*d.Pdd(k))*d.sigmasq_k(k)**2
                        - 0.25*(k*f)**2*(s + sb)*(0.5*(b1 + bb)*d.P02.mu2.no_velocity(k))
                        + 0.125*(k*f)**4*(s**2 + sb**2)*self.Phh_mu0(k))
            def mu6(k):
                k = arr(k)
                return (d.P22.mu6.no_velocity(k)
                        - 0.25*(k*f)**2*(self.sigsq + self.sigsq_bar)*(0.5*(b1 + bb)*d.P02.mu4.no_velocity(k)))
            return T(mu4=A(total=mu4), mu6=A(total=mu6))
        if name == 'P03_ss':
            return T(mu4=A(total=lambda k: -0.25*(f*arr(k))**2*(self.sigsq + self.sigsq_bar)*self.P01_mu2(k)))
        if name == 'P13_ss':
            P11 = self.term('P11_ss')                                    # biased/P13.py:3-34
            return T(mu4=A(total=lambda k: -0.5*(f*arr(k))**2*(self.sigsq + self.sigsq_bar)*P11.mu2(k)),
                     mu6=A(total=lambda k: -0.5*(f*arr(k))**2*(self.sigsq + self.sigsq_bar)*P11.mu4(k)))
        if name == 'P04_ss':
            def mu4(k):
                k = arr(k)                                               # biased/P04.py:4-26
                s, sb = self.sigsq, self.sigsq_bar
                return (-0.125*(b1 + bb)*(f*k)**2*(s + sb)*d.P02.mu2.no_velocity(k)
                        + (1./12)*(f*k)**4*self.Phh_mu0(k)*(1.5*(s**2 + sb**2) + g.velocity_kurtosis))
            def mu6(k):
                k = arr(k)
                return -0.125*(b1 + bb)*(f*k)**2*(self.sigsq + self.sigsq_bar)*d.P02.mu4.no_velocity(k)
            return T(mu4=A(total=mu4), mu6=A(total=mu6))
        raise AttributeError(name)

    def mu_correction(self, which, k):
        """mu2_model_correction / mu4_model_correction (power_biased.py:450-472): the residual fit at the mean bias"""
        from . import simulation as _sim
        g = self.g
        fit, cls = ((g.Pmu2_correction, _sim.Pmu2ResidualCorrection) if which == 2
                    else (g.Pmu4_correction, _sim.Pmu4ResidualCorrection))
        fit = fit if fit is not None else _sim._shared(cls)
        return np.asarray(fit(f=g.f, sigma8_z=g.sigma8_z, b1=(self.b1*self.b1_bar)**0.5, k=np.asarray(k, dtype=float)))


# ------------------------------------------------------------------ what the three classes share
def _halo(self):
    g = self if isinstance(self, _rsd.GalaxySpectrum) else self._g
    b1, b1_bar = (self.b1, self.b1_bar) if isinstance(self, _rsd.GalaxySpectrum) else self._pair()
    return HaloTerms(g, b1, b1_bar)


def _halo_attr(self, name):
    """names resolved through the pair's halo terms / the dark-matter view; raises AttributeError otherwise"""
    if name in HALO_TERMS:
        return _halo(self).term(name)
    if name in ('bs', 'bs_bar'):
        return getattr(_halo(self), name)
    if name == 'sigmav_halo':
        h = _halo(self)
        return h._sigv(h.b1)
    if name == 'sigmav_halo_bar':
        h = _halo(self)
        return h._sigv(h.b1_bar)
    if name == 'mu2_model_correction':
        return lambda k: _halo(self).mu_correction(2, k)
    if name == 'mu4_model_correction':
        return lambda k: _halo(self).mu_correction(4, k)
    g = self if isinstance(self, _rsd.GalaxySpectrum) else self._g
    if name in DM_NAMES:
        return getattr(dm_view(g), name)
    raise AttributeError(name)


def _generic_poles(self, k, ells, Nmu=40):
    """P_ell(k) = (2 ell + 1) int_0^1 L_ell(mu) power(k, mu) dmu by Simpson's rule on Nmu + 1 (odd) equally spaced mu
    (power_dm.py:1016-1044: the reference's rule on its mu grid); list of arrays, one per ell"""
    from scipy.special import legendre
    scalar = np.isscalar(ells)
    ells_ = [ells] if scalar else list(ells)
    k = np.atleast_1d(np.asarray(k, dtype=float))
    n = int(Nmu) + (1 - int(Nmu) % 2)
    mus = np.linspace(0., 1., n)
    P = np.atleast_2d(self.power(k, mus))                                # [nk][n]
    out = [pygcl.SimpsIntegrate(mus, (2*ell + 1.)*legendre(ell)(mus)[None, :]*P) for ell in ells_]
    return out[0] if scalar else out


def _generic_pole_method(ell):
    def method(self, k, **kwargs):
        return _generic_poles(self, k, ell, Nmu=41)
    return method


class _HZPTView(object):
    """``model.hzpt``: P00, P01, P11, Phm of the HZPT models (rsd/hzpt/__init__.py:185-287) at the model's sigma8_z, f"""

    def __init__(self, g):
        self._g = g

    def _block(self, name, k, extra=None):
        g = self._g
        saved = dict(g._models)
        try:
            for kw in ('use_P00_model', 'use_P01_model', 'use_P11_model'):
                g._models[kw] = True
            d = dm_view(g)
            y = d._block(name, k)
            if extra is not None:
                y = y + extra(d, np.asarray(k, dtype=float))
            return y
        finally:
            g._models.update(saved)

    def P00(self, k): return self._block('P00', k)
    def P01(self, k): return self._block('P01', k)
    # the kernel's block is hzpt.P11 - f^2 I31 (dm/P11.py:57-67)
    def P11(self, k): return self._block('P11_mu4', k, lambda d, q: d.f**2*d.I31(q))

    def Phm(self, b1, k):
        g = self._g
        saved = g._models['use_Phm_model']
        try:
            g._models['use_Phm_model'] = True
            return HaloTerms(g, b1, b1).phm(b1, k)
        finally:
            g._models['use_Phm_model'] = saved


def _jennings(self, which, k):
    g = self if isinstance(self, _rsd.GalaxySpectrum) else self._g
    saved = dict(g._models), g.Pdv_model_type, dict(g._loaded_data)
    try:
        g._models['use_Pdv_model'] = g._models['use_Pvv_model'] = True
        g.Pdv_model_type = 'jennings'
        g._loaded_data.pop('Pdv', None)
        return dm_view(g)._block(which, k)
    finally:
        g._models.update(saved[0])
        g.Pdv_model_type = saved[1]
        g._loaded_data.update(saved[2])


# ------------------------------------------------------------------ GalaxySpectrum
G = _rsd.GalaxySpectrum
G.b1 = G.b1_bar = 2.                       # power_biased.py:57-60
G.redshift_params = ()
G.cosmo_filename = G.linear_power_file = None
G.spline = staticmethod(_spline)
G.spline_kwargs = {}

_g_getattr = G.__getattr__


def _galaxy_getattr(self, name):
    try:
        return _g_getattr(self, name)
    except AttributeError:
        if name.startswith('_'):
            raise
        return _halo_attr(self, name)


G.__getattr__ = _galaxy_getattr
G.params = property(lambda self: self.cosmo)
G.k_interp = property(lambda self: self._ensure_tables().k_interp)
G.conformalH = property(lambda self: self.cosmo.H_z(self.z)/(1. + self.z))
G.fiducial_rs_drag = property(lambda self: self.cosmo.rs_drag())
G.transfer_fit_int = property(lambda self: getattr(pygcl.transfers, self.transfer_fit))
G.power_lin = property(lambda self: pygcl.LinearPS(self.cosmo, 0.))
G.power_lin_nw = property(lambda self: pygcl.LinearPS(self.cosmo.clone(tf=pygcl.transfers.EH_NoWiggle), 0.))
G.normed_power_lin = lambda self, k: self._power_norm*self.power_lin(np.asarray(k, dtype=float))
G.normed_power_lin_nw = lambda self, k: self._power_norm*self.power_lin_nw(np.asarray(k, dtype=float))
G.hzpt = property(lambda self: _HZPTView(self))
G.Pdv_jennings = lambda self, k: _jennings(self, 'Pdv', k)
G.Pvv_jennings = lambda self, k: _jennings(self, 'Pvv', k)
G.Phm = lambda self, k: _halo(self).phm(_halo(self).b1, k)
G.Phm_bar = lambda self, k: _halo(self).phm(_halo(self).b1_bar, k)


def _pair_moment(i):
    def method(self, k):
        v = object.__new__(_dm.BiasedSpectrum)
        v.__dict__.update(_g=self, b1=self.b1, b1_bar=self.b1_bar)
        return v._eval(i, k)
    return method


for _i in range(4):
    setattr(G, 'P_mu%d' % (2*_i), _pair_moment(_i))


def _config(self):
    """the model configuration (power_dm.py:775-784): the constructor keywords at their current values"""
    names = ('kmin', 'kmax', 'Nk', 'z', 'params', 'include_2loop', 'transfer_fit', 'max_mu', 'interpolate', 'k0_low',
             'Pdv_model_type', 'fog_model', 'use_so_correction', 'use_tidal_bias', 'use_mean_bias', 'vel_disp_from_sims',
             'correct_mu2', 'correct_mu4', 'use_vlah_biasing', 'linear_power_file', 'redshift_params', 'use_P00_model',
             'use_P01_model', 'use_P11_model', 'use_Pdv_model', 'use_Pvv_model', 'use_Phm_model')
    return {n: getattr(self, n) for n in names}


G.config = property(_config)


@contextlib.contextmanager
def _preserve(self, **kwargs):
    """save the model's state, apply ``kwargs``, restore on exit (power_dm.py:180-219)"""
    saved, models, loaded = dict(self.__dict__), dict(self._models), dict(self._loaded_data)
    try:
        self.update(**kwargs)
        yield
    finally:
        self.__dict__.clear()
        self.__dict__.update(saved)
        self._models.clear(); self._models.update(models)
        self._loaded_data.clear(); self._loaded_data.update(loaded)


@contextlib.contextmanager
def _use_spt(self):
    """every ``use_*_model`` off inside the context (power_dm.py:221-248)"""
    saved = dict(self._models)
    try:
        for kw in self._models:
            self._models[kw] = False
        yield
    finally:
        self._models.update(saved)


@contextlib.contextmanager
def _use_cache(self):
    """the reference memoises repeated (k, mu) calls here (power_dm.py:161-175); an evaluation is one kernel launch: no-op"""
    yield


@contextlib.contextmanager
def _cache_override(self, b2_00=None, b2_01=None, **kws):
    """constant b2_00 / b2_01 inside the context (power_biased.py:73-100)"""
    if kws:
        raise ValueError("only `b2_00` and `b2_01` can be overridden here")
    saved = self.b2_00, self.b2_01
    try:
        if b2_00 is not None:
            assert np.isscalar(b2_00)
            self.b2_00 = lambda b1: b2_00
        if b2_01 is not None:
            assert np.isscalar(b2_01)
            self.b2_01 = lambda b1: b2_01
        yield
    finally:
        self.b2_00, self.b2_01 = saved


G.preserve, G.use_spt, G.use_cache, G.cache_override = _preserve, _use_spt, _use_cache, _cache_override


def _fits(name):
    def get(self):
        from . import simulation as _sim
        return _sim._shared(getattr(_sim, name))
    return property(get)


G.auto_stochasticity_fits = _fits('AutoStochasticityFits')
G.cross_stochasticity_fits = _fits('CrossStochasticityFits')


def _bias_to_sigma(self):
    from . import tools as _tools
    key = (self.z, id(self.cosmo), self.interpolate)
    if self.__dict__.get('_b2s_key') != key:
        self.__dict__['_b2s'] = _tools.BiasToSigmaRelation(self.z, self.cosmo, interpolate=self.interpolate)
        self.__dict__['_b2s_key'] = key
    return self.__dict__['_b2s']


G.bias_to_sigma_relation = property(_bias_to_sigma)
G.sigmav_from_bias = lambda self, s8_z, bias: float(self.bias_to_sigma_relation(s8_z, bias))     # power_biased.py:288-306


def _pdv_sim_model(self):
    from . import simulation as _sim
    return _sim.SimulationPdv(self.cosmo, self.z, self.sigma8_z, self.f, self.Pdv_sim_cosmo)


G.Pdv_sim_model = property(_pdv_sim_model)


def _fog(self):
    """FOG(k, mu, sigma): the damping kernel of ``fog_model`` (gal/fog_kernels.py:20-84)"""
    model = self.fog_model

    def kernel(k, mu, sigma):
        x2 = (np.asarray(k, dtype=float)*np.asarray(mu, dtype=float)*sigma)**2
        if model == 'modified_lorentzian':
            return 1./(1. + 0.5*x2)**2
        if model == 'lorentzian':
            return 1./(1. + 0.5*x2)
        if model == 'gaussian':
            return np.exp(-0.5*x2)
        raise ValueError("unknown FOG model '%s'" % model)
    return kernel


G.FOG = property(_fog)

# ---- the ten sub-spectra: one batch of twelve walkers of the assembly kernel
_SUB = ('cAcA', 'cAcB', 'cBcB', 'cAsA', 'cAsB', 'cBsA', 'cBsB', 'sAsA', 'sAsB', 'sBsB')


def _subspectra(self, k, mu):
    """{name: P(k, mu)} of PcAcA ... PsBsB (power_gal.py:292-399): each G(sigma_1) G(sigma_2) [P_hh + 1-halo] + N with the
    Alcock-Paczynski treatment of ``power``.  ``power`` is a quadratic form in (fs, fcB, fsB) whose coefficients are these
    spectra: they are read off a batch of twelve walkers at fs, fcB, fsB in {0, 1/2, 1} -- one launch.  With the SO
    correction the three central-central spectra are undamped (sigma_c = 0 inside them, gal/Pcc/__init__.py:3-19)."""
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    if k.ndim == 1 and mu.ndim == 1 and len(mu) != len(k):
        k, mu = k[:, None], mu[None, :]
    kb, mub = np.broadcast_arrays(k, mu)
    saved = self.fs, self.fcB, self.fsB, self.sigma_c, self.f_so
    states = []
    try:
        # 0-2: the centrals alone, no SO re-damping; 3-4: the centrals as `power` sees them; 5-7: satellites; 8-11: mixed
        for fcB in (0., 0.5, 1.):
            self.fs, self.fcB, self.f_so = 0., fcB, 0.
            if self.use_so_correction:
                self.sigma_c = 0.
            states.append(self._state(self.k))
        self.sigma_c, self.f_so = saved[3], saved[4]
        for fcB in (0., 1.):
            self.fs, self.fcB = 0., fcB
            states.append(self._state(self.k))
        for fsB in (0., 0.5, 1.):
            self.fs, self.fsB = 1., fsB
            states.append(self._state(self.k))
        for fcB in (0., 1.):
            for fsB in (0., 1.):
                self.fs, self.fcB, self.fsB = 0.5, fcB, fsB
                states.append(self._state(self.k))
    finally:
        self.fs, self.fcB, self.fsB, self.sigma_c, self.f_so = saved
    P = self._evaluate(states, 'power', np.ascontiguousarray(kb).ravel(), np.ascontiguousarray(mub).ravel())[2]
    out = {'cAcA': P[0], 'cBcB': P[2], 'cAcB': 2.*P[1] - 0.5*(P[0] + P[2]),
           'sAsA': P[5], 'sBsB': P[7], 'sAsB': 2.*P[6] - 0.5*(P[5] + P[7])}
    cc = {0.: P[3], 1.: P[4]}
    ss = {0.: P[5], 1.: P[7]}
    i = 8
    for fcB, c in ((0., 'cA'), (1., 'cB')):
        for fsB, s in ((0., 'sA'), (1., 'sB')):
            out[c + s] = 2.*P[i] - 0.5*(cc[fcB] + ss[fsB])
            i += 1
    return {n: np.squeeze((v + self.N).reshape(kb.shape)) for n, v in out.items()}


G.subspectra = _subspectra
for _n in _SUB:
    setattr(G, 'Pgal_' + _n, (lambda n: lambda self, k, mu, flatten=False: _rsd._flat(_subspectra(self, k, mu)[n], flatten))(_n))

# ------------------------------------------------------------------ BiasedSpectrum / HaloSpectrum
B = _dm.BiasedSpectrum
_b_getattr = B.__getattr__


def _biased_getattr(self, name):
    if not name.startswith('_'):
        try:
            return _halo_attr(self, name)
        except AttributeError:
            pass
    return _b_getattr(self, name)


B.__getattr__ = _biased_getattr
B.poles = _generic_poles
B.monopole, B.quadrupole = _generic_pole_method(0), _generic_pole_method(2)
B.hexadecapole, B.tetrahexadecapole = _generic_pole_method(4), _generic_pole_method(6)
B.apply_transfer = _rsd._apply_transfer
B.derivative_k = lambda self, k, mu: _biased_derivative(self, k, mu, 'k')
B.derivative_mu = lambda self, k, mu: _biased_derivative(self, k, mu, 'mu')
B.to_npy = lambda self, filename: np.save(filename, np.array(self, dtype=object), allow_pickle=True)
B.from_npy = classmethod(lambda cls, filename: np.load(filename, allow_pickle=True).tolist())


def _biased_derivative(self, k, mu, wrt):
    # the AP parameters live on the wrapped assembly model: _derivative sets them on `self`
    class _Proxy(object):
        pass
    p = _Proxy()
    g = self._g
    p.__dict__.update(alpha_par=g.alpha_par, alpha_perp=g.alpha_perp, alpha_drag=g.alpha_drag)

    def power(kk, mm):
        saved = g.alpha_par, g.alpha_perp, g.alpha_drag
        try:
            g.alpha_par, g.alpha_perp, g.alpha_drag = p.alpha_par, p.alpha_perp, p.alpha_drag
            return self.power(kk, mm)
        finally:
            g.alpha_par, g.alpha_perp, g.alpha_drag = saved
    p.power = power
    return _rsd._derivative(p, k, mu, wrt)


# ------------------------------------------------------------------ DarkMatterSpectrum
D = _dm.DarkMatterSpectrum
_d_getattr = D.__getattr__
_D_FROM_G = ('params', 'k_interp', 'fiducial_rs_drag', 'transfer_fit_int', 'power_lin', 'power_lin_nw', 'normed_power_lin_nw',
             'hzpt', 'Pdv_jennings', 'Pvv_jennings', 'Pdv_sim_model', 'load_dm_sims', 'use_spt', 'use_cache', 'spline',
             'spline_kwargs', 'redshift_params', 'cosmo_filename', 'linear_power_file')


def _dm_getattr(self, name):
    if name in _D_FROM_G:
        return getattr(self.__dict__['_g'], name)
    return _d_getattr(self, name)


D.__getattr__ = _dm_getattr
D.poles = _generic_poles
D.monopole, D.quadrupole = _generic_pole_method(0), _generic_pole_method(2)
D.hexadecapole, D.tetrahexadecapole = _generic_pole_method(4), _generic_pole_method(6)
D.apply_transfer = _rsd._apply_transfer
D.derivative_k = lambda self, k, mu: _biased_derivative(self, k, mu, 'k')
D.derivative_mu = lambda self, k, mu: _biased_derivative(self, k, mu, 'mu')
D.to_npy = B.to_npy
D.from_npy = B.from_npy


def _d_config(self):
    names = ('kmin', 'kmax', 'Nk', 'z', 'params', 'include_2loop', 'transfer_fit', 'max_mu', 'interpolate', 'k0_low',
             'Pdv_model_type', 'linear_power_file', 'redshift_params', 'use_P00_model', 'use_P01_model', 'use_P11_model',
             'use_Pdv_model', 'use_Pvv_model')
    return {n: getattr(self, n) for n in names}


D.config = property(_d_config)


@contextlib.contextmanager
def _d_preserve(self, **kwargs):
    own = dict(self.__dict__)
    with self._g.preserve():
        try:
            self.update(**kwargs)
            yield
        finally:
            g = self.__dict__['_g']
            self.__dict__.clear()
            self.__dict__.update(own)
            self.__dict__['_g'] = g


D.preserve = _d_preserve
