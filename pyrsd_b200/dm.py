"""``DarkMatterSpectrum`` and ``HaloSpectrum`` with the term accessors of the reference
(``pyRSD/rsd/power/dm/power_dm.py:26-1221``, ``dm/P00.py`` .. ``P22.py``; ``biased/power_biased.py:26-560``): the classes
the reference's own tests plot (``tests/test_vlahetal_dm.py``: ``model.P00.mu0(model.k)``, ``model.P02.mu2.no_velocity(k)``,
``model.P11.mu4.scalar(k)`` ...; ``tests/test_vlahetal_halo.py``: ``model.P_mu0(k)``, ``model.Phm(k)``).

Everything that is an integral comes from the device: the PT tables of the plan (I_mn, J_mn, K_mn, ImnOneLoop, sigma_v^2(k),
one-loop spectra, velocity kurtosis) and the dark-matter blocks of the assembly kernel's first stage (``k_gal_base``: HZPT +
Zel'dovich P00 / P01 / P11, the Jennings et al. P_dv / P_vv) for ``DarkMatterSpectrum``; the second stage (``k_gal_moments``:
the four moment splines of a (b1, b1_bar) pair) for ``HaloSpectrum``.  What is done here is what the reference does in
Python above its ``interpolated_function`` tables: splines in k and the products with the velocity-dispersion scalars."""
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline as _spline

from . import pygcl
from . import rsd as _rsd

_J_ORDER = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 0): 3, (1, 1): 4, (2, 0): 5}
_K_ORDER = ['K00', 'K01', 'K00s', 'K01s', 'K02s', 'K10', 'K11', 'K10s', 'K11s', 'K20_a', 'K20_b', 'K20s_a', 'K20s_b']
_ML_ORDER = ['Ivvdd_h01', 'Ivvdd_h02', 'Idvdv_h03', 'Idvdv_h04', 'Ivvvv_f23', 'Ivvvv_f32', 'Ivvvv_f33']


class _Angular(object):
    """one angular piece of a power term (``AngularTerm``, rsd/power/__init__.py:22-49): callable = ``total``"""

    def __init__(self, **parts):
        self._parts = parts
        for name, fn in parts.items():
            setattr(self, name, fn)

    def __call__(self, k):
        return self.total(k)


class _Term(object):
    def __init__(self, **angular):
        for name, a in angular.items():
            setattr(self, name, a)


class DarkMatterSpectrum(object):
    """The dark matter redshift-space power spectrum in the distribution-function expansion (power_dm.py:26-1221).

    ``params`` must be a :class:`pyrsd_b200.pygcl.Cosmology`.  Same attribute names and defaults as the reference where
    they enter the spectrum."""

    def __init__(self, kmin=1e-3, kmax=0.5, Nk=200, z=0., params=None, include_2loop=False, transfer_fit="CLASS",
                 max_mu=4, interpolate=True, k0_low=5e-3, Pdv_model_type='jennings', engine=None, **kwargs):
        self._g = _rsd.GalaxySpectrum(kmin=kmin, kmax=kmax, Nk=Nk, z=z, params=params, max_mu=max_mu if max_mu <= 6 else 6,
                                      interpolate=interpolate, k0_low=k0_low, Pdv_model_type=Pdv_model_type, engine=engine,
                                      b2_00=lambda b: 0., b2_01=lambda b: 0., stochasticity=lambda k, s8, a, b: 0.*k, **kwargs)
        self.include_2loop = bool(include_2loop)
        self.transfer_fit = transfer_fit
        self.max_mu = max_mu
        self.sigma_v2 = self.sigma_bv2 = self.sigma_bv4 = 0.
        self._tabs = None

    # ------------------------------------------------------------ parameters shared with the assembly model
    _FORWARD = ('cosmo', 'z', 'kmin', 'kmax', 'Nk', 'k', 'sigma8_z', 'f', 'alpha_par', 'alpha_perp', 'alpha_drag', 'sigma_v',
                'interpolate', 'k0_low', 'D', '_power_norm', 'sigma_lin', 'velocity_kurtosis', 'Pdv_model_type',
                'use_P00_model', 'use_P01_model', 'use_P11_model', 'use_Pdv_model', 'use_Pvv_model')

    def __getattr__(self, name):
        if name in DarkMatterSpectrum._FORWARD:
            return getattr(self.__dict__['_g'], name)
        tabs = {'I%d%d' % (m, n): ('I', 4*m + n) for m in range(4) for n in range(4)}
        tabs.update({'J%d%d' % mn: ('J', j) for mn, j in _J_ORDER.items()})
        tabs.update({n: ('K', j) for j, n in enumerate(_K_ORDER)})
        if name in tabs:
            kind, row = tabs[name]
            return lambda k: self._pt(kind, row, k)
        if name in _ML_ORDER:
            return lambda k: self._ml(_ML_ORDER.index(name), k)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in ('z', 'kmin', 'kmax', 'Nk', 'sigma8_z', 'f', 'alpha_par', 'alpha_perp', 'alpha_drag', 'sigma_v', 'interpolate',
                    'k0_low', 'Pdv_model_type') and '_g' in self.__dict__:
            setattr(self._g, name, value)
        elif name.startswith('use_') and name.endswith('_model') and '_g' in self.__dict__:
            self._g.update(**{name: value})
        else:
            object.__setattr__(self, name, value)

    def update(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def initialize(self):
        self._g.initialize()

    @property
    def conformalH(self):
        """H(z) / (1 + z) (power_dm.py:547-552)"""
        return self.cosmo.H_z(self.z)/(1. + self.z)

    # -------------------------------------------------------------------- device tables -> splines in k
    def _tables(self):
        g = self._g
        eng = g._ensure_tables()
        tabs = g.__dict__.get('_dm_tabs')          # kept on the assembly model: every view of it shares one copy
        if tabs is None or tabs[0] is not g._result:
            t = {n: eng.get(g._result, n)[0] for n in ('I', 'J', 'K', 'ML', 'sigmasq_k')}
            tabs = g.__dict__['_dm_tabs'] = (g._result, eng.k_interp, t, {})
        return tabs

    def _spl(self, key, y):
        _, kint, _, cache = self._tables()
        if key not in cache:
            cache[key] = _spline(kint, y)
        return cache[key]

    def _pt(self, kind, row, k):
        """normalised I_mn / J_mn / K_mn at k: the reference's spline over the 250-point k_interp table (rsd/pt_integrals.py)"""
        _, _, t, _ = self._tables()
        norm = self._power_norm
        return (norm if kind == 'J' else norm**2)*self._spl((kind, row), t[kind][row])(np.asarray(k, dtype=float))

    def _ml(self, row, k):
        """norm^2 linear + norm^3 cross + norm^4 one-loop parts of an ImnOneLoop integral (pt_integrals.py:31-40)"""
        _, _, t, _ = self._tables()
        norm = self._power_norm
        k = np.asarray(k, dtype=float)
        return sum(norm**(2 + p)*self._spl(('ML', row, p), t['ML'][row, p])(k) for p in range(3))

    def sigmasq_k(self, k):
        _, _, t, _ = self._tables()
        return self._power_norm*self._spl(('sigk',), t['sigmasq_k'])(np.asarray(k, dtype=float))

    def normed_power_lin(self, k):
        return self._power_norm*pygcl.LinearPS(self.cosmo, 0.)(np.asarray(k, dtype=float))

    def _block(self, name, k):
        """a dark-matter block of the assembly kernel (model grid) splined to k"""
        blocks = self._g.dm_spectra()
        k = np.asarray(k, dtype=float)
        km = self.k
        if k.shape == km.shape and np.array_equal(k, km):
            return blocks[name].copy()
        return _spline(km, blocks[name])(k)

    def Pdd(self, k): return self._block('Pdd', k)
    def Pdv(self, k): return self._block('Pdv', k)
    def Pvv(self, k): return self._block('Pvv', k)

    # ------------------------------------------------------------------------------------ velocity scalars
    def _sigsq_eff(self, small):
        """sigma_lin^2 + (small-scale dispersion in km/s -> Mpc/h)^2 (dm/P02.py:27-29 and the same lines of P03 .. P22)"""
        sig = self.sigma_v if self.sigma_v is not None else self.sigma_lin
        s = small*self.cosmo.h()/(self.f*self.conformalH)
        return sig**2 + s**2

    # ---------------------------------------------------------------------------------------------- terms
    @property
    def P00(self):
        return _Term(mu0=_Angular(total=lambda k: self._block('P00', k)))

    @property
    def P01(self):
        return _Term(mu2=_Angular(total=lambda k: self._block('P01', k)))

    @property
    def P11(self):
        """dm/P11.py:3-77"""
        f = self.f

        def vector(k):
            if not self.include_2loop:
                return f**2*self.I31(k)
            return f**2*(self.Ivvdd_h01(k) + self.Idvdv_h03(k))

        def spt(k):
            k = np.asarray(k, dtype=float)
            C11 = (self.Ivvdd_h02(k) + self.Idvdv_h04(k)) if self.include_2loop else self.I13(k)
            Plin = self.normed_power_lin(k)
            part2 = 2*self.I11(k) + 4*self.I22(k) + 6*k**2*(self.J11(k) + 2*self.J10(k))*Plin
            return f**2*(Plin + part2 + C11)

        def scalar(k):
            # P11_mu4.scalar subtracts ITS vector part, which is minus the mu^2 one (dm/P11.py:33-55)
            return spt(k) + vector(k)

        def mu4_total(k):
            if self.use_P11_model:
                return self._block('P11_mu4', k)              # hzpt.P11(k) - f^2 I31(k)
            return spt(k)

        return _Term(mu2=_Angular(total=vector, vector=vector),
                     mu4=_Angular(total=mu4_total, scalar=scalar, vector=lambda k: -vector(k)))

    @property
    def P02(self):
        """dm/P02.py:3-62"""
        f = self.f

        def mu2_nv(k):
            k = np.asarray(k, dtype=float)
            return f**2*(self.I02(k) + 2.*k**2*self.J02(k)*self.normed_power_lin(k))

        def mu2_wv(k):
            k = np.asarray(k, dtype=float)
            P = self.P00.mu0(k) if self.include_2loop else self.normed_power_lin(k)
            return -(f*k)**2*self._sigsq_eff(self.sigma_bv2)*P

        def mu4_nv(k):
            k = np.asarray(k, dtype=float)
            return f**2*(self.I20(k) + 2*k**2*self.J20(k)*self.normed_power_lin(k))

        return _Term(mu2=_Angular(total=lambda k: mu2_wv(k) + mu2_nv(k), no_velocity=mu2_nv, with_velocity=mu2_wv),
                     mu4=_Angular(total=mu4_nv, no_velocity=mu4_nv))

    @property
    def P12(self):
        """dm/P12.py:3-63"""
        f = self.f

        def mu4_nv(k):
            k = np.asarray(k, dtype=float)
            return f**3*(self.I12(k) - self.I03(k) + 2*k**2*self.J02(k)*self.normed_power_lin(k))

        def mu4_wv(k):
            k = np.asarray(k, dtype=float)
            s2 = self._sigsq_eff(self.sigma_bv2)
            if self.include_2loop:
                return -0.5*(f*k)**2*s2*self.P01.mu2(k)
            return -f*(f*k)**2*s2*self.normed_power_lin(k)

        def mu6_nv(k):
            k = np.asarray(k, dtype=float)
            return f**3*(self.I21(k) - self.I30(k) + 2*k**2*self.J20(k)*self.normed_power_lin(k))

        return _Term(mu4=_Angular(total=lambda k: mu4_wv(k) + mu4_nv(k), no_velocity=mu4_nv, with_velocity=mu4_wv),
                     mu6=_Angular(total=mu6_nv, no_velocity=mu6_nv))

    @property
    def P22(self):
        """dm/P22.py:3-104"""
        f = self.f

        def mu4_nv(k):
            return 1./16*f**4*(self.Ivvvv_f23(k) if self.include_2loop else self.I23(k))

        def mu4_total(k):
            if not self.include_2loop:
                return mu4_nv(k)
            k = np.asarray(k, dtype=float)
            s2 = self._sigsq_eff(self.sigma_bv2)
            Plin, P00 = self.normed_power_lin(k), self.P00.mu0(k)
            extra = (f*k)**4*Plin*self.J02(k)**2 - 0.5*(f*k)**2*s2*self.P02.mu2.no_velocity(k) \
                + 0.25*(f*k)**4*s2**2*P00 + 0.5*(f*k)**4*P00*self.sigmasq_k(k)**2
            return mu4_nv(k) + extra

        def mu6_nv(k):
            return 1./8*f**4*(self.Ivvvv_f32(k) if self.include_2loop else self.I32(k))

        def mu6_total(k):
            if not self.include_2loop:
                return mu6_nv(k)
            k = np.asarray(k, dtype=float)
            s2 = self._sigsq_eff(self.sigma_bv2)
            extra = -0.5*(f*k)**2*s2*self.P02.mu4.no_velocity(k) + 2*(f*k)**4*self.normed_power_lin(k)*self.J02(k)*self.J20(k)
            return mu6_nv(k) + extra

        return _Term(mu4=_Angular(total=mu4_total, no_velocity=mu4_nv), mu6=_Angular(total=mu6_total, no_velocity=mu6_nv))

    @property
    def P03(self):
        """dm/P03.py:3-32"""
        f = self.f

        def mu4(k):
            k = np.asarray(k, dtype=float)
            s2 = self._sigsq_eff(self.sigma_v2)
            if self.include_2loop:
                return -0.5*(f*k)**2*s2*self.P01.mu2(k)
            return -f*(f*k)**2*s2*self.normed_power_lin(k)

        return _Term(mu4=_Angular(total=mu4, with_velocity=mu4))

    @property
    def P13(self):
        """dm/P13.py:3-54"""
        f = self.f

        def mu4(k):
            k = np.asarray(k, dtype=float)
            if not self.include_2loop:
                return 0.*k
            return -(f*k)**2*self._sigsq_eff(self.sigma_bv2)*self.P11.mu2(k)

        def mu6(k):
            k = np.asarray(k, dtype=float)
            s2 = self._sigsq_eff(self.sigma_v2)
            if self.include_2loop:
                return -(f*k)**2*s2*self.P11.mu4(k)
            # dm/P13.py:41 reads self.normed_power_lin instead of self.m.normed_power_lin and raises AttributeError in the
            # reference; the evident intent is evaluated here
            return -f**2*(f*k)**2*s2*self.normed_power_lin(k)

        return _Term(mu4=_Angular(total=mu4, with_velocity=mu4), mu6=_Angular(total=mu6, with_velocity=mu6))

    @property
    def P04(self):
        """dm/P04.py:3-59"""
        f = self.f

        def mu4_wv(k):
            k = np.asarray(k, dtype=float)
            if not self.include_2loop:
                return 0.*k
            s2 = self._sigsq_eff(self.sigma_bv4)
            return -0.5*(f*k)**2*s2*self.P02.mu2.no_velocity(k) + 0.25*(f*k)**4*s2**2*self.P00.mu0(k)

        def mu4_nv(k):
            k = np.asarray(k, dtype=float)
            if not self.include_2loop:
                return 0.*k
            return 1./12.*(f*k)**4*self.P00.mu0(k)*self.velocity_kurtosis

        def mu6(k):
            k = np.asarray(k, dtype=float)
            if not self.include_2loop:
                return 0.*k
            return -0.5*(f*k)**2*self._sigsq_eff(self.sigma_bv4)*self.P02.mu4.no_velocity(k)

        return _Term(mu4=_Angular(total=lambda k: mu4_nv(k) + mu4_wv(k), no_velocity=mu4_nv, with_velocity=mu4_wv),
                     mu6=_Angular(total=mu6, with_velocity=mu6))

    # ------------------------------------------------------------------------- sums over the terms (power_dm.py:811-846)
    def _on_grid(self, fn, k):
        """``interpolated_function(..., interp="k")``: tabulated on the model grid, splined to k"""
        km = self.k
        k = np.asarray(k, dtype=float)
        y = fn(km)
        if k.shape == km.shape and np.array_equal(k, km):
            return y
        return _spline(km, y)(k)

    def P_mu0(self, k):
        return self._on_grid(self.P00.mu0, k)

    def P_mu2(self, k):
        return self._on_grid(lambda q: self.P01.mu2(q) + self.P11.mu2(q) + self.P02.mu2(q), k)

    def P_mu4(self, k):
        def fn(q):
            Pk = self.P11.mu4(q) + self.P02.mu4(q) + self.P12.mu4(q) + self.P22.mu4(q) + self.P03.mu4(q)
            if self.include_2loop:
                Pk = Pk + self.P13.mu4(q) + self.P04.mu4(q)
            return Pk
        return self._on_grid(fn, k)

    def P_mu6(self, k):
        def fn(q):
            Pk = self.P12.mu6(q) + self.P22.mu6(q) + self.P13.mu6(q)
            if self.include_2loop:
                Pk = Pk + self.P04.mu6(q)
            return Pk
        return self._on_grid(fn, k)

    def power(self, k, mu, flatten=False):
        """sum of mu^(2i) P[mu^(2i)](k) up to ``max_mu`` with the Alcock-Paczynski remapping and volume factors of
        ``tools.alcock_paczynski`` (power_dm.py:984-1012, 1145-1163; rsd/tools.py:54-75, 195-216)"""
        if self.max_mu > 6:
            raise NotImplementedError("cannot compute power spectrum including terms with order higher than mu^6")
        k = np.atleast_1d(np.asarray(k, dtype=np.float64))
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        if k.ndim == 1 and mu.ndim == 1 and len(mu) != len(k):
            k, mu = k[:, None], mu[None, :]
        kb, mub = np.broadcast_arrays(k, mu)
        F = self.alpha_par/self.alpha_perp
        fac = np.sqrt(1. + mub**2*(1./F**2 - 1.))
        kt, mut = kb/self.alpha_perp*fac, mub/F/fac
        funcs = [self.P_mu0, self.P_mu2, self.P_mu4, self.P_mu6]
        shape = kt.shape
        P = np.zeros(kt.size)
        for i in range(self.max_mu//2 + 1):
            P += mut.ravel()**(2*i)*funcs[i](kt.ravel())
        P = np.nan_to_num(P).reshape(shape)*self.alpha_drag**3/(self.alpha_perp**2*self.alpha_par)
        P = np.squeeze(P)
        return np.ravel(P, order='F') if flatten else P


class BiasedSpectrum(object):
    """The redshift-space spectrum of two halo samples with linear biases ``b1``, ``b1_bar`` (``BiasedSpectrum``,
    ``power_biased.py:26-560``): the moment functions ``P_mu0`` .. ``P_mu6`` are the spline knots the assembly kernel's second
    stage builds for the pair (``k_gal_moments``); ``Phm`` / ``Phm_bar`` / ``stochasticity`` are its ingredients."""

    _BASE_ROWS = {'P00': 1, 'K00': 11, 'K00s': 12, 'Z00': 25}        # rows of prsd_galaxy_base (csrc/galaxy.cu)

    def __init__(self, b1=2., b1_bar=2., engine=None, **kwargs):
        self._g = _rsd.GalaxySpectrum(engine=engine, **kwargs)
        self.b1 = b1
        self.b1_bar = b1_bar

    def __getattr__(self, name):
        g = self.__dict__.get('_g')
        if g is not None and not name.startswith('__'):
            return getattr(g, name)
        raise AttributeError(name)

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if k in ('b1', 'b1_bar'):
                setattr(self, k, v)
            else:
                self._g.update(**{k: v})

    def _pair(self):
        return self.b1, self.b1_bar

    def _moments(self):
        """knots of P[mu^0], P[mu^2], P[mu^4], P[mu^6] of the pair on the model grid, and the dark-matter blocks"""
        g = self._g
        saved = g.b1_cA, g.b1_cB
        try:
            g.b1_cA, g.b1_cB = self._pair()                  # pair 1 of the assembly kernel is (cA, cB)
            km = g.k
            eng, gal, _ = g._evaluate([g._state(km)], 'power', km[len(km)//2:len(km)//2 + 1], np.zeros(1))      # one point inside the grid whatever the AP parameters
            return km, eng.galaxy_moments(gal, 1)[0, 1], eng.galaxy_base(gal, 1)[0]
        finally:
            g.b1_cA, g.b1_cB = saved

    @staticmethod
    def _to_k(km, y, k):
        k = np.asarray(k, dtype=float)
        if k.shape == km.shape and np.array_equal(k, km):
            return np.array(y, copy=True)
        return _spline(km, y)(k)

    def _eval(self, row, k):
        km, mom, _ = self._moments()
        return self._to_k(km, mom[row], k)

    def P_mu0(self, k): return self._eval(0, k)
    def P_mu2(self, k): return self._eval(1, k)
    def P_mu4(self, k): return self._eval(2, k)
    def P_mu6(self, k): return self._eval(3, k)

    def _phm(self, b1, k):
        return phm_from_base(self._g, self._moments()[2], b1, k)

    def Phm(self, k): return self._phm(self._pair()[0], k)
    def Phm_bar(self, k): return self._phm(self._pair()[1], k)

    def stochasticity(self, k):
        """power_biased.py:342-366: the fit on 20 log-spaced k between the ends of the model grid, splined"""
        g = self._g
        km = g.k
        ks = _rsd.stochasticity_grid(km)
        lo, hi = sorted(self._pair())
        return _spline(ks, g.stochasticity(ks, g.sigma8_z, lo, hi))(np.asarray(k, dtype=float))

    def power(self, k, mu, flatten=False):
        """sum of mu^(2i) P[mu^(2i)](k) up to ``max_mu`` (no Finger-of-God damping: that belongs to ``GalaxySpectrum``)"""
        k = np.atleast_1d(np.asarray(k, dtype=np.float64))
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        if k.ndim == 1 and mu.ndim == 1 and len(mu) != len(k):
            k, mu = k[:, None], mu[None, :]
        kb, mub = np.broadcast_arrays(k, mu)
        g = self._g
        F = g.alpha_par/g.alpha_perp
        fac = np.sqrt(1. + mub**2*(1./F**2 - 1.))
        kt, mut = (kb/g.alpha_perp*fac).ravel(), (mub/F/fac).ravel()
        km, mom, _ = self._moments()
        P = np.zeros(kt.size)
        for i in range(g.max_mu//2 + 1):
            P += mut**(2*i)*_spline(km, mom[i])(kt)
        P = np.squeeze(np.nan_to_num(P).reshape(kb.shape)*g.alpha_drag**3/(g.alpha_perp**2*g.alpha_par))
        return np.ravel(P, order='F') if flatten else P


def phm_from_base(g, base, b1, k):
    """the halo-matter spectrum of linear bias ``b1`` from the blocks of the assembly kernel's first stage
    (power_biased.py:310-340; HZPT form: rsd/hzpt/Phm.py:24-68, 231-296 -- Pade broadband + b1 x Zel'dovich P00)"""
    km = g.k
    R = BiasedSpectrum._BASE_ROWS
    if g.use_Phm_model:
        hm = np.asarray(_rsd.galaxy_config()[35:55]).reshape(5, 4)       # A0, R1, R1h, R2h, R x (amp, alpha, beta, run)
        x, lb = g.sigma8_z/0.8, np.log(b1)
        A0, R1, R1h, R2h, Rr = [a*b1**(al + run*lb)*x**be for a, al, be, run in hm]
        k2 = km**2
        F = 1. - 1./(1. + k2*Rr**2)
        y = F*A0*(1. + k2*R1**2)/(1. + k2*R1h**2 + (km*R2h)**4) + b1*base[R['Z00']]
    else:
        bs = -2./7*(b1 - 1.) if g.use_tidal_bias else 0.
        y = b1*base[R['P00']] + g.b2_00(b1)*base[R['K00']] + bs*base[R['K00s']]
    return BiasedSpectrum._to_k(km, y, k)


class HaloSpectrum(BiasedSpectrum):
    """halos of one linear bias: ``b1_bar`` is ``b1`` (``power_halo.py:4-23``)"""

    def __init__(self, b1=2., engine=None, **kwargs):
        BiasedSpectrum.__init__(self, b1=b1, b1_bar=b1, engine=engine, **kwargs)

    def _pair(self):
        return self.b1, self.b1


from . import model_api  # noqa: E402,F401  (attaches the rest of the reference's model surface to the three classes)
