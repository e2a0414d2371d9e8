"""``QuasarSpectrum`` (``pyRSD/rsd/power/qso/power_qso.py:48-334``): the Kaiser redshift-space spectrum of a tracer with the
scale-dependent bias of local primordial non-Gaussianity, a Finger-of-God damping and the Alcock-Paczynski distortion,

    P(k, mu) = G(k mu sigma_fog)^2 [b_tot(k)^2 + 2 f b_tot(k) mu^2 + f^2 mu^4] P_L(k) / (alpha_perp^2 alpha_par) alpha_drag^3 + N

at the distorted (k', mu').  As in the reference every k-dependent piece is a cubic spline through the 500 points of
``k_interp = logspace(1e-8, 100, 500)`` (``interpolated_function`` of rsd/_cache.py:344-389 over ``RSDSpline`` = FITPACK's
interpolating spline) and the only call into the integration library is the linear spectrum on those 500 points -- here
``pygcl.LinearPS`` on the device."""
import numpy as np
from scipy.integrate import quad
from scipy.interpolate import InterpolatedUnivariateSpline as _spline

from . import pygcl
from . import rsd as _rsd
from . import model_api as _api

C_LIGHT_KMS = 2.99792458e5            # pygcl.Constants.c_light / pygcl.Constants.km (Common.i:36, 52)


def growth_function(cosmo, z):
    """unnormalised growth factor E(a) int_0^a da' / (a' E(a'))^3 (power_qso.py:32-46)"""
    Om0, Ode0, Ok0 = cosmo['Omega0_m'], cosmo['Omega0_lambda'], cosmo['Omega0_k']

    def efunc(a):
        return np.sqrt(a**-3*Om0 + Ok0*a**-2 + Ode0)
    return efunc(1./(1 + z))*quad(lambda a: 1.0/(a*efunc(a))**3, 0., 1./(1 + z))[0]


class QuasarSpectrum(object):
    """The quasar redshift-space power spectrum.  ``params`` must be a :class:`pyrsd_b200.pygcl.Cosmology`."""

    k_interp = np.logspace(np.log10(1e-8), np.log10(100.0), 500)
    delta_crit = 1.686

    def __init__(self, kmin=1e-3, kmax=0.5, Nk=200, z=0., params=None, fog_model='gaussian', max_mu=4, **kwargs):
        if not isinstance(params, pygcl.Cosmology):
            raise TypeError("`params` must be a pyrsd_b200.pygcl.Cosmology (the linear T(k) is a host input)")
        self.cosmo = params
        self.kmin, self.kmax, self.Nk, self.z, self.max_mu = kmin, kmax, Nk, z, max_mu
        self.fog_model = fog_model
        self.sigma8_z = params.Sigma8_z(z)
        self.f = params.f_z(z)
        self.alpha_par = self.alpha_perp = self.alpha_drag = 1.
        self.b1 = kwargs.pop('b1', 2.)
        self.sigma_fog, self.N, self.f_nl, self.p = 4.0, 0., 0., 1.6
        self.include_2loop = False
        if kwargs:
            raise TypeError("unexpected keyword arguments %s" % sorted(kwargs))
        self._plin = None

    def __setattr__(self, name, value):
        if name == 'fog_model' and value not in ('modified_lorentzian', 'lorentzian', 'gaussian'):
            raise ValueError("`fog_model` must be one of %s" % ['modified_lorentzian', 'lorentzian', 'gaussian'])
        if name == 'p' and not 1 <= value <= 1.6:
            raise ValueError("f_nl ``p`` value should be between 1 and 1.6 (inclusive)")
        object.__setattr__(self, name, value)

    def update(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getstate__(self):
        d = dict(self.__dict__)
        d['_plin'] = None
        return d

    # ---- pieces
    @property
    def k(self):
        return np.logspace(np.log10(self.kmin), np.log10(self.kmax), self.Nk)

    @property
    def D(self):
        return self.cosmo.D_z(self.z)

    @property
    def _power_norm(self):
        return (self.sigma8_z/self.cosmo.sigma8())**2

    FOG = _api.G.FOG

    def _plin_interp(self):
        """P_L(z = 0) on k_interp: the one device evaluation behind every spline of the model"""
        if self._plin is None:
            self._plin = np.asarray(pygcl.LinearPS(self.cosmo, 0.)(self.k_interp))
        return self._plin

    def normed_power_lin(self, k):
        return self._power_norm*pygcl.LinearPS(self.cosmo, 0.)(np.asarray(k, dtype=float))

    def _knots(self):
        """knot values of alpha_png, delta_bias, b_tot, P[mu^0], P[mu^2], P[mu^4] on k_interp (power_qso.py:186-260)"""
        k = self.k_interp
        Pn = self._power_norm*self._plin_interp()
        Tk = (Pn/k**self.cosmo.n_s())**0.5
        Tk = Tk/Tk.max()
        g_ratio = growth_function(self.cosmo, 0.)/((1 + 100.)*growth_function(self.cosmo, 100.))
        alpha = 2*k**2*Tk*self.D/(3*self.cosmo.Omega0_m())*(C_LIGHT_KMS/100.)**2*g_ratio
        dbias = 0.*k if self.f_nl == 0 else 2*(self.b1 - self.p)*self.f_nl*self.delta_crit/alpha
        btot = self.b1 + dbias
        return dict(alpha_png=alpha, delta_bias=dbias, btot=btot, P_mu0=btot**2*Pn, P_mu2=2*self.f*btot*Pn, P_mu4=self.f**2*Pn)

    def _fn(self, name, k):
        k = np.asarray(k, dtype=float)
        if k.size and (k.min() < self.k_interp[0] or k.max() > self.k_interp[-1]):
            raise ValueError("wavenumber range outside interpolation domain [%.1e, %.1e]" % (self.k_interp[0], self.k_interp[-1]))
        return _spline(self.k_interp, self._knots()[name])(k)

    def alpha_png(self, k): return self._fn('alpha_png', k)
    def delta_bias(self, k): return self._fn('delta_bias', k)
    def btot(self, k): return self._fn('btot', k)
    def P_mu0(self, k): return self._fn('P_mu0', k)
    def P_mu2(self, k): return self._fn('P_mu2', k)
    def P_mu4(self, k): return self._fn('P_mu4', k)
    def P_mu6(self, k): return 0.*np.asarray(k, dtype=float)

    def power(self, k, mu, flatten=False):
        if self.max_mu > 6:
            raise NotImplementedError("cannot compute power spectrum including terms with order higher than mu^6")
        k = np.atleast_1d(np.asarray(k, dtype=np.float64))
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        if k.ndim == 1 and mu.ndim == 1 and len(mu) != len(k):
            k, mu = k[:, None], mu[None, :]
        kb, mub = np.broadcast_arrays(k, mu)
        F = self.alpha_par/self.alpha_perp
        fac = np.sqrt(1. + mub**2*(1./F**2 - 1.))
        kt, mut = (kb/self.alpha_perp*fac).ravel(), (mub/F/fac).ravel()
        knots = self._knots()
        P = np.zeros(kt.size)
        for i, name in enumerate(('P_mu0', 'P_mu2', 'P_mu4')[:self.max_mu//2 + 1]):
            P += mut**(2*i)*_spline(self.k_interp, knots[name])(kt)
        P = np.nan_to_num(P)*self.FOG(kt, mut, self.sigma_fog)**2
        # power_qso.py:262-292: the constant N is added inside the Alcock-Paczynski wrapper, i.e. it carries the volume factor
        P = (P + self.N)*self.alpha_drag**3/(self.alpha_perp**2*self.alpha_par)
        P = np.squeeze(P.reshape(kb.shape))
        return np.ravel(P, order='F') if flatten else P

    poles = _api._generic_poles
    monopole, quadrupole = _api._generic_pole_method(0), _api._generic_pole_method(2)
    hexadecapole, tetrahexadecapole = _api._generic_pole_method(4), _api._generic_pole_method(6)
    apply_transfer = _rsd._apply_transfer
    derivative_k = lambda self, k, mu: _rsd._derivative(self, k, mu, 'k')
    derivative_mu = lambda self, k, mu: _rsd._derivative(self, k, mu, 'mu')
