"""Drop-in for ``pyRSD.pygcl`` on the perturbation-theory hot path.

Same class names, constructor / call signatures and return types as the reference's SWIG
classes (``pyRSD/_gcl/python/*.i`` re-exported by ``pyRSD/pygcl.py``); every numerical result is
computed by ``libpyrsd_b200.so`` on the GPU (hand-written sm_100a CUDA through the C-ABI of
``include/pyrsd_b200.h``).  There is no CPU fallback.

Deliberate differences:

* ``epsrel`` constructor arguments are accepted and stored but do not steer anything: the adaptive
  Genz-Malik / Gauss-Kronrod quadrature is replaced by a fixed, converged rule (accuracy ~1e-6,
  i.e. what the reference delivers only at ``epsrel ~ 1e-8``).
* ``Cosmology`` does not run CLASS: the linear transfer function is a host input
  (``Cosmology(params, tf, sigma8, k, Tk)`` -- the reference's own pickle-state constructor,
  ``pyRSD/pygcl.py:60-63``) and the growth history (``D_z``, ``f_z``, ``H_z``) is supplied through
  ``params`` (keys ``'growth'``, ``'f'``, ``'H'``: dicts or callables of z).
"""
import ctypes as C

import numpy as np

from . import _native as N

__all__ = ['Cosmology', 'ClassParams', 'transfers', 'IntegrationMethods', 'LinearPS', 'Imn', 'Jmn', 'Kmn',
           'ImnOneLoop', 'OneLoopPdd', 'OneLoopPdv', 'OneLoopPvv', 'OneLoopP22Bar', 'ZeldovichPS', 'ZeldovichP00',
           'ZeldovichP01', 'ZeldovichP11', 'ZeldovichCF', 'pk_to_xi', 'xi_to_pk', 'ComputeXiLM', 'ComputeXiLM_fftlog',
           'CubicSpline', 'LinearSpline', 'FFTLog', 'CorrelationFunction', 'Kaiser', 'SimpsIntegrate', 'TrapzIntegrate']

_vp = C.c_void_p
_dp = N._dp


def _L():
    L = N.lib()
    if not getattr(L, '_pygcl_sigs', False):
        i, d = C.c_int, C.c_double
        L.prsd_fftlog.argtypes = [_vp, i, i, d, d, d, d, i, _dp, _dp, C.POINTER(d)]
        L.prsd_fftlog.restype = i
        L.prsd_zeldovich.argtypes = [_vp, i, _dp, _vp, _dp, _dp, _vp, i, d, _vp, _vp, i, _dp, _dp]
        L.prsd_zeldovich.restype = i
        L.prsd_compute_xilm.argtypes = [_vp, i, i, i, i, _dp, _dp, i, _dp, d, d, _dp]
        L.prsd_compute_xilm.restype = i
        L.prsd_compute_xilm_discrete.argtypes = [_vp, i, i, i, i, i, _dp, _dp, i, _dp, d, d, _dp]
        L.prsd_compute_xilm_discrete.restype = i
        L.prsd_discrete_integrate.argtypes = [_vp, i, i, i, _dp, _dp, _dp]
        L.prsd_discrete_integrate.restype = i
        L.prsd_compute_xilm_fftlog.argtypes = [_vp, i, i, i, i, _dp, _dp, d, d, _dp, _dp]
        L.prsd_compute_xilm_fftlog.restype = i
        L.prsd_oneloop_tables.argtypes = [_vp, _vp, _vp]
        L.prsd_oneloop_tables.restype = i
        L.prsd_spectra_set_oneloop_one.argtypes = [_vp, i, _dp]
        L.prsd_spectra_set_oneloop_one.restype = i
        L.prsd_spectra_get_oneloop.argtypes = [_vp, _dp]
        L.prsd_spectra_get_oneloop.restype = i
        L.prsd_cubic_spline_eval.argtypes = [_vp, i, i, _dp, _dp, i, _dp, _dp]
        L.prsd_cubic_spline_eval.restype = i
        L._pygcl_sigs = True
    return L


def _scalar_or_array(x):
    """SWIG dispatches on scalar vs sequence (Common.i:92-110): return (array, was_scalar)"""
    a = np.asarray(x, dtype=np.float64)
    return np.ascontiguousarray(np.atleast_1d(a)), a.ndim == 0


def _ret(arr, scalar):
    return float(arr[0]) if scalar else arr


class _Enum(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


# Cosmology.h:9-16 and CorrelationFunction.h:8
transfers = _Enum(CLASS=0, EH=1, EH_NoWiggle=2, BBKS=3, FromArrays=4)
IntegrationMethods = _Enum(FFTLOG=0, SIMPS=1, TRAPZ=2)


class ClassParams(dict):
    """dict of CLASS-style parameters (ClassParams.i); ``from_dict`` as used by pygcl.py:66"""
    @classmethod
    def from_dict(cls, d):
        return cls(d)


class PickalableSWIG(object):
    """same pickling contract as pyRSD/pygcl.py:28-34: the state is the constructor arguments"""
    def __setstate__(self, state):
        self.__init__(*state['args'])

    def __getstate__(self):
        return {'args': self.args}


def _lookup(table, z, what):
    if table is None:
        raise ValueError("Cosmology: `%s` was not supplied (CLASS is not run by this library)" % what)
    if callable(table):
        return float(table(z))
    try:
        return float(table[float(z)])
    except KeyError:
        raise ValueError("Cosmology: no `%s` value for z = %s" % (what, z))


def eh_nowiggle_transfer(k, h, Omega0_m, Omega0_b, Tcmb=2.7255):
    """Eisenstein & Hu (1998) no-wiggle transfer function, eqs. 26, 28-31, with the parameters of
    ``Cosmology::SetEisensteinHuParameters`` (Cosmology.cpp:233-246, 311-362); k in h/Mpc"""
    k = np.asarray(k, dtype=float)
    omh2, obh2 = Omega0_m*h*h, Omega0_b*h*h
    fb = obh2/omh2
    theta = Tcmb/2.7
    keq = 0.0746*omh2/theta**2
    s = h*44.5*np.log(9.83/omh2)/np.sqrt(1. + 10.*obh2**0.75)
    alpha_gamma = 1. - 0.328*np.log(431.*omh2)*fb + 0.38*np.log(22.3*omh2)*fb*fb
    kk = k*h
    ks = kk*s/h
    q = kk/(13.41*keq)
    geff = omh2*(alpha_gamma + (1. - alpha_gamma)/(1. + (0.43*ks)**4))
    qeff = q*omh2/geff
    L0 = np.log(2.*np.e + 1.8*qeff)
    C0 = 14.2 + 731./(1. + 62.5*qeff)
    return L0/(L0 + C0*qeff**2)


class Cosmology(PickalableSWIG):
    """Host container of a linear transfer function and its normalisation.

    ``Cosmology(params, tf, sigma8, k, Tk)`` -- the state constructor of the reference
    (Cosmology.i:16-74, pygcl.py:48-95).  ``params`` must hold ``h``, ``Omega_b``, ``Omega_cdm`` (and
    optionally ``m_ncdm``, ``T_cmb``, ``n_s``); the spline + Eisenstein-Hu extrapolation of T(k) and
    the sigma8 normalisation follow Cosmology.cpp:153-246, 283-369 (evaluated on the GPU)."""

    def __init__(self, params, tf=0, sigma8=None, k=None, Tk=None, context=None):
        if k is None or Tk is None or sigma8 is None:
            raise ValueError("pyrsd_b200.Cosmology needs the discrete transfer function: "
                             "Cosmology(params, tf, sigma8, k, Tk); CLASS is not run by this library")
        if tf not in (transfers.CLASS, transfers.FromArrays, transfers.EH_NoWiggle):
            raise ValueError("only tabulated transfer functions (transfers.CLASS) and their no-wiggle clones are supported")
        self._params = ClassParams(params)
        self._tf = int(tf)
        self._sigma8 = float(sigma8)
        self._k = np.ascontiguousarray(k, dtype=np.float64)
        self._Tk = np.ascontiguousarray(Tk, dtype=np.float64)
        self._ctx = context if context is not None else N.default_context()
        self._delta_H = None

    # ---- pickling: (params dict, tf, sigma8, k_i, T_i) as pygcl.py:60-63
    @property
    def args(self):
        return [dict(self._params), self._tf, self._sigma8, self._k, self._Tk]

    def __setstate__(self, state):
        a = state['args']
        self.__init__(ClassParams.from_dict(a[0]), *a[1:])

    def clone(self, tf=None):
        """a copy; ``tf=transfers.EH_NoWiggle``: the same parameters with the Eisenstein-Hu no-wiggle transfer function
        tabulated on the same k (Cosmology.cpp:233-246 on the grid of :270, as ``SetTransferFunction`` does)"""
        if tf is None or tf == self._tf:
            return Cosmology(*self.args)
        if tf != transfers.EH_NoWiggle:
            raise ValueError("only tabulated transfer functions and their Eisenstein-Hu no-wiggle clones are supported")
        a = self.args
        return Cosmology(a[0], transfers.EH_NoWiggle, a[2], a[3], self.GetNoWiggleTransfer(a[3]))

    def GetNoWiggleTransfer(self, k):
        """Eisenstein & Hu (1998) no-wiggle fit at this cosmology's parameters (Cosmology.cpp:233-246)"""
        return eh_nowiggle_transfer(k, self.h(), self.Omega0_m(), self.Omega0_b(), self.Tcmb())

    # ---- accessors
    def GetParams(self): return self._params
    def GetTransferFit(self): return self._tf
    def GetDiscreteK(self): return self._k.copy()
    def GetDiscreteTk(self): return self._Tk.copy()
    def sigma8(self): return self._sigma8
    def h(self): return float(self._params['h'])
    def n_s(self): return float(self._params.get('n_s', 0.96))
    def Tcmb(self): return float(self._params.get('T_cmb', 2.7255))
    def Omega0_b(self): return float(self._params['Omega_b'])

    def Omega0_m(self):
        """CLASS's background Omega_m: baryons + CDM + massive neutrinos (ClassEngine.cpp:61-62)"""
        p = self._params
        if 'Omega_m' in p:
            return float(p['Omega_m'])
        return float(p['Omega_b']) + float(p['Omega_cdm']) + float(p.get('m_ncdm', 0.0))/93.14/self.h()**2

    def Omega0_k(self):
        return float(self._params.get('Omega_k', 0.0))

    def Omega0_lambda(self):
        """dark-energy density today; flat LCDM (1 - Omega_m - Omega_k) unless ``Omega_Lambda`` is among the parameters"""
        return float(self._params.get('Omega_Lambda', 1.0 - self.Omega0_m() - self.Omega0_k()))

    def D_z(self, z): return 1.0 if (z == 0 and self._params.get('growth') is None) else _lookup(self._params.get('growth'), z, 'growth')
    def f_z(self, z): return _lookup(self._params.get('f'), z, 'f')
    def H_z(self, z): return 100.*self.h() if (z == 0 and self._params.get('H') is None) else _lookup(self._params.get('H'), z, 'H')
    def Sigma8_z(self, z): return self._sigma8*self.D_z(z)
    def H0(self): return 100.*self.h()

    def Omega_m_z(self, z):
        """ClassEngine.cpp:572-577"""
        Ez = self.H_z(z)/self.H0()
        return self.Omega0_m()*(1. + z)**3/Ez**2

    def rho_crit_z(self, z):
        """critical density in h^2 M_sun / Mpc^3 with the reference's constants (ClassEngine.cpp:552-570, Common.h:101-140)"""
        G, M_sun, au = 6.67384e-8, 1.9891e33, 149597870700.*1e2
        Mpc = 1e6*au/(np.pi/180./60./60.)
        H = 100.*1e5/Mpc
        rho = 3.*H**2/(8.*np.pi*G)/(M_sun/Mpc**3)
        return rho*(self.H_z(z)/self.H0())**2

    def rho_bar_z(self, z):
        """mean matter density in h^2 M_sun / Mpc^3 (ClassEngine.cpp:579-582)"""
        return self.Omega_m_z(z)*self.rho_crit_z(z)
    def rs_drag(self):
        if self._params.get('rs_drag') is None:
            raise ValueError("Cosmology: `rs_drag` was not supplied (CLASS is not run by this library)")
        return float(self._params['rs_drag'])

    def __getitem__(self, key):
        if hasattr(self, key):
            f = getattr(self, key)
            if callable(f):
                return f()
        raise KeyError("Sorry, cannot return parameter '%s' in dict-like fashion" % key)

    def _pars(self, amp):
        return np.array([[self.n_s(), amp, self.h(), self.Omega0_m(), self.Omega0_b(), self.Tcmb(), 0., 0.]])

    def _make_spectra(self, amp):
        return _Spectra(self._ctx, self._k[None, :], self._Tk[None, :], self._pars(amp))

    def delta_H(self):
        """normalisation of P_L at z = 0 such that sigma(R = 8) = sigma8 (Cosmology.cpp:167-184)"""
        if self._delta_H is None:
            sp = self._make_spectra(1.0)
            sig = sp.ps_integral(5, [8.0])[0, 0]
            self._delta_H = self._sigma8/sig
        return self._delta_H

    def SetSigma8(self, sigma8):
        self._sigma8 = float(sigma8)
        self._delta_H = None

    def NormalizeTransferFunction(self, sigma8):
        """delta_H such that sigma(R = 8) = sigma8 (Cosmology.cpp:175-184); here the stored sigma8 follows"""
        self.SetSigma8(sigma8)

    @classmethod
    def from_power(cls, params, k, Pk, sigma8=None, context=None):
        """a cosmology whose transfer function is read off a linear power spectrum: T = sqrt(P / k^n_s), largest value
        1e7 (``Cosmology::FromPower``, Cosmology.cpp:82-118).  CLASS supplies sigma8 in the reference; here it is the
        argument or ``params['sigma8']``."""
        p = ClassParams(params)
        if sigma8 is None:
            if 'sigma8' not in p:
                raise ValueError("Cosmology.from_power: `sigma8` is needed (CLASS is not run by this library)")
            sigma8 = p['sigma8']
        k, Pk = np.asarray(k, dtype=float), np.asarray(Pk, dtype=float)
        T = np.sqrt(Pk/k**float(p.get('n_s', 0.96)))
        return cls(p, transfers.CLASS, sigma8, k, T/T.max()*1e7, context=context)

    @classmethod
    def from_file(cls, params, k, Tk, sigma8=None, context=None):
        """``Cosmology(param_file, k, Tk)`` of pygcl.py:86-88 with the parameters as a mapping"""
        p = ClassParams(params)
        if sigma8 is None:
            if 'sigma8' not in p:
                raise ValueError("Cosmology.from_file: `sigma8` is needed (CLASS is not run by this library)")
            sigma8 = p['sigma8']
        return cls(p, transfers.CLASS, sigma8, k, Tk, context=context)

    def EvaluateTransfer(self, k):
        ka, scalar = _scalar_or_array(k)
        sp = self._make_spectra(1.0)
        T = np.sqrt(sp.plin(ka)[0]/ka**self.n_s())
        return _ret(T, scalar)


class _Spectra(N.Handle):
    """owner of a ``prsd_spectra`` handle (batch of cosmologies)"""
    _destroy = 'prsd_spectra_destroy'

    def __init__(self, ctx, k, T, pars):
        k = np.ascontiguousarray(k, dtype=np.float64)
        T = np.ascontiguousarray(T, dtype=np.float64)
        pars = np.ascontiguousarray(pars, dtype=np.float64)
        self.ncosmo, self.nspl = k.shape
        assert T.shape == k.shape and pars.shape == (self.ncosmo, 8)
        p = _vp()
        N.check(_L().prsd_spectra_create(ctx.ptr, self.ncosmo, self.nspl, k, T, pars, C.byref(p)))
        N.Handle.__init__(self, p)
        self.ctx = ctx
        self.have_oneloop = False

    def plin(self, q):
        q = N.as_f64(q)
        out = np.empty((self.ncosmo, len(q)))
        N.check(_L().prsd_spectra_plin(self.ptr, len(q), q, out))
        return out

    def rescale(self, fac):
        N.check(_L().prsd_spectra_rescale(self.ptr, N.as_f64(fac)))
        self._computed_oneloop = None        # a one-loop object built from now on sees the new normalisation

    def ps_integral(self, what, x, factor=0.5):
        x = N.as_f64(x)
        out = np.empty((self.ncosmo, max(len(x), 1)))
        N.check(_L().prsd_eval_ps_integral(self.ptr, what, len(x), x, factor, out))
        return out

    def correlation(self, r, kmin, kmax):
        r = N.as_f64(r)
        out = np.empty((self.ncosmo, max(len(r), 1)))
        N.check(_L().prsd_eval_correlation(self.ptr, len(r), r, float(kmin), float(kmax), out))
        return out

    def ensure_oneloop(self):
        """the four one-loop tables COMPUTED from this batch's linear spectra (cached: objects built on explicit
        ``oneloop_power`` arrays overwrite the device copy, never this cache)"""
        if getattr(self, '_computed_oneloop', None) is None:
            out = np.empty((self.ncosmo, 4, 1000))
            N.check(_L().prsd_oneloop_tables(self.ptr, None, out.ctypes.data_as(_vp)))
            self._computed_oneloop = out
            self.have_oneloop = True
        return self._computed_oneloop

    def oneloop_tables(self):
        out = np.empty((self.ncosmo, 4, 1000))
        N.check(_L().prsd_spectra_get_oneloop(self.ptr, out))
        return out


class PowerSpectrum(object):
    """methods of the reference's PowerSpectrum base class (PowerSpectrum.i:5-45)"""
    _sp = None

    def _eval(self, k):
        raise NotImplementedError

    def __call__(self, k):
        ka, scalar = _scalar_or_array(k)
        return _ret(self._eval(ka), scalar)

    Evaluate = __call__
    EvaluateMany = __call__

    def _integral(self, what, x, factor=0.5):
        xa, scalar = _scalar_or_array(x)
        return _ret(self._sp.ps_integral(what, xa, factor)[0], scalar)

    def Sigma(self, R): return self._integral(5, R)
    def X_Zel(self, r): return self._integral(2, r)
    def Y_Zel(self, r): return self._integral(3, r)
    def Q3_Zel(self, k): return self._integral(4, k)
    def sigma3_squared(self, k): return self._integral(6, k)

    def VelocityDispersion(self, k=None, factor=0.5):
        if k is None:
            return float(self._sp.ps_integral(0, np.zeros(1))[0, 0])
        return self._integral(1, k, factor)

    def NonlinearScale(self):
        return 1.0/np.sqrt(self.VelocityDispersion())

    def Save(self, filename, kmin=1e-3, kmax=1., Nk=1001, log=False):
        """two-column text file of (k, P(k)) (PowerSpectrum.i:17, PowerSpectrum.cpp:77-92)"""
        import time
        i = np.arange(int(Nk))
        k = kmin*np.exp(i*np.log(kmax/kmin)/(Nk - 1)) if log else kmin + i*(kmax - kmin)/(Nk - 1)
        P = np.atleast_1d(self(k))
        with open(filename, 'w+') as fp:
            fp.write("# Power spectrum saved at %s\n\n" % time.ctime())
            fp.write("# k -- P(k)\n")
            for a, b in zip(k, P):
                fp.write("%e %e\n" % (a, b))


class LinearPS(PowerSpectrum, PickalableSWIG):
    """``LinearPS(cosmo, z=0)`` (LinearPS.i:5-14): P_L(k) = delta_H^2 (sigma8_z/sigma8)^2 k^ns T(k)^2"""

    def __init__(self, cosmo, z=0.):
        self.args = (cosmo, z)
        self._cosmo, self._z = cosmo, float(z)
        self._sigma8_z = cosmo.Sigma8_z(z)
        amp = cosmo.delta_H()**2*(self._sigma8_z/cosmo.sigma8())**2
        self._sp = cosmo._make_spectra(amp)

    def __setstate__(self, state):
        self.__init__(*state['args'][:2])
        self.SetSigma8AtZ(state['args'][2])

    def __getstate__(self):
        return {'args': self.args + (self.GetSigma8AtZ(),)}

    def _eval(self, k):
        return self._sp.plin(k)[0]

    def GetRedshift(self): return self._z
    def GetSigma8AtZ(self): return self._sigma8_z
    def GetCosmology(self): return self._cosmo

    def SetSigma8AtZ(self, sigma8_z):
        if sigma8_z <= 0:
            raise ValueError("sigma8_z must be positive")
        self._sp.rescale([(sigma8_z/self._sigma8_z)**2])
        self._sigma8_z = float(sigma8_z)
        self._sp.have_oneloop = False


def _kernel_call(fn, P_L, k, *ints):
    ka, scalar = _scalar_or_array(k)
    out = np.empty((1, len(ka)))
    N.check(fn(P_L._sp.ptr, *ints, len(ka), ka, None, out))
    return _ret(out[0], scalar)


class Imn(PickalableSWIG):
    """``Imn(P_L, epsrel=1e-4)(k, m, n)`` (Imn.i:8-14, Imn.cpp:142-189)"""
    def __init__(self, P_L, epsrel=1e-4):
        self.args = (P_L, epsrel)
        self._P, self._eps = P_L, epsrel

    def GetLinearPS(self): return self._P
    def GetEpsrel(self): return self._eps

    def __call__(self, k, m, n):
        return _kernel_call(_L().prsd_eval_imn, self._P, k, int(m), int(n))

    Evaluate = EvaluateMany = __call__


class Jmn(PickalableSWIG):
    """``Jmn(P_L, epsrel=1e-4)(k, m, n)`` (Jmn.i:8-14, Jmn.cpp:96-139)"""
    def __init__(self, P_L, epsrel=1e-4):
        self.args = (P_L, epsrel)
        self._P, self._eps = P_L, epsrel

    def GetLinearPS(self): return self._P
    def GetEpsrel(self): return self._eps

    def __call__(self, k, m, n):
        ka, scalar = _scalar_or_array(k)
        out = np.empty((1, len(ka)))
        N.check(_L().prsd_eval_jmn(self._P._sp.ptr, int(m), int(n), len(ka), ka, out))
        return _ret(out[0], scalar)

    Evaluate = EvaluateMany = __call__


class Kmn(PickalableSWIG):
    """``Kmn(P_L, epsrel=1e-3)(k, m, n, tidal=False, part=0)`` (Kmn.i:8-14, Kmn.cpp:109-184)"""
    def __init__(self, P_L, epsrel=1e-3):
        self.args = (P_L, epsrel)
        self._P, self._eps = P_L, epsrel

    def GetLinearPS(self): return self._P
    def GetEpsrel(self): return self._eps

    def __call__(self, k, m, n, tidal=False, part=0):
        return _kernel_call(_L().prsd_eval_kmn, self._P, k, int(m), int(n), int(bool(tidal)), int(part))

    Evaluate = EvaluateMany = __call__


class OneLoopPS(PowerSpectrum, PickalableSWIG):
    """base of the one-loop spectra (OneLoopPS.i:5-63): 1000-point table on logspace(1e-5, 10) + spline"""
    _which = None

    def __init__(self, P_L, epsrel=1e-4, oneloop_power=None):
        self.args = (P_L,)
        self._P, self._eps = P_L, epsrel
        self._sp = P_L._sp
        if oneloop_power is None:
            self._table = self._sp.ensure_oneloop()[0, self._which].copy()
        else:
            t = np.ascontiguousarray(np.asarray(oneloop_power, dtype=np.float64).reshape(1, -1))
            if t.shape[1] != 1000:
                raise ValueError("oneloop_power must have 1000 entries (OneLoopPS.cpp:11-13)")
            self._table = t[0].copy()
        self._install()

    def __getstate__(self):
        return {'args': self.args + (self.GetEpsrel(), self.GetOneLoopPower())}

    def _install(self):
        """make sure the shared device tables hold THIS object's table"""
        cur = self._sp.oneloop_tables()[0, self._which] if self._sp.have_oneloop else None
        if cur is None or not np.array_equal(cur, self._table):
            N.check(_L().prsd_spectra_set_oneloop_one(self._sp.ptr, self._which, self._table.reshape(1, -1)))
            self._sp.have_oneloop = True

    def _ev(self, k, full):
        self._install()
        out = np.empty((1, len(k)))
        N.check(_L().prsd_spectra_oneloop_eval(self._sp.ptr, self._which, int(full), len(k), k, out))
        return out[0]

    def _eval(self, k):
        return self._ev(k, False)

    def EvaluateFull(self, k):
        ka, scalar = _scalar_or_array(k)
        return _ret(self._ev(ka, True), scalar)

    def GetLinearPS(self): return self._P
    def GetRedshift(self): return self._P.GetRedshift()
    def GetCosmology(self): return self._P.GetCosmology()
    def GetEpsrel(self): return self._eps
    def GetOneLoopPower(self): return self._table.copy()


class OneLoopPdd(OneLoopPS):
    """2 I00 + 6 k^2 J00 P_L (OneLoopPS.cpp:56-71)"""
    _which = 0


class OneLoopPdv(OneLoopPS):
    """2 I01 + 6 k^2 J01 P_L (OneLoopPS.cpp:85-101)"""
    _which = 1


class OneLoopPvv(OneLoopPS):
    """2 I11 + 6 k^2 J11 P_L (OneLoopPS.cpp:116-132)"""
    _which = 2


class OneLoopP22Bar(OneLoopPS):
    """I23 + 2/3 I32 + 1/5 I33 (OneLoopPS.cpp:164-185)"""
    _which = 3

    def VelocityKurtosis(self):
        self._install()
        out = np.empty(1)
        N.check(_L().prsd_eval_kurtosis(self._sp.ptr, out))
        return float(out[0])


class ImnOneLoop(PickalableSWIG):
    """``ImnOneLoop(P1[, P2], epsrel=1e-4)`` (ImnOneLoop.i:8-23, ImnOneLoop.cpp:96-216)"""
    def __init__(self, P_1, *args):
        self.args = (P_1,) + tuple(args)
        self._P1, self._P2, self._eps = P_1, None, 1e-4
        for a in args:
            if isinstance(a, OneLoopPS):
                self._P2 = a
            else:
                self._eps = a
        for P in (P_1, self._P2):
            if P is not None and (not isinstance(P, OneLoopPS) or P._which > 2):
                raise TypeError("ImnOneLoop needs OneLoopPdd / OneLoopPdv / OneLoopPvv spectra")
        if self._P2 is not None and self._P2._sp is not self._P1._sp:
            raise ValueError("both one-loop spectra must be built on the same LinearPS")

    def _ev(self, which, k, m, n):
        self._P1._install()
        p2 = -1
        if self._P2 is not None:
            self._P2._install()
            p2 = self._P2._which
        ka, scalar = _scalar_or_array(k)
        out = np.empty((1, len(ka)))
        N.check(_L().prsd_eval_imn_oneloop(self._P1._sp.ptr, self._P1._which, p2, which, int(m), int(n), len(ka), ka, None, out))
        return _ret(out[0], scalar)

    def EvaluateLinear(self, k, m, n): return self._ev(0, k, m, n)
    def EvaluateCross(self, k, m, n): return self._ev(1, k, m, n)
    def EvaluateOneLoop(self, k, m, n): return self._ev(2, k, m, n)
    def GetEpsrel(self): return self._eps
    def GetOneLoopPS1(self): return self._P1
    def GetOneLoopPS2(self): return self._P2 if self._P2 is not None else self._P1
    def GetEqual(self): return self._P2 is None          # single-spectrum constructor (ImnOneLoop.cpp:14-22)


# ------------------------------------------------------------------------------ Zel'dovich
_ZEL_N = 1024


def _zel_rgrid():
    # ZeldovichPS::InitializeR (ZeldovichPS.cpp:41-53)
    i = np.arange(1, _ZEL_N + 1)
    return 10.**(1.5 + (i - 0.5*(_ZEL_N + 1))*(7.0/_ZEL_N))


class ZeldovichPS(PickalableSWIG):
    """``ZeldovichPS(C, z, approx_lowk=False)`` or the 8-argument state constructor
    ``ZeldovichPS(C, approx_lowk, sigma8_z, k0_low, sigma_sq, X0, XX, YY)`` (ZeldovichPS.i:5-71)"""
    _index = None

    def __init__(self, *args):
        if len(args) == 1 and isinstance(args[0], ZeldovichPS):
            src = args[0]
            self._init_state(src._cosmo, src._lowk, src._s8z, src._k0, src._ssq, src._X0.copy(), src._XX.copy(), src._YY.copy())
            self.args = (src._cosmo,)
        elif len(args) == 8:
            self._init_state(args[0], bool(args[1]), float(args[2]), float(args[3]), float(args[4]),
                             np.array(args[5], dtype=np.float64), np.array(args[6], dtype=np.float64),
                             np.array(args[7], dtype=np.float64))
            self.args = (args[0],)
        elif len(args) in (2, 3):
            cosmo, z = args[0], args[1]
            lowk = bool(args[2]) if len(args) == 3 else False
            # integrals at z = 0, then scaled by (sigma8_z/sigma8)^2 (ZeldovichPS.cpp:15-30)
            P0 = LinearPS(cosmo, 0.)
            s8z = cosmo.Sigma8_z(z)
            norm = (s8z/cosmo.sigma8())**2
            r = _zel_rgrid()
            ssq = norm*P0.VelocityDispersion()
            X0 = norm*P0.X_Zel(r)
            YY = norm*P0.Y_Zel(r)
            self._init_state(cosmo, lowk, s8z, 5e-3, ssq, X0, X0 + 2.*ssq, YY)
            self.args = (cosmo,)
        else:
            raise TypeError("ZeldovichPS(C, z[, approx_lowk]) | ZeldovichPS(C, approx_lowk, sigma8_z, k0_low, "
                            "sigma_sq, X0, XX, YY) | ZeldovichPS(other)")

    def _init_state(self, cosmo, lowk, s8z, k0, ssq, X0, XX, YY):
        for a in (X0, XX, YY):
            if a.shape != (_ZEL_N,):
                raise ValueError("X0, XX, YY must have %d entries" % _ZEL_N)
        self._cosmo, self._lowk, self._s8z, self._k0, self._ssq = cosmo, lowk, s8z, k0, ssq
        self._X0, self._XX, self._YY = X0, XX, YY
        self._plin = None

    def __getstate__(self):
        return {'args': (self.args[0], self.GetApproxLowKFlag(), self.GetSigma8AtZ(), self.GetK0Low(), self.GetSigmaSq(),
                         self.GetX0Zel(), self.GetXZel(), self.GetYZel())}

    # accessors
    def GetCosmology(self): return self._cosmo
    def GetSigma8AtZ(self): return self._s8z
    def GetApproxLowKFlag(self): return self._lowk
    def GetK0Low(self): return self._k0
    def GetXZel(self): return self._XX.copy()
    def GetYZel(self): return self._YY.copy()
    def GetX0Zel(self): return self._X0.copy()
    def GetSigmaSq(self): return self._ssq
    def SetLowKApprox(self, approx_lowk_=True): self._lowk = bool(approx_lowk_)
    def SetLowKTransition(self, k0): self._k0 = float(k0)

    def SetSigma8AtZ(self, new_sigma8_z):
        # ZeldovichPS.cpp:95-109
        ratio = (new_sigma8_z/self._s8z)**2
        self._X0 *= ratio
        self._XX *= ratio
        self._YY *= ratio
        self._ssq *= ratio
        self._s8z = float(new_sigma8_z)

    def _lowk_inputs(self, k):
        if self._plin is None:
            self._plin = LinearPS(self._cosmo, 0.)
        if self._plin.GetSigma8AtZ() != self._s8z:
            self._plin.SetSigma8AtZ(self._s8z)
        plin = np.ascontiguousarray(self._plin(k).reshape(1, -1))
        q3 = np.array([self._plin.Q3_Zel(1.0)])
        return plin, q3

    def _evaluate(self, k, lowk):
        if self._index is None:
            raise RuntimeError("ZeldovichPS base class cannot be evaluated; use ZeldovichP00/P01/P11")
        ka, scalar = _scalar_or_array(k)
        out = np.empty((1, 3, len(ka)))
        plin_p = q3_p = None
        keep = []
        if lowk:
            plin, q3 = self._lowk_inputs(ka)
            keep = [plin, q3]
            plin_p, q3_p = plin.ctypes.data_as(_vp), q3.ctypes.data_as(_vp)
        X0 = np.ascontiguousarray(self._X0.reshape(1, -1))
        XX = np.ascontiguousarray(self._XX.reshape(1, -1))
        YY = np.ascontiguousarray(self._YY.reshape(1, -1))
        N.check(_L().prsd_zeldovich(self._cosmo._ctx.ptr, 1, X0, XX.ctypes.data_as(_vp), YY, np.array([self._ssq]), None,
                                    int(lowk), self._k0, plin_p, q3_p, len(ka), ka, out))
        del keep
        return _ret(out[0, self._index], scalar)

    def __call__(self, k):
        return self._evaluate(k, self._lowk)

    Evaluate = EvaluateMany = __call__

    def LowKApprox(self, k):
        """the SPT low-k form itself (ZeldovichPS.cpp:205-208, 229-232, 262-265)"""
        ka, scalar = _scalar_or_array(k)
        save = self._k0
        self._k0 = np.inf
        try:
            out = self._evaluate(ka, True)
        finally:
            self._k0 = save
        return _ret(np.atleast_1d(out), scalar)


class _ZelSpectrum(ZeldovichPS):
    def __init__(self, *args):
        ZeldovichPS.__init__(self, *args)
        self.args = args


class ZeldovichP00(_ZelSpectrum):
    _index = 0


class ZeldovichP01(_ZelSpectrum):
    _index = 1


class ZeldovichP11(_ZelSpectrum):
    _index = 2


class ZeldovichCF(PickalableSWIG):
    """``ZeldovichCF(C, z, kmin=1e-5, kmax=10)`` | ``ZeldovichCF(ZelPS, kmin, kmax)`` (ZeldovichCF.i:7-20):
    P00_zel (low-k approximation on) on logspace(kmin, kmax, 1024) -> xi(r) via pk_to_xi"""

    def __init__(self, *args):
        self.args = args
        if isinstance(args[0], ZeldovichPS):
            self._P = ZeldovichP00(args[0])
            rest = args[1:]
        else:
            self._P = ZeldovichP00(args[0], args[1], True)
            rest = args[2:]
        self._kmin = rest[0] if len(rest) > 0 else 1e-5
        self._kmax = rest[1] if len(rest) > 1 else 10.
        self._P.SetLowKApprox()
        self._sigma8_z = self._P.GetSigma8AtZ()
        # parray::logspace (parray.cpp:101-121)
        n = _ZEL_N
        self._k = np.exp(np.log(self._kmin) + np.arange(n)*(np.log(self._kmax) - np.log(self._kmin))/(n - 1))
        self._k[0], self._k[-1] = self._kmin, self._kmax
        self._grid = self._P(self._k)

    def SetSigma8AtZ(self, s):
        self._sigma8_z = float(s)

    def __call__(self, r, smoothing=0.):
        if self._sigma8_z != self._P.GetSigma8AtZ():
            self._P.SetSigma8AtZ(self._sigma8_z)
            self._grid = self._P(self._k)
        ra, scalar = _scalar_or_array(r)
        return _ret(pk_to_xi(0, self._k, self._grid, ra, smoothing), scalar)

    Evaluate = EvaluateMany = __call__


# ------------------------------------------------------------------- correlation functions
def ComputeXiLM(l, m, k, pk, r, smoothing=0., method=IntegrationMethods.FFTLOG):
    """xi_l^m(r) = int dk/(2 pi^2) k^m j_l(kr) P(k) (CorrelationFunction.i:16-37, .cpp:56-125); ``pk`` may hold a
    batch of spectra [nbatch][nk] on the shared ``k`` (the SWIG call takes one)"""
    k, pk, r = N.as_f64(k), np.ascontiguousarray(pk, dtype=np.float64), N.as_f64(r)
    batch = pk.reshape(-1, len(k))
    out = np.empty((batch.shape[0], len(r)))
    if method == IntegrationMethods.FFTLOG:
        N.check(_L().prsd_compute_xilm(N.default_context().ptr, batch.shape[0], int(l), int(m), len(k), k, batch, len(r), r,
                                       float(smoothing), 1.0, out))
    elif method in (IntegrationMethods.SIMPS, IntegrationMethods.TRAPZ):
        N.check(_L().prsd_compute_xilm_discrete(N.default_context().ptr, int(method), batch.shape[0], int(l), int(m), len(k), k,
                                                batch, len(r), r, float(smoothing), 1.0, out))
    else:
        raise ValueError("the integration method in `ComputeXiLM` must be one of [`fftlog=0`, `simps=1`, `trapz=2`]")
    return out[0] if pk.ndim == 1 else out


def _discrete(method, x, y):
    x, y = N.as_f64(x), np.ascontiguousarray(y, dtype=np.float64)
    batch = y.reshape(-1, len(x))
    out = np.empty(batch.shape[0])
    N.check(_L().prsd_discrete_integrate(N.default_context().ptr, method, batch.shape[0], len(x), x, batch, out))
    return float(out[0]) if y.ndim == 1 else out


def SimpsIntegrate(x, y):
    """Simpson's rule on arbitrary abscissae (DiscreteQuad.i; DiscreteQuad.cpp:3-28, 41-66); ``y`` may be [nbatch][n]"""
    return _discrete(1, x, y)


def TrapzIntegrate(x, y):
    """trapezoid rule (DiscreteQuad.i; DiscreteQuad.cpp:30-39); ``y`` may be [nbatch][n]"""
    return _discrete(2, x, y)


def pk_to_xi(ell, k, pk, r, smoothing=0.5, method=IntegrationMethods.FFTLOG):
    """(-1)^(ell/2) ComputeXiLM(ell, 2, ...) (CorrelationFunction.cpp:127-131)"""
    return (-1.)**(ell//2)*ComputeXiLM(ell, 2, k, pk, r, smoothing, method)


def xi_to_pk(ell, r, xi, k, smoothing=0.005, method=IntegrationMethods.FFTLOG):
    """(-1)^(ell/2) 8 pi^3 ComputeXiLM(ell, 2, r, xi, k, ...) (CorrelationFunction.cpp:133-141)"""
    return (-1.)**(ell//2)*8*np.pi**3*ComputeXiLM(ell, 2, r, xi, k, smoothing, method)


def ComputeXiLM_fftlog(l, m, k, pk, *args, **kwargs):
    """log-spaced input (``compute_xilm_fftlog`` of CorrelationFunction.i:28-37, .cpp:23-54).  Two call forms:

    * the reference's: ``ComputeXiLM_fftlog(l, m, k, pk, r, xi, q=0, smoothing=0)`` fills the caller's arrays ``r`` and ``xi``
      (length ``len(k)``) in place and returns None (rsd/window.py:238, 251);
    * ``ComputeXiLM_fftlog(l, m, k, pk, q=0, smoothing=0)`` returns ``(r, xi)``."""
    inplace = len(args) >= 2 and isinstance(args[0], np.ndarray) and isinstance(args[1], np.ndarray)
    if inplace:
        r_out, xi_out, args = args[0], args[1], args[2:]
        if r_out.shape != (len(k),) or xi_out.shape != (len(k),) or r_out.dtype != np.float64 or xi_out.dtype != np.float64:
            raise ValueError("ComputeXiLM_fftlog: `r` and `xi` must be float64 arrays of the length of `k`")
    q = args[0] if len(args) > 0 else kwargs.pop('q', 0.)
    smoothing = args[1] if len(args) > 1 else kwargs.pop('smoothing', 0.)
    if kwargs or len(args) > 2:
        raise TypeError("ComputeXiLM_fftlog: unexpected arguments")
    k, pk = N.as_f64(k), N.as_f64(pk)
    r, xi = np.empty(len(k)), np.empty((1, len(k)))
    N.check(_L().prsd_compute_xilm_fftlog(N.default_context().ptr, 1, int(l), int(m), len(k), k, pk.reshape(1, -1),
                                          float(q), float(smoothing), r, xi))
    if inplace:
        r_out[...] = r
        xi_out[...] = xi[0]
        return None
    return r, xi[0]


class CorrelationFunction(PickalableSWIG):
    """``CorrelationFunction(P, kmin=1e-4, kmax=10)`` (CorrelationFunction.i:39-55): xi(r) = 1/(2 pi^2 r) int_kmin^kmax
    k sin(k r) P(k) dk (CorrelationFunction.cpp:145-168, adaptive quadrature there).  For a ``LinearPS`` the integral is the
    device's product-integration operator on the knot table of P_L (the oscillating factor is integrated exactly against
    the interpolant, as for X(r), Y(r)); for any other spectrum P is sampled on 16384 log-spaced k and integrated by the
    device Simpson rule of ``ComputeXiLM(0, 2, ..., method=SIMPS)``."""
    def __init__(self, P, kmin=1e-4, kmax=1e1):
        self.args = (P, kmin, kmax)
        self._P, self._kmin, self._kmax = P, float(kmin), float(kmax)

    def __getstate__(self):
        return {'args': (self._P, self._kmin, self._kmax)}

    def __call__(self, r):
        ra, scalar = _scalar_or_array(r)
        if isinstance(self._P, LinearPS):
            return _ret(self._P._sp.correlation(ra, self._kmin, self._kmax)[0], scalar)
        k = np.logspace(np.log10(self._kmin), np.log10(self._kmax), 16384)
        return _ret(ComputeXiLM(0, 2, k, self._P(k), ra, 0., IntegrationMethods.SIMPS), scalar)

    Evaluate = EvaluateMany = __call__

    def GetPowerSpectrum(self): return self._P
    def GetKmin(self): return self._kmin
    def GetKmax(self): return self._kmax
    def SetKmin(self, kmin): self._kmin = float(kmin)
    def SetKmax(self, kmax): self._kmax = float(kmax)


class Kaiser(PickalableSWIG):
    """``Kaiser(P_L, f, b1=1)`` (Kaiser.i, Kaiser.cpp:9-62): linear redshift-space spectrum, its multipoles and theirs of
    the correlation function.  (As the reference: ``P_s`` is (1 + f mu^2)^2 P(k) -- no b1 -- while the multipole
    prefactors carry b1^2, Kaiser.cpp:14-31.)"""
    def __init__(self, P_L, f, b1=1.0):
        self.args = (P_L, f, b1)
        self._P, self._f, self._b1 = P_L, float(f), float(b1)

    def __getstate__(self):
        return {'args': (self._P, self._f, self._b1)}

    def P_s(self, k, mu):
        return (1. + self._f*np.asarray(mu, dtype=float)**2)**2*self._P(k)

    __call__ = P_s

    def _prefactor(self, ell):
        beta = self._f/self._b1
        pre = {0: 1. + (2./3.)*beta + (1./5.)*beta**2, 2: (4./3.)*beta + (4./7.)*beta**2, 4: (8./35.)*beta**2}.get(int(ell), 0.)
        return pre*self._b1**2

    def P_ell(self, ell, k):
        return self._prefactor(ell)*self._P(k)

    def Xi_ell(self, ell, r, Nk=32768, kmin=1e-5, kmax=100):
        ra, scalar = _scalar_or_array(r)
        if int(Nk) > 4096:
            raise ValueError("Kaiser.Xi_ell: the device ComputeXiLM takes at most 4096 samples (the reference's default is "
                             "Nk = 32768); pass Nk <= 4096")
        k = np.exp(np.linspace(np.log(kmin), np.log(kmax), int(Nk)))
        k[0], k[-1] = kmin, kmax                                     # parray::logspace (parray.cpp:101-121)
        xi = ComputeXiLM(int(ell), 2, k, self._P(k), ra, 0.)
        return _ret((-1.)**(int(ell)//2)*self._prefactor(ell)*xi, scalar)

    def SetGrowthRate(self, f): self._f = float(f)
    def GetGrowthRate(self): return self._f
    def SetLinearBias(self, b1): self._b1 = float(b1)
    def GetLinearBias(self): return self._b1
    def GetPowerSpectrum(self): return self._P


class FFTLog(object):
    """``FortranFFTLog(N, dlnr, mu=0.5, q=0, kr=1, kropt=1)`` (FortranFFTLog.h:9-34)"""
    def __init__(self, N_, dlnr, mu=0.5, q=0., kr=1., kropt=1):
        self.N, self.dlnr, self.mu, self.q, self.kropt = int(N_), float(dlnr), float(mu), float(q), int(kropt)
        self._kr_in = float(kr)
        self._kr = None

    def Transform(self, a, direction=1):
        if direction not in (1, -1):
            raise ValueError("FFTLog.Transform: direction must be +1 or -1")
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        batch = a.reshape(-1, self.N)
        out = np.empty_like(batch)
        kr = C.c_double()
        L = _L()
        L.prsd_fftlog_dir.argtypes = [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                      N._dp, N._dp, C.POINTER(C.c_double)]
        L.prsd_fftlog_dir.restype = C.c_int
        N.check(L.prsd_fftlog_dir(N.default_context().ptr, batch.shape[0], self.N, self.mu, self.q, self.dlnr, self._kr_in,
                                  self.kropt, int(direction), batch, out, C.byref(kr)))
        self._kr = kr.value
        return out.reshape(a.shape)

    def KR(self):
        if self._kr is None:
            self.Transform(np.zeros(self.N))
        return self._kr


Spline = object          # the SWIG base of LinearSpline / CubicSpline (Spline.i:5-22): both expose Evaluate / EvaluateDerivative


class LinearSpline(object):
    """``LinearSpline(x, y)`` (Spline.i:24, Spline.cpp:58-130): piecewise-linear, zero outside the domain"""
    def __init__(self, x, y):
        self._x, self._y = N.as_f64(x), N.as_f64(y)
        if len(self._x) != len(self._y) or len(self._x) < 2:
            raise ValueError("LinearSpline needs two arrays of the same length >= 2")

    def _eval(self, xq, deriv):
        xa, scalar = _scalar_or_array(xq)
        out = np.empty((1, len(xa)))
        L = _L()
        L.prsd_linear_spline_eval.argtypes = [_vp, C.c_int, C.c_int, N._dp, N._dp, C.c_int, N._dp, C.c_int, N._dp]
        L.prsd_linear_spline_eval.restype = C.c_int
        N.check(L.prsd_linear_spline_eval(N.default_context().ptr, 1, len(self._x), self._x, self._y.reshape(1, -1), len(xa), xa,
                                          int(deriv), out))
        return _ret(out[0], scalar)

    def __call__(self, xq):
        return self._eval(xq, 0)

    Evaluate = __call__

    def EvaluateDerivative(self, xq):
        return self._eval(xq, 1)


class CubicSpline(object):
    """``CubicSpline(x, y)`` (Spline.i; netlib spline of Spline.cpp:211-273, zero outside the domain)"""
    def __init__(self, x, y):
        self._x, self._y = N.as_f64(x), N.as_f64(y)

    def Evaluate(self, xq):
        return self(xq)

    def __call__(self, xq):
        xa, scalar = _scalar_or_array(xq)
        out = np.empty((1, len(xa)))
        N.check(_L().prsd_cubic_spline_eval(N.default_context().ptr, 1, len(self._x), self._x, self._y.reshape(1, -1),
                                            len(xa), xa, out))
        return _ret(out[0], scalar)
