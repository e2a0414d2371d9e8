"""TEST INFRASTRUCTURE ONLY: a stand-in for the SWIG-generated ``pyRSD.gcl`` module.

The reference's Python model layer (``pyRSD.rsd``) talks only to ``pyRSD.pygcl``, which
subclasses the SWIG classes of ``pyRSD.gcl`` (a build artefact that cannot be produced here: no
swig / CLASS / gfortran).  This module provides classes with the same names and call shapes,
backed by ``oracle/_ref/libgcl_ref.so`` -- i.e. by the reference's own, unmodified C++ -- so that
the UNMODIFIED reference Python (imported from /root/reference) can run in the build container
and produce golden P(k, mu) vectors (tests/golden/make_galaxy_ref.py).

Nothing here is used at test time on the GPU box, and nothing in the product imports it.
"""
import numpy as np

from oracle import gcl_ref as R


class _Transfers(object):
    CLASS, EH, EH_NoWiggle, BBKS, FromArrays = range(5)


transfers = _Transfers()


class _IntegrationMethods(object):
    FFTLOG, SIMPS, TRAPZ = range(3)


IntegrationMethods = _IntegrationMethods()


class Constants(object):
    # Common.i:36, 52 (cgs)
    c_light = 2.99792458e10
    km = 1e5


class ClassParams(dict):
    @classmethod
    def from_dict(cls, d):
        return cls(d)


class ClassEngine(object):
    pass


class Cosmology(object):
    """Cosmology(params, tf, sigma8, k, Tk) -- the pickle-state constructor (pygcl.py:60-63).

    Background quantities that come from CLASS in the reference (f, D, H, r_s) are supplied through
    ``params`` under the keys ``_f_z0``, ``_growth`` (dict z -> D), ``_f`` (dict z -> f)."""

    def __init__(self, params, tf=0, sigma8=None, k=None, Tk=None):
        self._params = ClassParams(params)
        self._tf = tf
        p = self._params
        h = p['h']
        Om = p['Omega_b'] + p['Omega_cdm'] + p.get('m_ncdm', 0.0)/93.14/h**2
        if k is None:
            # Cosmology(params) alone would run CLASS in the reference (hzpt/__init__.py:47 does this only to
            # hold a handle); here it is a shell without a transfer function
            self._ref, self._sigma8, self._k, self._Tk = None, sigma8, None, None
            return
        self._ref = R.RefCosmology(k, Tk, sigma8, p['n_s'], h, Om, p['Omega_b'], p.get('T_cmb', 2.7255))
        self._sigma8 = sigma8
        self._k, self._Tk = np.asarray(k), np.asarray(Tk)

    # accessors used by rsd
    def GetParams(self): return self._params
    def GetTransferFit(self): return self._tf
    def GetDiscreteK(self): return self._k
    def GetDiscreteTk(self): return self._Tk
    def sigma8(self): return self._sigma8
    def h(self): return self._params['h']
    def n_s(self): return self._params['n_s']
    def delta_H(self): return self._ref.delta_H()
    def D_z(self, z): return self._params.get('_growth', {0.0: 1.0})[float(z)]
    def f_z(self, z): return self._params.get('_f', {})[float(z)]
    def H_z(self, z): return self._params.get('_H', {0.0: 100.*self._params['h']})[float(z)]
    def Sigma8_z(self, z): return self._sigma8*self.D_z(z)
    def rs_drag(self): return self._params.get('_rs_drag', 147.48679403288168)
    def SetTransferFunction(self, tf): raise NotImplementedError("only the CLASS-table transfer function is available")

    # background densities CLASS would supply (ClassEngine.i:60-72); flat LCDM unless given in params
    def Omega0_m(self):
        p = self._params
        return p['Omega_b'] + p['Omega_cdm'] + p.get('m_ncdm', 0.0)/93.14/p['h']**2
    def Omega0_k(self): return self._params.get('Omega_k', 0.0)
    def Omega0_lambda(self): return self._params.get('Omega_Lambda', 1.0 - self.Omega0_m() - self.Omega0_k())

    def __getitem__(self, key):
        # pygcl.py:91-95
        if hasattr(self, key):
            f = getattr(self, key)
            if callable(f):
                return f()
        raise KeyError("Sorry, cannot return parameter '%s' in dict-like fashion" % key)


class _PS(object):
    """mixin for wrappers around a reference PowerSpectrum*"""
    def __call__(self, k):
        out = self._r(np.atleast_1d(np.asarray(k, dtype=float)))
        return out if np.ndim(k) else float(out[0])

    def VelocityDispersion(self, *args):
        if len(args) == 0:
            return self._r.VelocityDispersion()
        k = args[0]
        factor = args[1] if len(args) > 1 else 0.5
        out = self._r.VelocityDispersion(np.atleast_1d(np.asarray(k, dtype=float)), factor)
        return out if np.ndim(k) else float(out[0])

    def X_Zel(self, r): return self._r.X_Zel(r)
    def Y_Zel(self, r): return self._r.Y_Zel(r)
    def Q3_Zel(self, k): return self._r.Q3_Zel(k)
    def Sigma(self, R_): return self._r.Sigma(R_)


class LinearPS(_PS):
    def __init__(self, cosmo, z=0.):
        self._cosmo, self._z = cosmo, z
        self._s8z = cosmo.Sigma8_z(z)
        self._build()

    def _build(self):
        self._r = R.RefLinearPS(self._cosmo._ref, self._s8z)

    def GetRedshift(self): return self._z
    def GetCosmology(self): return self._cosmo
    def GetSigma8AtZ(self): return self._s8z

    def SetSigma8AtZ(self, s):
        self._s8z = s
        self._build()


class _OneLoop(_PS):
    which = None

    def __init__(self, P_L, epsrel=1e-4, oneloop_power=None):
        self._P_L, self._eps = P_L, epsrel
        self._r = R.RefOneLoop(P_L._r, self.which, epsrel, oneloop_power)

    def GetEpsrel(self): return self._eps
    def GetOneLoopPower(self): return self._r.table()
    def GetLinearPS(self): return self._P_L
    def EvaluateFull(self, k): return self._r.EvaluateFull(k)


class OneLoopPdd(_OneLoop):
    which = 'dd'


class OneLoopPdv(_OneLoop):
    which = 'dv'


class OneLoopPvv(_OneLoop):
    which = 'vv'


class OneLoopP22Bar(_OneLoop):
    which = '22bar'

    def VelocityKurtosis(self): return self._r.VelocityKurtosis()


class Imn(object):
    def __init__(self, P_L, epsrel=1e-4):
        self._P, self._eps = P_L, epsrel

    def __call__(self, k, m, n):
        return R.Imn(self._P._r, k, m, n, self._eps)


class Jmn(Imn):
    def __call__(self, k, m, n):
        return R.Jmn(self._P._r, k, m, n, self._eps)


class Kmn(object):
    def __init__(self, P_L, epsrel=1e-3):
        self._P, self._eps = P_L, epsrel

    def __call__(self, k, m, n, tidal=False, part=0):
        return R.Kmn(self._P._r, k, m, n, tidal, part, self._eps)


class ImnOneLoop(object):
    def __init__(self, P1, *args):
        self._P1 = P1
        self._P2 = None
        self._eps = 1e-4
        for a in args:
            if isinstance(a, _OneLoop):
                self._P2 = a
            else:
                self._eps = a

    def _ev(self, which, k, m, n):
        return R.ImnOneLoop(self._P1._r, None if self._P2 is None else self._P2._r, k, m, n, which, self._eps)

    def EvaluateLinear(self, k, m, n): return self._ev('linear', k, m, n)
    def EvaluateCross(self, k, m, n): return self._ev('cross', k, m, n)
    def EvaluateOneLoop(self, k, m, n): return self._ev('oneloop', k, m, n)


class ZeldovichPS(object):
    which = None

    def __init__(self, *args):
        if isinstance(args[0], ZeldovichPS):
            src = args[0]
            self._cosmo, self._r = src._cosmo, src._r        # shares the base state like the C++ copy
            self._lowk, self._k0 = src._lowk, src._k0
            self._s8z = src._s8z
            # the C++ copy constructor duplicates X0/XX/YY: emulate with a fresh state object
            ssq, X0, XX, YY = src._r.state()
            self._r = R.RefZeldovich(self._cosmo._ref, self._lowk, (self._s8z, self._k0, ssq, X0, XX, YY))
        elif len(args) == 8:
            cosmo, lowk, s8z, k0, ssq, X0, XX, YY = args
            self._cosmo, self._lowk, self._k0, self._s8z = cosmo, lowk, k0, s8z
            self._r = R.RefZeldovich(cosmo._ref, lowk, (s8z, k0, ssq, X0, XX, YY))
        else:
            cosmo, z = args[0], args[1]
            lowk = args[2] if len(args) > 2 else False
            self._cosmo, self._lowk, self._k0 = cosmo, lowk, 5e-3
            self._s8z = cosmo.Sigma8_z(z)
            cosmo._ref.set_sigma8_z(self._s8z)
            self._r = R.RefZeldovich(cosmo._ref, lowk)

    def GetCosmology(self): return self._cosmo
    def GetSigma8AtZ(self): return self._s8z
    def GetApproxLowKFlag(self): return self._lowk
    def GetK0Low(self): return self._k0
    def GetSigmaSq(self): return self._r.state()[0]
    def GetX0Zel(self): return self._r.state()[1]
    def GetXZel(self): return self._r.state()[2]
    def GetYZel(self): return self._r.state()[3]
    def SetLowKApprox(self, flag=True): self._lowk = flag
    def SetLowKTransition(self, k0): self._k0 = k0

    def SetSigma8AtZ(self, s):
        self._r.SetSigma8AtZ(s)
        self._s8z = s

    def __call__(self, k):
        out = self._r(self.which, np.atleast_1d(np.asarray(k, dtype=float)), int(bool(self._lowk)), self._k0)
        return out if np.ndim(k) else float(out[0])


class ZeldovichP00(ZeldovichPS):
    which = 'P00'


class ZeldovichP01(ZeldovichPS):
    which = 'P01'


class ZeldovichP11(ZeldovichPS):
    which = 'P11'


class ZeldovichCF(object):
    def __init__(self, *args):
        raise NotImplementedError


class CorrelationFunction(object):
    pass


class Kaiser(object):
    pass


class NonlinearPS(object):
    pass


class Spline(object):
    pass


def CubicSpline(x, y):
    return lambda xq: R.CubicSpline(x, y, xq)


def LinearSpline(x, y):
    return lambda xq: np.interp(xq, x, y)


def pk_to_xi(ell, k, pk, r, smoothing=0.5, method=0): return R.pk_to_xi(ell, k, pk, r, smoothing, method)
def xi_to_pk(ell, r, xi, k, smoothing=0.005, method=0): return R.xi_to_pk(ell, r, xi, k, smoothing, method)
def ComputeXiLM(l, m, k, pk, r, smoothing=0., method=0): return R.ComputeXiLM(l, m, k, pk, r, smoothing, method)
def compute_xilm_fftlog(l, m, k, pk, r, xi, q=0., smoothing=0.):
    """SWIG's in-place form (CorrelationFunction.i:28-37): fills r, xi"""
    rr, xx = R.ComputeXiLM_fftlog(l, m, k, pk, q, smoothing)
    r[...] = rr
    xi[...] = xx
def SimpsIntegrate(x, y): return R.SimpsIntegrate(x, y)
def TrapzIntegrate(x, y): return R.TrapzIntegrate(x, y)


class _CVar(object):
    """SWIG's holder of C++ static variables (pyRSD/__init__.py:60-71 sets CLASS data-file paths)"""
    pass


cvar = _CVar()
