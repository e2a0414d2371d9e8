"""TEST INFRASTRUCTURE ONLY -- ctypes handle on ``oracle/_ref/libgcl_ref.so``.

``libgcl_ref.so`` is the UNMODIFIED reference C++ for the hot path compiled by
``oracle/Makefile`` (plus the Cosmology stand-in and the C restatement of the
Fortran FFTLog).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this.
The product package ``pyrsd_b200`` never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_ref', 'libgcl_ref.so')

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')
_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libgcl_ref.so missing: run `make -C oracle` "
                               "in the build container (needs /root/reference)")
        L = C.CDLL(_LIB_PATH)
        vp, d, i = C.c_void_p, C.c_double, C.c_int
        sig = {
            'ref_cosmo_new': (vp, [i, _dp, _dp, d, d, d, d, d, d]),
            'ref_cosmo_free': (None, [vp]),
            'ref_cosmo_delta_H': (d, [vp]),
            'ref_cosmo_set_sigma8_z': (None, [vp, d]),
            'ref_cosmo_transfer': (None, [vp, i, _dp, _dp]),
            'ref_plin_new': (vp, [vp, d]),
            'ref_plin_free': (None, [vp]),
            'ref_ps_eval': (None, [vp, i, _dp, _dp]),
            'ref_ps_velocity_dispersion': (d, [vp]),
            'ref_ps_sigmasq_k': (None, [vp, d, i, _dp, _dp]),
            'ref_ps_sigma_R': (None, [vp, i, _dp, _dp]),
            'ref_correlation': (None, [vp, d, d, i, _dp, _dp]),
            'ref_ps_xzel': (None, [vp, i, _dp, _dp]),
            'ref_ps_yzel': (None, [vp, i, _dp, _dp]),
            'ref_ps_q3zel': (None, [vp, i, _dp, _dp]),
            'ref_ps_sigma3sq': (None, [vp, i, _dp, _dp]),
            'ref_ps_xzel_tol': (None, [vp, i, _dp, d, d, _dp]),
            'ref_ps_yzel_tol': (None, [vp, i, _dp, d, d, _dp]),
            'ref_ps_sigv_tol': (d, [vp, d, d, d]),
            'ref_ps_q3_tol': (d, [vp, d, d, d]),
            'ref_p22bar_kurtosis_tol': (d, [vp, d, d]),
            'ref_ps_sigma_R_tol': (d, [vp, d, d, d]),
            'ref_imn': (None, [vp, d, i, i, i, _dp, _dp]),
            'ref_jmn': (None, [vp, d, i, i, i, _dp, _dp]),
            'ref_kmn': (None, [vp, d, i, i, i, i, i, _dp, _dp]),
            'ref_oneloop_new': (vp, [vp, i, d]),
            'ref_oneloop_from_table': (vp, [vp, i, d, i, _dp]),
            'ref_oneloop_free': (None, [vp]),
            'ref_oneloop_table': (None, [vp, _dp]),
            'ref_oneloop_eval': (None, [vp, i, i, _dp, _dp]),
            'ref_p22bar_kurtosis': (d, [vp]),
            'ref_imn_oneloop': (None, [vp, vp, d, i, i, i, i, _dp, _dp]),
            'ref_zel_new_full': (vp, [vp, i]),
            'ref_zel_new': (vp, [vp, i, d, d, d, i, _dp, _dp, _dp]),
            'ref_zel_free': (None, [vp]),
            'ref_zel_state': (None, [vp, C.POINTER(d), _dp, _dp, _dp]),
            'ref_zel_set_sigma8': (None, [vp, d]),
            'ref_zel_eval': (None, [vp, i, i, d, i, _dp, _dp]),
            'ref_zelcf_eval': (None, [vp, d, d, d, i, _dp, _dp]),
            'ref_pk_to_xi': (None, [i, i, _dp, _dp, i, _dp, d, i, _dp]),
            'ref_xi_to_pk': (None, [i, i, _dp, _dp, i, _dp, d, i, _dp]),
            'ref_compute_xilm': (None, [i, i, i, _dp, _dp, i, _dp, d, i, _dp]),
            'ref_compute_xilm_fftlog': (None, [i, i, i, _dp, _dp, _dp, _dp, d, d]),
            'ref_fftlog': (d, [i, d, d, d, d, i, i, _dp]),
            'ref_simps': (d, [i, _dp, _dp]),
            'ref_trapz': (d, [i, _dp, _dp]),
            'ref_cubic_spline': (None, [i, _dp, _dp, i, _dp, _dp]),
            'ref_logspace': (None, [d, d, i, _dp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _a(x):
    return np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)


ONELOOP = {'dd': 0, 'dv': 1, 'vv': 2, '22bar': 3}


class RefCosmology(object):
    """Stand-in for pygcl.Cosmology fed with a discrete transfer function."""

    def __init__(self, k, T, sigma8, n_s, h, Omega0_m, Omega0_b, Tcmb):
        k, T = _a(k), _a(T)
        self.args = (k, T, sigma8, n_s, h, Omega0_m, Omega0_b, Tcmb)
        self.ptr = lib().ref_cosmo_new(len(k), k, T, sigma8, n_s, h, Omega0_m, Omega0_b, Tcmb)

    def delta_H(self):
        return lib().ref_cosmo_delta_H(self.ptr)

    def sigma8(self):
        return self.args[2]

    def set_sigma8_z(self, s):
        lib().ref_cosmo_set_sigma8_z(self.ptr, s)

    def transfer(self, k):
        k = _a(k)
        out = np.empty_like(k)
        lib().ref_cosmo_transfer(self.ptr, len(k), k, out)
        return out


class RefPower(object):
    """Any reference PowerSpectrum* (LinearPS or OneLoopPS)."""
    ptr = None

    def __call__(self, k):
        k = _a(k)
        out = np.empty_like(k)
        lib().ref_ps_eval(self.ptr, len(k), k, out)
        return out

    def VelocityDispersion(self, k=None, factor=0.5):
        if k is None:
            return lib().ref_ps_velocity_dispersion(self.ptr)
        k = _a(k)
        out = np.empty_like(k)
        lib().ref_ps_sigmasq_k(self.ptr, factor, len(k), k, out)
        return out

    def _vec(self, fn, x):
        x = _a(x)
        out = np.empty_like(x)
        getattr(lib(), fn)(self.ptr, len(x), x, out)
        return out

    def Sigma(self, R):
        return self._vec('ref_ps_sigma_R', R)

    def X_Zel(self, r, epsrel=None, epsabs=1e-10):
        if epsrel is None:
            return self._vec('ref_ps_xzel', r)
        r = _a(r)
        out = np.empty_like(r)
        lib().ref_ps_xzel_tol(self.ptr, len(r), r, epsrel, epsabs, out)
        return out

    def Y_Zel(self, r, epsrel=None, epsabs=1e-10):
        if epsrel is None:
            return self._vec('ref_ps_yzel', r)
        r = _a(r)
        out = np.empty_like(r)
        lib().ref_ps_yzel_tol(self.ptr, len(r), r, epsrel, epsabs, out)
        return out

    def Q3_Zel(self, k):
        return self._vec('ref_ps_q3zel', k)

    # tight-tolerance variants (the reference hard-codes 1e-4..1e-5 in these integrals)
    def sigv_tol(self, qmax=100.0, epsrel=1e-10, epsabs=1e-14):
        return lib().ref_ps_sigv_tol(self.ptr, qmax, epsrel, epsabs)

    def Q3_tol(self, k, epsrel=1e-10, epsabs=1e-14):
        return np.array([lib().ref_ps_q3_tol(self.ptr, float(kk), epsrel, epsabs) for kk in _a(k)])

    def Sigma_tol(self, R, epsrel=1e-10, epsabs=1e-14):
        return np.array([lib().ref_ps_sigma_R_tol(self.ptr, float(r), epsrel, epsabs) for r in _a(R)])

    def sigma3_squared(self, k):
        return self._vec('ref_ps_sigma3sq', k)

    def correlation(self, r, kmin=1e-4, kmax=1e1):
        """CorrelationFunction(P, kmin, kmax)(r)"""
        r = _a(r)
        out = np.empty_like(r)
        lib().ref_correlation(self.ptr, kmin, kmax, len(r), r, out)
        return out


class RefLinearPS(RefPower):
    def __init__(self, cosmo, sigma8_z=None):
        self.cosmo = cosmo
        self.sigma8_z = cosmo.sigma8() if sigma8_z is None else sigma8_z
        self.ptr = lib().ref_plin_new(cosmo.ptr, self.sigma8_z)


def Imn(P, k, m, n, epsrel=1e-4):
    k = _a(k)
    out = np.empty_like(k)
    lib().ref_imn(P.ptr, epsrel, m, n, len(k), k, out)
    return out


def Jmn(P, k, m, n, epsrel=1e-4):
    k = _a(k)
    out = np.empty_like(k)
    lib().ref_jmn(P.ptr, epsrel, m, n, len(k), k, out)
    return out


def Kmn(P, k, m, n, tidal=False, part=0, epsrel=1e-3):
    k = _a(k)
    out = np.empty_like(k)
    lib().ref_kmn(P.ptr, epsrel, m, n, int(tidal), part, len(k), k, out)
    return out


class RefOneLoop(RefPower):
    def __init__(self, P_L, which, epsrel=1e-4, table=None):
        self.P_L = P_L
        w = ONELOOP[which]
        if table is None:
            self.ptr = lib().ref_oneloop_new(P_L.ptr, w, epsrel)
        else:
            table = _a(table)
            self.ptr = lib().ref_oneloop_from_table(P_L.ptr, w, epsrel, len(table), table)

    def table(self):
        out = np.empty(1000)
        lib().ref_oneloop_table(self.ptr, out)
        return out

    def EvaluateFull(self, k):
        k = _a(k)
        out = np.empty_like(k)
        lib().ref_oneloop_eval(self.ptr, 1, len(k), k, out)
        return out

    def VelocityKurtosis(self, epsrel=None, epsabs=1e-14):
        if epsrel is None:
            return lib().ref_p22bar_kurtosis(self.ptr)
        return lib().ref_p22bar_kurtosis_tol(self.ptr, epsrel, epsabs)


def ImnOneLoop(P1, P2, k, m, n, which, epsrel=1e-4):
    """which in {'linear','cross','oneloop'}; P2 may be None (equal spectra)."""
    w = {'linear': 0, 'cross': 1, 'oneloop': 2}[which]
    k = _a(k)
    out = np.empty_like(k)
    lib().ref_imn_oneloop(P1.ptr, None if P2 is None else P2.ptr, epsrel, w, m, n, len(k), k, out)
    return out


class RefZeldovich(object):
    NUM_PTS = 1024

    def __init__(self, cosmo, approx_lowk=False, state=None):
        """state = (sigma8_z, k0_low, sigmasq, X0, XX, YY) or None for the full ctor."""
        self.cosmo = cosmo
        if state is None:
            self.ptr = lib().ref_zel_new_full(cosmo.ptr, int(approx_lowk))
        else:
            s8z, k0, ssq, X0, XX, YY = state
            X0, XX, YY = _a(X0), _a(XX), _a(YY)
            self.ptr = lib().ref_zel_new(cosmo.ptr, int(approx_lowk), s8z, k0, ssq, len(X0), X0, XX, YY)

    def state(self):
        ssq = C.c_double()
        X0, XX, YY = (np.empty(self.NUM_PTS) for _ in range(3))
        lib().ref_zel_state(self.ptr, C.byref(ssq), X0, XX, YY)
        return ssq.value, X0, XX, YY

    def SetSigma8AtZ(self, s):
        lib().ref_zel_set_sigma8(self.ptr, s)

    def __call__(self, which, k, lowk=-1, k0_low=-1.0):
        w = {'P00': 0, 'P01': 1, 'P11': 2}[which]
        k = _a(k)
        out = np.empty_like(k)
        lib().ref_zel_eval(self.ptr, w, lowk, k0_low, len(k), k, out)
        return out

    def cf(self, r, smoothing=0.0, kmin=1e-5, kmax=10.0):
        r = _a(r)
        out = np.empty_like(r)
        lib().ref_zelcf_eval(self.ptr, kmin, kmax, smoothing, len(r), r, out)
        return out


def pk_to_xi(ell, k, pk, r, smoothing=0.5, method=0):
    k, pk, r = _a(k), _a(pk), _a(r)
    out = np.empty_like(r)
    lib().ref_pk_to_xi(ell, len(k), k, pk, len(r), r, smoothing, method, out)
    return out


def xi_to_pk(ell, r, xi, k, smoothing=0.005, method=0):
    r, xi, k = _a(r), _a(xi), _a(k)
    out = np.empty_like(k)
    lib().ref_xi_to_pk(ell, len(r), r, xi, len(k), k, smoothing, method, out)
    return out


def ComputeXiLM(l, m, k, pk, r, smoothing=0.0, method=0):
    k, pk, r = _a(k), _a(pk), _a(r)
    out = np.empty_like(r)
    lib().ref_compute_xilm(l, m, len(k), k, pk, len(r), r, smoothing, method, out)
    return out


def ComputeXiLM_fftlog(l, m, k, pk, q=0.0, smoothing=0.0):
    k, pk = _a(k), _a(pk)
    r, xi = np.empty_like(k), np.empty_like(k)
    lib().ref_compute_xilm_fftlog(l, m, len(k), k, pk, r, xi, q, smoothing)
    return r, xi


def fftlog(a, dlnr, mu, q=0.0, kr=1.0, kropt=1, direction=1):
    a = _a(a).copy()
    kr_out = lib().ref_fftlog(len(a), dlnr, mu, q, kr, kropt, direction, a)
    return a, kr_out


def SimpsIntegrate(x, y):
    x, y = _a(x), _a(y)
    return lib().ref_simps(len(x), x, y)


def TrapzIntegrate(x, y):
    x, y = _a(x), _a(y)
    return lib().ref_trapz(len(x), x, y)


def CubicSpline(x, y, xq):
    x, y, xq = _a(x), _a(y), _a(xq)
    out = np.empty_like(xq)
    lib().ref_cubic_spline(len(x), x, y, len(xq), xq, out)
    return out


def logspace(a, b, n):
    out = np.empty(n)
    lib().ref_logspace(a, b, n, out)
    return out


def golden_cosmology(golden):
    """RefCosmology for the Planck15 transfer function stored in the golden npz.

    Omega0_m = Omega_b + Omega_cdm + Omega_ncdm(m=0.06 eV) (CLASS background value,
    ClassEngine.cpp:61-62) and T_cmb from Omega_g; both enter only the Eisenstein-Hu
    extrapolation outside the tabulated k range (Cosmology.cpp:187-199)."""
    h = float(golden['cosmo_h'])
    Om = float(golden['cosmo_Omega_b']) + float(golden['cosmo_Omega_cdm']) \
        + float(golden['cosmo_m_ncdm'])/93.14/h**2
    return RefCosmology(golden['cosmo_k'], golden['cosmo_T'], float(golden['cosmo_sigma8']),
                        float(golden['cosmo_n_s']), h, Om, float(golden['cosmo_Omega_b']), 2.7255)
