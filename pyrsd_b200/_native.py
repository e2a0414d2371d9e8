"""ctypes binding of ``libpyrsd_b200.so`` (the C-ABI in ``include/pyrsd_b200.h``).

There is NO CPU fallback: if the shared library is missing, or no sm_100 device is
usable, the error propagates to the caller.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpyrsd_b200.so')

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')
_dp_or_null = C.c_void_p
_lib = None


class NativeError(RuntimeError):
    """Raised for any non-zero status of the C-ABI (mirrors the reference's translation of
    C++ exceptions into Python exceptions, pyRSD/gcl.i:22-30)."""

    def __init__(self, code, msg):
        RuntimeError.__init__(self, "pyrsd_b200 [%d]: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -m pyrsd_b200.build` "
                          "(the CUDA library is mandatory, there is no CPU path)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, d, i, ll = C.c_void_p, C.c_double, C.c_int, C.c_longlong
    pvp = C.POINTER(C.c_void_p)
    sig = {
        'prsd_version': (i, []),
        'prsd_ctx_create': (i, [i, pvp]),
        'prsd_ctx_destroy': (i, [vp]),
        'prsd_ctx_sync': (i, [vp]),
        'prsd_ctx_launches': (i, [vp, C.POINTER(ll)]),
        'prsd_ctx_stream': (i, [vp, pvp]),
        'prsd_knot_grid': (i, [C.POINTER(i), vp]),
        'prsd_quad_config_default': (i, [_dp]),
        'prsd_spectra_create': (i, [vp, i, i, _dp, _dp, _dp, pvp]),
        'prsd_spectra_destroy': (i, [vp]),
        'prsd_spectra_rescale': (i, [vp, _dp]),
        'prsd_spectra_plin': (i, [vp, i, _dp, _dp]),
        'prsd_spectra_table': (i, [vp, i, _dp]),
        'prsd_spectra_set_oneloop': (i, [vp, _dp]),
        'prsd_spectra_oneloop_eval': (i, [vp, i, i, i, _dp, _dp]),
        'prsd_eval_imn': (i, [vp, i, i, i, _dp, vp, _dp]),
        'prsd_eval_kmn': (i, [vp, i, i, i, i, i, _dp, vp, _dp]),
        'prsd_eval_jmn': (i, [vp, i, i, i, _dp, _dp]),
        'prsd_eval_imn_oneloop': (i, [vp, i, i, i, i, i, i, _dp, vp, _dp]),
        'prsd_eval_ps_integral': (i, [vp, i, i, _dp, d, _dp]),
        'prsd_eval_kurtosis': (i, [vp, _dp]),
        'prsd_eval_correlation': (i, [vp, i, _dp, d, d, _dp]),
        'prsd_plan_create': (i, [vp, i, _dp, vp, i, pvp]),
        'prsd_plan_destroy': (i, [vp]),
        'prsd_plan_info': (i, [vp, _dp]),
        'prsd_plan_run': (i, [vp, vp, pvp]),
        'prsd_result_destroy': (i, [vp]),
        'prsd_result_get': (i, [vp, i, _dp, ll]),
        'prsd_plan_run_async': (i, [vp, vp, pvp]),
        'prsd_spectra_reload': (i, [vp, vp, vp, vp]),
        'prsd_galaxy_power_async': (i, [vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, C.POINTER(ll)]),
        'prsd_galaxy_wait': (i, [vp]),
        'prsd_galaxy_wait_copy': (i, [vp, ll]),
        'prsd_host_alloc': (i, [C.c_size_t, pvp]),
        'prsd_host_free': (i, [vp]),
    }
    L.prsd_last_error.restype = C.c_char_p
    L.prsd_last_error.argtypes = []
    for name, (res, args) in sig.items():
        if not hasattr(L, name):
            continue            # optional entry points are checked by tests/test_abi.py
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(status):
    if status != 0:
        msg = lib().prsd_last_error()
        raise NativeError(status, msg.decode('utf-8', 'replace') if msg else 'unknown error')


def as_f64(x):
    return np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)


def cfg_ptr(cfg):
    """cfg: None or array of 56 doubles (include/pyrsd_b200.h) -> (keepalive, void*)"""
    if cfg is None:
        return None, None
    a = as_f64(cfg)
    if a.size != 56:
        raise ValueError("quadrature configuration must hold 56 numbers (see prsd_quad_config_default)")
    return a, a.ctypes.data_as(C.c_void_p)


class PinnedArray(object):
    """page-locked host memory (prsd_host_alloc) viewed as a NumPy float64 / int32 array; freed with the object"""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(int(x) for x in np.atleast_1d(shape))
        dt = np.dtype(dtype)
        nbytes = max(int(np.prod(self.shape))*dt.itemsize, 8)
        p = C.c_void_p()
        check(lib().prsd_host_alloc(nbytes, C.byref(p)))
        self.ptr = p
        buf = (C.c_char*nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dt, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if self.ptr is not None and _lib is not None:
            self.array = None
            _lib.prsd_host_free(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def knot_grid():
    n = C.c_int()
    check(lib().prsd_knot_grid(C.byref(n), None))
    x = np.empty(n.value)
    check(lib().prsd_knot_grid(C.byref(n), x.ctypes.data_as(C.c_void_p)))
    return x


def quad_config_default():
    c = np.empty(56)
    check(lib().prsd_quad_config_default(c))
    return c


class Handle(object):
    """Owns an opaque C handle and destroys it exactly once."""
    _destroy = None

    def __init__(self, ptr):
        self.ptr = ptr

    def close(self):
        if self.ptr is not None and _lib is not None:
            getattr(_lib, self._destroy)(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context(Handle):
    _destroy = 'prsd_ctx_destroy'

    def __init__(self, device=0):
        p = C.c_void_p()
        check(lib().prsd_ctx_create(int(device), C.byref(p)))
        Handle.__init__(self, p)
        self.device = int(device)

    def sync(self):
        check(lib().prsd_ctx_sync(self.ptr))

    def launches(self):
        n = C.c_longlong()
        check(lib().prsd_ctx_launches(self.ptr, C.byref(n)))
        return n.value


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
