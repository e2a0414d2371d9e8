"""TEST INFRASTRUCTURE ONLY: ctypes handle on the CPU emulation of the host logic."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')


class Emul(object):
    def __init__(self, L):
        self.L = L
        L.emul_knot_count.restype = C.c_int
        L.emul_knots.argtypes = [dp]
        L.emul_set_config.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
        L.emul_plin_table.argtypes = [C.c_int, dp, dp] + [C.c_double]*6 + [dp]
        L.emul_plin_eval.argtypes = [C.c_int, dp, dp] + [C.c_double]*6 + [C.c_int, dp, dp]
        L.emul_table_read.argtypes = [dp, C.c_int, dp, dp]
        L.emul_spline.argtypes = [C.c_int, dp, dp, C.c_int, dp, dp]
        L.emul_ik.argtypes = [C.c_int, C.c_int, dp, C.c_double, dp, dp, C.c_int, dp, C.POINTER(C.c_longlong)]
        L.emul_linear.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
        self.M = L.emul_knot_count()

    def knots(self):
        x = np.empty(self.M)
        self.L.emul_knots(x)
        return x

    def plin_table(self, k, T, ns, amp, h, Om, Ob, Tcmb):
        tab = np.empty(self.M)
        self.L.emul_plin_table(len(k), np.ascontiguousarray(k), np.ascontiguousarray(T), ns, amp, h, Om, Ob, Tcmb, tab)
        return tab

    def plin_eval(self, k, T, ns, amp, h, Om, Ob, Tcmb, q):
        q = np.ascontiguousarray(q, dtype='f8')
        out = np.empty_like(q)
        self.L.emul_plin_eval(len(k), np.ascontiguousarray(k), np.ascontiguousarray(T), ns, amp, h, Om, Ob, Tcmb, len(q), q, out)
        return out

    def fill_map(self, which, k):
        """keep mask [n], stencil indices [n][8] (into the kept points) and weights [n][8] of the thinned-list fill maps"""
        k = np.ascontiguousarray(k, dtype='f8')
        n = len(k)
        keep = np.zeros(n, dtype=np.int8)
        idx = np.zeros((n, 8), dtype=np.int32)
        w = np.zeros((n, 8))
        self.L.emul_fill_map.restype = C.c_int
        self.L.emul_fill_map.argtypes = [C.c_int, C.c_int, dp, np.ctypeslib.ndpointer(dtype=np.int8, flags='C_CONTIGUOUS'),
                                         np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS'), dp]
        nsub = self.L.emul_fill_map(which, n, k, keep, idx, w)
        return keep.astype(bool), idx, w, nsub

    def table_read(self, tab, q):
        q = np.ascontiguousarray(q, dtype='f8')
        out = np.empty_like(q)
        self.L.emul_table_read(np.ascontiguousarray(tab), len(q), q, out)
        return out

    def spline(self, x, y, xq):
        xq = np.ascontiguousarray(xq, dtype='f8')
        out = np.empty_like(xq)
        self.L.emul_spline(len(x), np.ascontiguousarray(x, 'f8'), np.ascontiguousarray(y, 'f8'), len(xq), xq, out)
        return out

    def ik(self, kid, k, tabQ, tabR, mode, qmax=1e5):
        k = np.ascontiguousarray(k, dtype='f8')
        out = np.empty_like(k)
        nn = C.c_longlong()
        self.L.emul_ik(kid, len(k), k, qmax, np.ascontiguousarray(tabQ), np.ascontiguousarray(tabR), mode, out, C.byref(nn))
        return out, nn.value

    def linear(self, type_, sub, param, lo, hi, tab, scale=1.0):
        param = np.ascontiguousarray(np.atleast_1d(np.asarray(param, float)))
        n = len(param)
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, float), (n,)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, float), (n,)))
        sc = np.ascontiguousarray(np.broadcast_to(np.asarray(scale, float), (n,)))
        out = np.empty(n)
        self.L.emul_linear(type_, sub, n, param, lo, hi, sc, np.ascontiguousarray(tab), out)
        return out


def load():
    path = os.path.join(HERE, 'libprsd_emul.so')
    if not os.path.exists(path):
        from pyrsd_b200 import build
        build.build_emul()
    return Emul(C.CDLL(path))
