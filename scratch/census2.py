import sys
sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from tests.host_emul import emul_lib
E = emul_lib.load()
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS'); dp = emul_lib.dp
E.L.emul_nodes.restype = C.c_longlong
E.L.emul_nodes.argtypes = [C.c_int, dp, C.c_double, C.c_longlong, ip, ip, dp]
def census(ks, qmax, ml):
    E.L.emul_set_config_default()
    if ml: E.L.emul_use_ml_variant()
    tot = 0; units = 0; single = 0; hist = {}
    for k in ks:
        cap = 200000
        o = np.empty(cap, np.int32); j = np.empty(cap, np.int32); w = np.empty(cap)
        n = E.L.emul_nodes(1, np.array([k]), qmax, cap, o, j, w)
        key = o[:n].astype(np.int64)*100000 + j[:n]
        u, cnt = np.unique(key, return_counts=True)
        tot += n
        single += (cnt == 1).sum()
        units += 4*(cnt >= 2).sum()
        for c in cnt: hist[c] = hist.get(c, 0) + 1
    return tot, single, units, hist
k1 = np.exp(np.log(1e-5)+np.arange(1000)*np.log(1e6)/999)[::10]
ki = np.logspace(np.log10(5e-6), 0, 250)[::3]
for name, ks, qmax, ml in (('1loop', k1, 1e5, False), ('ik', ki, 1e5, False), ('ml', ki, 10., True)):
    t, s, u, h = census(ks, qmax, ml)
    n = len(ks)
    reads_now = 4*t; reads_new = 4*s + u
    print('%s: nodes/k %.0f  singletons %.0f  unit pseudo-nodes %.0f -> total %.0f (%.2fx nodes), table reads/slot %.0f -> %.0f (%.2fx fewer)' % (name, t/n, s/n, u/n, (s+u)/n, (s+u)/t, reads_now/n, reads_new/n, reads_now/reads_new))
    print('   item-size histogram', dict(sorted(h.items())))
