import sys
sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from tests.host_emul import emul_lib
E = emul_lib.load()
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS'); dp = emul_lib.dp
E.L.emul_nodes.restype = C.c_longlong
E.L.emul_nodes.argtypes = [C.c_int, dp, C.c_double, C.c_longlong, ip, ip, dp]
def census(ks, qmax):
    tot = 0; comp = 0
    for k in ks:
        cap = 200000
        o = np.empty(cap, np.int32); j = np.empty(cap, np.int32); w = np.empty(cap)
        n = E.L.emul_nodes(1, np.array([k]), qmax, cap, o, j, w)
        o, j = o[:n], j[:n]
        key = o.astype(np.int64)*100000 + j
        u, cnt = np.unique(key, return_counts=True)
        c = np.minimum(cnt, 4).sum()
        tot += n; comp += c
    return tot, comp
k1 = np.exp(np.log(1e-5)+np.arange(1000)*np.log(1e6)/999)[::20]
ki = np.logspace(np.log10(5e-6), 0, 250)[::5]
for name, ks, qmax in (('1loop', k1, 1e5), ('ik', ki, 1e5), ('ml', ki, 10.)):
    t, c = census(ks, qmax)
    print(name, 'nodes', t/len(ks), 'after (outer, j0) merge', c/len(ks), 'ratio %.2f'%(t/c))
