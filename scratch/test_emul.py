import sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import gcl_ref as R
g = np.load('tests/golden/golden_pt.npz')
cosmo = R.golden_cosmology(g); PL = R.RefLinearPS(cosmo)
L = C.CDLL('tests/host_emul/libprsd_emul.so')
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')
L.emul_knot_count.restype = C.c_int
L.emul_knots.argtypes=[dp]
L.emul_plin_table.argtypes=[C.c_int, dp, dp]+[C.c_double]*6+[dp]
L.emul_plin_eval.argtypes=[C.c_int, dp, dp]+[C.c_double]*6+[C.c_int, dp, dp]
L.emul_table_read.argtypes=[dp, C.c_int, dp, dp]
L.emul_ik.argtypes=[C.c_int, C.c_int, dp, C.c_double, dp, dp, C.c_int, dp, C.POINTER(C.c_longlong)]
M = L.emul_knot_count(); print('M', M)
k_i, T_i, s8, ns, h, Om, Ob, Tc = cosmo.args
amp = cosmo.delta_H()**2
tab = np.empty(M)
L.emul_plin_table(len(k_i), k_i, T_i, ns, amp, h, Om, Ob, Tc, tab)
# check plin_eval vs oracle
q = np.exp(np.random.default_rng(0).uniform(np.log(1e-8), np.log(1e5), 5000))
out = np.empty_like(q); L.emul_plin_eval(len(k_i), k_i, T_i, ns, amp, h, Om, Ob, Tc, len(q), q, out)
ex = PL(q); print('plin_eval max rel', np.max(np.abs(out/ex-1)))
L.emul_table_read(tab, len(q), q, out); print('table read max rel', np.max(np.abs(out/ex-1)), 'median', np.median(np.abs(out/ex-1)))

ref = np.load('scratch/ref_tight.npy', allow_pickle=True).item()
ks = np.array([1e-4, 1e-3, 5e-3, 0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 0.8, 1.0, 3.0, 10.0])
Pk = PL(ks)
kid = {'I00':0,'I22':10,'I13':7,'I33':15,'K00':16,'K11':22,'K20s_b':28,'K01':17}
for name in ref:
    out = np.empty_like(ks); nn = C.c_longlong()
    t=time.time()
    L.emul_ik(kid[name], len(ks), ks, 1e5, tab, tab, 0 if name[0]=='I' else 1, out, C.byref(nn))
    err = np.abs(out-ref[name])/np.maximum(np.abs(ref[name]), Pk)
    print('%-7s'%name, ' '.join('%.0e'%e for e in err), 'nodes/k', nn.value//len(ks), '%.2fs'%(time.time()-t))
