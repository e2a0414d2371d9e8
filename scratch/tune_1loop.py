import sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import gcl_ref as R
g = np.load('tests/golden/golden_pt.npz')
cosmo = R.golden_cosmology(g); PL = R.RefLinearPS(cosmo)
L = C.CDLL('tests/host_emul/libprsd_emul.so')
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')
L.emul_knot_count.restype = C.c_int
L.emul_plin_table.argtypes=[C.c_int, dp, dp]+[C.c_double]*6+[dp]
L.emul_ik.argtypes=[C.c_int, C.c_int, dp, C.c_double, dp, dp, C.c_int, dp, C.POINTER(C.c_longlong)]
L.emul_set_config.argtypes=[C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
M = L.emul_knot_count()
k_i, T_i, s8, ns, h, Om, Ob, Tc = cosmo.args
tab = np.empty(M); L.emul_plin_table(len(k_i), k_i, T_i, ns, cosmo.delta_H()**2, h, Om, Ob, Tc, tab)
ks = np.array([0.04, 0.16, 0.3, 0.64, 1.27, 2.5, 5.0, 10.0])
import os
if not os.path.exists('scratch/ref_tight2.npy'):
    ref = {'I00': R.Imn(PL, ks, 0,0, epsrel=1e-11), 'I11': R.Imn(PL, ks, 1,1, epsrel=1e-11), 'I33': R.Imn(PL, ks, 3,3, epsrel=1e-11)}
    np.save('scratch/ref_tight2.npy', ref, allow_pickle=True)
ref = np.load('scratch/ref_tight2.npy', allow_pickle=True).item()
kid={'I00':0,'I11':5,'I33':15}
def run(label, qmin=1e-8, no=8, ni=8, ln_far=2.0, ln_mid=0.8, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05):
    L.emul_set_config(qmin, no, ni, ln_far, ln_mid, bao_lo, bao_hi, bao_dq)
    print(label)
    for name in ['I00','I11']:
        out = np.empty_like(ks); nn = C.c_longlong()
        L.emul_ik(kid[name], len(ks), ks, 1e5, tab, tab, 0, out, C.byref(nn))
        print('  %s'%name, ' '.join('%.0e'%e for e in np.abs(out/ref[name]-1)), 'nodes/k', nn.value//len(ks))
run('default')
run('ln_mid .5', ln_mid=0.5)
run('ln_mid .4 n10', ln_mid=0.4, no=10, ni=10)
run('n12', no=12, ni=12)
run('dq .03', bao_dq=0.03)
run('dq .03 ln_mid .5 n10', bao_dq=0.03, ln_mid=0.5, no=10, ni=10)
L.emul_ik_exactP.argtypes=[C.c_int, C.c_int, dp, C.c_double, C.c_int, dp, dp]+[C.c_double]*6+[dp]
def run_exact(label, qmin=1e-8, no=8, ni=8, ln_far=2.0, ln_mid=0.8, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05):
    L.emul_set_config(qmin, no, ni, ln_far, ln_mid, bao_lo, bao_hi, bao_dq)
    out = np.empty_like(ks)
    L.emul_ik_exactP(0, len(ks), ks, 1e5, len(k_i), k_i, T_i, ns, cosmo.delta_H()**2, h, Om, Ob, Tc, out)
    print(label, 'exactP I00', ' '.join('%.0e'%e for e in np.abs(out/ref['I00']-1)))
run_exact('default'); run_exact('ln_mid .5', ln_mid=0.5); run_exact('dq .03 ln_mid .5 n10', bao_dq=0.03, ln_mid=0.5, no=10, ni=10)
print('---- more')
run_exact('ln_mid .5 ln_far 1', ln_mid=0.5, ln_far=1.0)
run_exact('ln_mid .5 ni 12', ln_mid=0.5, ni=12)
run_exact('ln_mid .5 no 12', ln_mid=0.5, no=12)
run_exact('ln_mid .5 bao_hi 1.5', ln_mid=0.5, bao_hi=1.5)
run_exact('ln_mid .5 qmin1e-7 ', ln_mid=0.5, qmin=1e-7)
