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
L.emul_linear.argtypes=[C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
M = L.emul_knot_count()
k_i, T_i, s8, ns, h, Om, Ob, Tc = cosmo.args
tab = np.empty(M); L.emul_plin_table(len(k_i), k_i, T_i, ns, cosmo.delta_H()**2, h, Om, Ob, Tc, tab)
def lin(type, sub, param, lo, hi, scale=1.0, table=tab):
    param=np.ascontiguousarray(np.atleast_1d(np.asarray(param,float))); n=len(param)
    lo=np.broadcast_to(np.asarray(lo,float),(n,)).copy(); hi=np.broadcast_to(np.asarray(hi,float),(n,)).copy(); sc=np.broadcast_to(np.asarray(scale,float),(n,)).copy()
    out=np.empty(n); L.emul_linear(type, sub, n, param, lo, hi, sc, table, out); return out
ks = np.array([1e-4, 1e-3, 5e-3, 0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 0.8, 1.0, 3.0, 10.0])
for (m,n) in []:
    ref = R.Jmn(PL, ks, m, n, epsrel=1e-10)
    out = lin(0, 3*m+n, ks, 1e-5, 1e5)
    print('J%d%d'%(m,n), ' '.join('%.0e'%e for e in np.abs(out/ref-1)))
print('sigv', lin(1,0,[0.],1e-5,100.)[0], PL.sigv_tol(), PL.VelocityDispersion())
kk = g['k_interp'][40::20]
ref = np.array([PL.sigv_tol(0.5*k) for k in kk]); out = lin(1,0,kk,1e-5,0.5*kk)
print('sigmasq_k', np.max(np.abs(out/ref-1)), 'vs golden default', np.max(np.abs(out/g['pt_sigmasq_k'][40::20]-1)))
rs = np.array([0.01, 0.1, 1., 5., 20., 60., 105., 150., 300., 1000.])
t=time.time(); outx = lin(2,0,rs,1e-5,100.); outy = lin(3,0,rs,1e-5,100.); print('emul X/Y time', time.time()-t)
t=time.time(); refx = PL.X_Zel(rs, 1e-9, 1e-13); refy = PL.Y_Zel(rs, 1e-9, 1e-13); print('oracle tight time', time.time()-t)
print('X abs err', ' '.join('%.0e'%e for e in np.abs(outx-refx))); print('  X', refx)
print('Y abs err', ' '.join('%.0e'%e for e in np.abs(outy-refy))); print('  Y', refy)
# Q3 on P^2 table
ref = PL.Q3_tol(np.array([1.0]))[0]; out = lin(4,0,[0.],1e-5,100., 1/(10*np.pi**2), table=tab*tab)[0]
print('Q3', out/ref-1)
print('Sigma(8)', np.sqrt(lin(5,0,[8.],1e-5,100.)[0]), PL.Sigma_tol([8.])[0])
print('Sigma(1,30)', np.sqrt(lin(5,0,[1.,30.],1e-5,100.)), PL.Sigma_tol([1.,30.]))
