import sys, time, ctypes as C, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0, '.')
import numpy as np
from pyrsd_b200 import _native as N
from oracle import gcl_ref as R
g = np.load('tests/golden/golden_pt.npz')
L = N.lib(); ctx = N.Context(0)
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')
L.prsd_fftlog.argtypes=[C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, C.POINTER(C.c_double)]
L.prsd_zeldovich.argtypes=[C.c_void_p, C.c_int, dp, C.c_void_p, dp, dp, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, dp, dp]
L.prsd_compute_xilm.argtypes=[C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp, C.c_double, C.c_double, dp]
L.prsd_compute_xilm_fftlog.argtypes=[C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, dp, dp]
def section(s): print('=== '+s, flush=True)
try:
    section('fftlog vs oracle (reference wrapper + C restatement)')
    rng = np.random.default_rng(3)
    for (Nn, mu, q) in [(1024, 0.5, 0.0), (1024, 7.5, 0.0), (4096, 0.5, 0.0), (512, 2.5, 0.3), (300, 0.5, 0.0), (64, 1.5, -0.2)]:
        dlnr = np.log(1e7)/Nn
        x = np.exp((np.arange(Nn)-Nn/2)*dlnr)
        a = np.vstack([x**1.5*np.exp(-x**2/2/s**2) for s in (0.5, 2.0, 7.0)])
        out = np.empty_like(a); kr = C.c_double()
        N.check(L.prsd_fftlog(ctx.ptr, 3, Nn, mu, q, dlnr, 1.0, 1, a, out, C.byref(kr)))
        ref = np.vstack([R.fftlog(a[i], dlnr, mu, q, 1.0, 1)[0] for i in range(3)]); krr = R.fftlog(a[0], dlnr, mu, q, 1.0, 1)[1]
        print('N=%d mu=%.1f q=%.1f  max|diff|/max|ref| = %.2e   kr %.15g vs %.15g'%(Nn, mu, q, np.max(np.abs(out-ref))/np.max(np.abs(ref)), kr.value, krr))
except Exception: traceback.print_exc()
try:
    section('Zeldovich P00/P01/P11 vs golden sigma8 x k tables')
    kk = g['hzpt_k']; s8 = g['hzpt_sigma8'][g['hzpt_rows']]
    sel = kk >= 5e-3; ks = np.ascontiguousarray(kk[sel])
    nrow = len(s8)
    X0 = np.tile(g['zel_X0'], (nrow,1)); XX = np.tile(g['zel_XX'], (nrow,1)); YY = np.tile(g['zel_YY'], (nrow,1))
    ss = np.full(nrow, float(g['zel_sigmasq'])); ratio = (s8/float(g['zel_sigma8_z']))**2
    out = np.empty((nrow,3,len(ks)))
    t=time.time()
    N.check(L.prsd_zeldovich(ctx.ptr, nrow, X0, XX.ctypes.data_as(C.c_void_p), YY, ss, ratio.ctypes.data_as(C.c_void_p), 0, 5e-3, None, None, len(ks), ks, out))
    print('time %.3fs (incl. plan build)'%(time.time()-t))
    t=time.time()
    N.check(L.prsd_zeldovich(ctx.ptr, nrow, X0, XX.ctypes.data_as(C.c_void_p), YY, ss, ratio.ctypes.data_as(C.c_void_p), 0, 5e-3, None, None, len(ks), ks, out))
    print('time %.4fs (cached plan), %d rows x %d k'%(time.time()-t, nrow, len(ks)))
    for j,name in enumerate(['P00','P01','P11']):
        gold = g['hzpt_'+name][:, sel]
        rel = np.abs(out[:,j,:]/gold-1)
        print(name, 'max rel vs golden', rel.max(), 'median', np.median(rel))
    # XX=None path (X0 + 2 sigma^2)
    out2 = np.empty_like(out)
    N.check(L.prsd_zeldovich(ctx.ptr, nrow, X0, None, YY, ss, ratio.ctypes.data_as(C.c_void_p), 0, 5e-3, None, None, len(ks), ks, out2))
    print('XX implicit vs explicit', np.max(np.abs(out2/out-1)))
except Exception: traceback.print_exc()
try:
    section('pk_to_xi / ComputeXiLM vs oracle')
    cosmo = R.golden_cosmology(g); PL = R.RefLinearPS(cosmo)
    k = np.logspace(-4, 1.5, 700); pk = np.vstack([PL(k), 0.7*PL(k)])
    r = np.logspace(-1, 2.3, 50)
    for ell in (0, 2, 4):
        out = np.empty((2, len(r)))
        N.check(L.prsd_compute_xilm(ctx.ptr, 2, ell, 2, len(k), k, pk, len(r), r, 0.5, (-1.)**(ell//2), out))
        ref = R.pk_to_xi(ell, k, pk[0], r, 0.5)
        print('pk_to_xi ell=%d N=700 (direct kernel): max rel %.2e'%(ell, np.max(np.abs(out[0]-ref))/np.max(np.abs(ref))))
    k = np.logspace(-4, 1.5, 1024); pk = np.vstack([PL(k)])
    out = np.empty((1, len(r))); N.check(L.prsd_compute_xilm(ctx.ptr, 1, 0, 2, len(k), k, pk, len(r), r, 0.5, 1.0, out))
    ref = R.pk_to_xi(0, k, pk[0], r, 0.5); print('pk_to_xi N=1024 (fft kernel): max rel %.2e'%(np.max(np.abs(out[0]-ref))/np.max(np.abs(ref))))
    rr = np.empty(1024); xi = np.empty((1,1024)); N.check(L.prsd_compute_xilm_fftlog(ctx.ptr, 1, 0, 2, 1024, k, pk, 0.0, 0.0, rr, xi))
    r2, xi2 = R.ComputeXiLM_fftlog(0, 2, k, pk[0]); print('ComputeXiLM_fftlog r diff %.1e xi diff %.2e'%(np.max(np.abs(rr/r2-1)), np.max(np.abs(xi[0]-xi2))/np.max(np.abs(xi2))))
except Exception: traceback.print_exc()
print('launches', ctx.launches())
