"""time the pieces of the host-buffer (e2e) path on the GPU box"""
import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, ctypes as C
import bench
from pyrsd_b200 import _native as N, pygcl, rsd
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
data = bench.synthetic_batch(B, 0)
L = rsd._L()
eng = rsd.PTEngine()
gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
res = rsd._new_result()
npts = len(data['kk']); out = np.empty((B, npts))
def sync(): N.check(L.prsd_ctx_sync(eng.ctx.ptr))
L.prsd_ctx_sync.argtypes = [C.c_void_p]
for it in range(6):
    t0 = time.perf_counter()
    sp = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
    sync(); t1 = time.perf_counter()
    eng.run(sp, result=res); sync(); t2 = time.perf_counter()
    N.check(L.prsd_galaxy_power(gal.ptr, sp.ptr, res.ptr, B, None, data['wpar'], data['stoch'], npts, data['kk'], data['mm'], out))
    t3 = time.perf_counter()
    sp.close(); t4 = time.perf_counter()
    print('iter %d: spectra %.2f ms  run %.2f ms  galaxy_power(host) %.2f ms  close %.2f ms' % (it, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)), flush=True)
