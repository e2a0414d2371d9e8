"""timings of BASELINE configs[2] (PT tables of 1024 variants) and configs[4] (2048-walker ensemble step, chi^2 out)"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np, ctypes as C
import bench
from pyrsd_b200 import _native as N, rsd
L = rsd._L(); L.prsd_ctx_sync.argtypes = [C.c_void_p]
eng = rsd.PTEngine()
def sync(): N.check(L.prsd_ctx_sync(eng.ctx.ptr))
for B in (1024, 2048):
    data = bench.synthetic_batch(B, 0)
    res = rsd._new_result()
    sp = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
    gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
    kd = np.linspace(0.02, 0.4, 40); n = 3*len(kd)
    datav = np.ones(n); cinv = np.eye(n)
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        sp2 = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
        eng.run(sp2, result=res); sync()
        t1 = time.perf_counter()
        chi2 = eng.galaxy_chi2(gal, sp2, res, data['wpar'], data['stoch'], kd, [0, 2, 4], datav, cinv, 40)
        t2 = time.perf_counter()
        sp2.close()
        ts.append((t1 - t0, t2 - t1))
    a = np.array(ts[1:]).mean(axis=0)
    print('batch %d: all PT tables (16 I, 6 J, 13 K, 4 one-loop x 1000 k, 21 ImnOneLoop, sigma^2(k), X, Y) from host T(k): %.2f ms = %.1f us per cosmology;'
          ' + multipoles and chi^2 per walker: %.2f ms; total %.1f us per walker' % (B, 1e3*a[0], 1e6*a[0]/B, 1e3*a[1], 1e6*a.sum()/B), flush=True)
