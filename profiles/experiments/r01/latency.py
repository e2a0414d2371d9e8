"""latency of one full evaluation as a function of the batch size (device-resident inputs, host-visible sync)"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np, ctypes as C
import bench
from pyrsd_b200 import _native as N, rsd
L = rsd._L()
L.prsd_ctx_sync.argtypes = [C.c_void_p]
eng = rsd.PTEngine()
for B in (1, 2, 8, 32, 128, 512):
    data = bench.synthetic_batch(B, 0)
    gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
    res = rsd._new_result()
    npts = len(data['kk']); out = np.empty((B, npts))
    sp = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
    ts = []
    for it in range(6):
        t0 = time.perf_counter()
        sp2 = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
        eng.run(sp2, result=res)
        poles = eng.galaxy_poles(gal, sp2, res, data['wpar'], data['stoch'], np.logspace(-2.95, np.log10(0.49), 1000), [0, 2, 4], 40)
        t1 = time.perf_counter()
        P = eng.galaxy_power(gal, sp2, res, data['wpar'], data['stoch'], data['kk'], data['mm'])
        t2 = time.perf_counter()
        sp2.close()
        ts.append((t1 - t0, t2 - t1))
    a = np.array(ts[2:]).mean(axis=0)
    print('batch %4d: tables + multipoles (host in, host out) %.3f ms = %.1f us per cosmology; extra P(k,mu) 1000x40 to host %.3f ms' % (B, 1e3*a[0], 1e6*a[0]/B, 1e3*a[1]), flush=True)
