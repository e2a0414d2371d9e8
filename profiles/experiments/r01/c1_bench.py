"""BASELINE configs[0]: pygcl OneLoopP00 + ZeldovichP00 at z = 0.55 on 200 k, single cosmology, through the pygcl drop-in"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np
from pyrsd_b200 import pygcl
from tests.gpu_common import golden_pygcl_cosmology
golden = np.load('/root/repo/tests/golden/golden_pt.npz')
k = np.logspace(np.log10(0.005), np.log10(0.5), 200)
for it in range(4):
    cosmo = golden_pygcl_cosmology(golden)
    t0 = time.perf_counter()
    P = pygcl.LinearPS(cosmo, 0.55)
    Pdd = pygcl.OneLoopPdd(P)
    p00 = Pdd.EvaluateFull(k)
    t1 = time.perf_counter()
    zel = pygcl.ZeldovichP00(cosmo, 0.55, False)
    z00 = zel(k)
    t2 = time.perf_counter()
    print('call %d: LinearPS + OneLoopPdd(P_L) + EvaluateFull(200 k): %.2f ms;  ZeldovichP00(cosmo, z) [X(r), Y(r), sigma^2 set-up] + 200 k: %.2f ms'
          % (it, 1e3*(t1 - t0), 1e3*(t2 - t1)), flush=True)
try:
    from oracle import gcl_ref as R
    rc = R.golden_cosmology(golden); rc.set_sigma8_z(cosmo.Sigma8_z(0.55)); rp = R.RefLinearPS(rc, cosmo.Sigma8_z(0.55))
    t0 = time.perf_counter(); ref = R.RefOneLoop(rp, 'dd', 1e-4); ref.EvaluateFull(k); t1 = time.perf_counter()
    z = R.RefZeldovich(rc, False); z('P00', k); t2 = time.perf_counter()
    print('reference C++ (1 core): OneLoopPdd ctor + eval %.0f ms; ZeldovichP00 ctor + 200 k %.0f ms' % (1e3*(t1 - t0), 1e3*(t2 - t1)))
except Exception as e:
    print('oracle not available:', e)
