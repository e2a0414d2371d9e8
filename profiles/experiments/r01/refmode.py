"""reference-mode ensemble step: ONE tabulated cosmology, 2048 walkers varying sigma8_z, f, biases, FOG, AP (what
rsdfit does during a fit): time of galaxy_power / galaxy_poles per step"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np
import bench
from pyrsd_b200 import rsd
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
data = bench.synthetic_batch(nw, 0)
eng = rsd.PTEngine()
sp = eng.spectra(data['k'][:1], data['T'][:1], data['pars'][0, 0], data['pars'][0, 1], data['pars'][0, 2], data['pars'][0, 3], data['pars'][0, 4], data['pars'][0, 5])
t0 = time.perf_counter(); res = eng.run(sp); print('PT tables of the cosmology: %.2f ms' % (1e3*(time.perf_counter() - t0)))
gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
cmap = np.zeros(nw, dtype=np.int32)
kp = np.logspace(-2, np.log10(0.4), 40)
for it in range(5):
    t0 = time.perf_counter()
    poles = eng.galaxy_poles(gal, sp, res, data['wpar'], data['stoch'], kp, [0, 2, 4], 40, cmap=cmap)
    t1 = time.perf_counter()
    print('step %d: multipoles (3 x 40 k) of %d walkers %.2f ms = %.2f us per walker' % (it, nw, 1e3*(t1 - t0), 1e6*(t1 - t0)/nw), flush=True)
