"""soak: 300 full steps (PT + galaxy, host-buffer path and device path alternating) -- free memory must not drift"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from pyrsd_b200 import rsd
B = 128
data = bench.synthetic_batch(B, 0)
eng = rsd.PTEngine()
gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
free0 = None
ref = None
for it in range(300):
    sp = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
    res = eng.run(sp)
    P = eng.galaxy_power(gal, sp, res, data['wpar'], data['stoch'], data['kk'][:4000], data['mm'][:4000])
    if ref is None:
        ref = P.copy()
    assert np.array_equal(P, ref), it               # bit-reproducible from step to step
    sp.close(); res.close()
    if it in (20, 299):
        torch.cuda.synchronize()
        free, total = torch.cuda.mem_get_info()
        print('step %d: free %.1f MB of %.0f MB' % (it, free/1e6, total/1e6), flush=True)
        if free0 is None:
            free0 = free
assert abs(free - free0) < 64e6, (free0, free)
print('soak ok: bit-identical results over 300 steps, free-memory drift %.1f MB' % ((free0 - free)/1e6))
