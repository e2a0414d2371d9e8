import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
from pyrsd_b200 import rsd
for rank in (0, 1, 2, 3):
    data = bench.synthetic_batch(128, rank)
    eng = rsd.PTEngine()
    gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
    sp = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
    res = eng.run(sp)
    P = eng.galaxy_power(gal, sp, res, data['wpar'], data['stoch'], data['kk'], data['mm'])
    bad = np.where(~np.isfinite(P).all(axis=1))[0]; neg = np.where((P <= 0).any(axis=1))[0]
    print('rank', rank, 'nonfinite walkers', bad, 'nonpositive walkers', neg)
    for w in list(neg[:3]):
        j = np.argmin(P[w]); print('   walker', w, 'min P', P[w].min(), 'at k, mu', data['kk'][j], data['mm'][j], 'sigma8_z f b1', data['wpar'][w][[0, 1, 5]])
    del eng
