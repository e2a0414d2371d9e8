import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from pyrsd_b200 import pygcl, rsd
from tests.gpu_common import golden_pygcl_cosmology, relerr_floor
golden = np.load('/root/repo/tests/golden/golden_pt.npz'); tight = np.load('/root/repo/tests/golden/oracle_tight.npz')
cosmo = golden_pygcl_cosmology(golden); plin = pygcl.LinearPS(cosmo, 0.)
eng = rsd.PTEngine()
sp = eng.spectra(cosmo.GetDiscreteK(), cosmo.GetDiscreteTk(), cosmo.n_s(), cosmo.delta_H()**2, cosmo.h(), cosmo.Omega0_m(), cosmo.Omega0_b(), cosmo.Tcmb())
res = eng.run(sp)
ML = eng.get(res, 'ML')[0]; k = eng.k_interp; sel = np.array([60, 130, 170, 195, 215, 235, 249]); Pk = plin(k)
olp = eng.get(res, 'oneloop')[0]
ol = {w: cls(plin) for w, cls in (('dd', pygcl.OneLoopPdd), ('dv', pygcl.OneLoopPdv), ('vv', pygcl.OneLoopPvv))}
for j, w in enumerate(('dd', 'dv', 'vv')):
    print('one-loop table', w, 'plan vs one-off max rel', np.max(np.abs(olp[j] - ol[w].GetOneLoopPower())/np.maximum(np.abs(olp[j]), 1e-30)))
spec = [('vv', 'dd', 0, 1), ('vv', 'dd', 0, 2), ('dv', None, 0, 3), ('dv', None, 0, 4), ('vv', None, 2, 3), ('vv', None, 3, 2), ('vv', None, 3, 3)]
for j, (p1, p2, m, n) in enumerate(spec):
    I = pygcl.ImnOneLoop(ol[p1]) if p2 is None else pygcl.ImnOneLoop(ol[p1], ol[p2], 1e-4)
    for w, fn in enumerate((I.EvaluateLinear, I.EvaluateCross, I.EvaluateOneLoop)):
        one = fn(k[sel], m, n)
        e_po = relerr_floor(ML[j, w][sel], one, Pk[sel]).max()
        e_pt = relerr_floor(ML[j, w][sel], tight['ML'][j, w], Pk[sel]).max()
        e_ot = relerr_floor(one, tight['ML'][j, w], Pk[sel]).max()
        print('ML%d part%d: plan-vs-oneoff %.1e  plan-vs-tight %.1e  oneoff-vs-tight %.1e' % (j, w, e_po, e_pt, e_ot))
