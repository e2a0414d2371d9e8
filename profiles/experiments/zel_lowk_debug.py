"""debug: Zel'dovich P00/P01/P11 at low k, GPU vs the reference's shipped tables and vs the oracle on the same state"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyrsd_b200 import pygcl
from tests.gpu_common import golden_pygcl_cosmology
from oracle import gcl_ref as R
g = np.load(os.path.join(ROOT, 'tests/golden/golden_pt.npz'))
cosmo = golden_pygcl_cosmology(g)
state = (float(g['zel_sigma8_z']), float(g['zel_k0_low']), float(g['zel_sigmasq']), g['zel_X0'], g['zel_XX'], g['zel_YY'])
print('state sigma8_z', state[0], 'k0_low', state[1])
k = g['hzpt_k']
rc = R.golden_cosmology(g)
for lowk in (False, True):
    zr = R.RefZeldovich(rc, lowk, state)
    for name, cls in (('P00', pygcl.ZeldovichP00), ('P01', pygcl.ZeldovichP01), ('P11', pygcl.ZeldovichP11)):
        z = cls(pygcl.ZeldovichPS(cosmo, lowk, *state))
        mine = z(k)
        ref = zr(name, k, int(lowk), state[1])
        err = np.abs(mine/ref - 1)
        sel = k >= 1e-3
        i = np.argmax(np.where(sel, err, 0))
        print('lowk', lowk, name, 'max rel err (k >= 1e-3) %.2e at k = %.4g' % (err[i], k[i]), ' at k in 4e-3..8e-3:',
              np.array2string(err[(k > 4e-3) & (k < 8e-3)], precision=1))
# shipped table, last sigma8
z = pygcl.ZeldovichP00(pygcl.ZeldovichPS(cosmo, False, *state))
z.SetSigma8AtZ(float(g['hzpt_sigma8'][99]))
err = np.abs(z(k)/g['hzpt_P00'][-1] - 1)
print('vs shipped table (no lowk): worst %.2e at k = %.4g; k < 0.02: %.2e' % (err.max(), k[err.argmax()], err[k < 0.02].max()))
z = pygcl.ZeldovichP00(pygcl.ZeldovichPS(cosmo, True, *state))
z.SetSigma8AtZ(float(g['hzpt_sigma8'][99]))
err = np.abs(z(k)/g['hzpt_P00'][-1] - 1)
print('vs shipped table (lowk): worst %.2e at k = %.4g; k < 0.02: %.2e' % (err.max(), k[err.argmax()], err[k < 0.02].max()))
