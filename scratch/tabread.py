import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import gcl_ref as R
from tests.host_emul import emul_lib
t = np.load('tests/golden/oracle_tight.npz')
E = emul_lib.load()
x = E.knots(); q = np.exp(x)
k1 = R.logspace(1e-5, 10, 1000)
com = t['oneloop_common']
for j,name in enumerate(['dd','dv','vv','22']):
    y = com[j]
    out = np.zeros(E.M)
    sel = (x >= np.log(1e-5)-1e-12) & (x <= np.log(10.)+1e-12)
    out[sel] = E.spline(k1, y, np.clip(q[sel],1e-5,10.))
    for lo,hi in [(1e-5,1e-3),(1e-3,1e-2),(1e-2,0.1),(0.1,1),(1,3),(3,9.99)]:
        qq = np.exp(np.random.default_rng(1).uniform(np.log(lo),np.log(hi),4000))
        ex = E.spline(k1,y,qq); rd = E.table_read(out,qq)
        print(name, lo, hi, 'max|d|/max|y| %.1e'%(np.abs(rd-ex).max()/np.abs(ex).max()), 'max rel %.1e'%np.max(np.abs(rd-ex)/np.abs(ex)))
