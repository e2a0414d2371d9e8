import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import gcl_ref as R
from tests.host_emul import emul_lib
g = np.load('tests/golden/golden_pt.npz'); t = np.load('tests/golden/oracle_tight.npz')
cosmo = R.golden_cosmology(g); PL = R.RefLinearPS(cosmo)
E = emul_lib.load()
k_i, T_i, s8, ns, h, Om, Ob, Tc = cosmo.args
tabL = E.plin_table(k_i, T_i, ns, cosmo.delta_H()**2, h, Om, Ob, Tc)
x = E.knots(); q = np.exp(x)
k1 = R.logspace(1e-5, 10, 1000)
def tab1(y):
    out = np.zeros(E.M)
    sel = (x >= np.log(1e-5)-1e-12) & (x <= np.log(10.)+1e-12)
    i0 = np.where(sel)[0][0]; i1 = np.where(sel)[0][-1]
    sel[i0] = False; sel[i1] = False   # duplicate knots belonging to the outer zones
    qq = np.clip(q[sel], 1e-5, 10.)
    out[sel] = E.spline(k1, y, qq)
    return out
com = t['oneloop_common']
tabs = {'L': tabL, 'dd': tab1(com[0]), 'dv': tab1(com[1]), 'vv': tab1(com[2])}
kml = t['k_ml']; floor = PL(kml)
KID = dict(h01=29, h02=30)
import re
src = open('pyrsd_b200/csrc/pt_math.cuh').read()
def run(label, **kw):
    cfg = dict(qmin=1e-8, n_outer=12, n_inner=8, ln_far=2.0, ln_mid=0.5, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05); cfg.update(kw)
    E.L.emul_set_config(cfg['qmin'], cfg['n_outer'], cfg['n_inner'], cfg['ln_far'], cfg['ln_mid'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq'])
    print(label)
    for name, kid, j in (('h01', int(sys.argv[1]), 0), ('h02', int(sys.argv[1])+1, 1)):
        lin, nn = E.ik(kid, kml, tabs['L'], tabs['L'], 2, qmax=10.)
        c1, _ = E.ik(kid, kml, tabs['L'], tabs['dd'], 2, qmax=10.)
        c2, _ = E.ik(kid, kml, tabs['vv'], tabs['L'], 2, qmax=10.)
        ol, _ = E.ik(kid, kml, tabs['vv'], tabs['dd'], 2, qmax=10.)
        for w, v in enumerate((lin, c1+c2, ol)):
            print('  %s part %d'%(name, w), ' '.join('%.0e'%e for e in np.abs(v-t['ML'][j,w])/np.maximum(np.abs(t['ML'][j,w]), floor)), 'nodes/k', nn//len(kml))
run('default')
run('n16/12', n_outer=16, n_inner=12)
run('ln_mid .25', ln_mid=0.25)
run('ln_mid .25 n16/12', ln_mid=0.25, n_outer=16, n_inner=12)
