import sys
sys.path.insert(0,'/root/repo/scratch')
from acc_all import *
def setcfg(**kw):
    cfg = dict(qmin=1e-8, n_outer=10, n_inner=7, ln_far=2.0, ln_mid=0.5, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05,
               n_inner_thin=5, ln_thin=0.2, ln_inner=0.5, mid_lo=1e-3, mid_hi=10.0, n_outer_hi=12, k_switch=0.3); cfg.update(kw)
    E.L.emul_set_config(cfg['qmin'], cfg['n_outer'], cfg['n_inner'], cfg['ln_far'], cfg['ln_mid'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq'])
    E.L.emul_set_config_ext(cfg['n_inner_thin'], cfg['ln_thin'], cfg['ln_inner'], cfg['mid_lo'], cfg['mid_hi'])
    E.L.emul_set_config_hi.argtypes = [C.c_int, C.c_double]
    E.L.emul_set_config_hi(cfg['n_outer_hi'], cfg['k_switch'])
kint = np.logspace(np.log10(5e-6), 0, 250)
klo = np.ascontiguousarray(kint[120:215:4]); Plo = PL(klo)
k1lo = np.ascontiguousarray(k1[380:680:12]); P1lo = PL(k1lo)
def allk(kk, mode_ml=False):
    res = []
    for kid in list(range(16)):
        res.append(E.ik(kid, kk, tabL, tabL, 0)[0])
    for kid in KSPEC:
        res.append(E.ik(kid, kk, tabL, tabL, 1)[0])
    return np.array(res)
def mlk(kk):
    res = []
    for j,(p1,p2,kid) in enumerate(spec):
        res.append(E.ik(kid, kk, tabs['L'], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs['L'], tabs[p2], 2, qmax=10.)[0] + E.ik(kid, kk, tabs[p1], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs[p1], tabs[p2], 2, qmax=10.)[0])
    return np.array(res)
setcfg(); ref = allk(klo); ref1 = allk(k1lo); n_ref = E.ik(0, klo, tabL, tabL, 0)[1]/len(klo)
setcfg(n_outer=12, n_outer_hi=12, n_inner=8, n_inner_thin=8, ln_thin=0.); refml = mlk(klo); n_refml = E.ik(0, klo, tabL, tabL, 2, qmax=10.)[1]/len(klo)
for label, kw in (('no6 ni4 t3', dict(n_outer=6, n_inner=4, n_inner_thin=3)),
                  ('no6 ni4 t3 lnmid.7', dict(n_outer=6, n_inner=4, n_inner_thin=3, ln_mid=0.7)),
                  ('no5 ni4 t3 lnmid1 far3', dict(n_outer=5, n_inner=4, n_inner_thin=3, ln_mid=1.0, ln_far=3.0)),
                  ('no8 ni5 t4', dict(n_outer=8, n_inner=5, n_inner_thin=4)),
                  ('no6 ni4 t3 dq.1', dict(n_outer=6, n_inner=4, n_inner_thin=3, bao_dq=0.1)),
                  ('no4 ni3 t2 lnmid1 far3 dq.1', dict(n_outer=4, n_inner=3, n_inner_thin=2, ln_mid=1.0, ln_far=3.0, bao_dq=0.1)),
                  ):
    setcfg(**kw)
    out = allk(klo); n = E.ik(0, klo, tabL, tabL, 0)[1]/len(klo)
    e = np.abs(out - ref)/np.maximum(np.abs(ref), Plo[None,:])
    o1 = allk(k1lo); e1 = np.abs(o1 - ref1)/np.maximum(np.abs(ref1), P1lo[None,:])
    # 1-loop combination 2I + 6k^2 J P: only the I error matters
    kw2 = dict(kw); kw2.update(ln_thin=0.2 if 'ln_thin' not in kw else kw['ln_thin'])
    oml = mlk(klo); nml = E.ik(0, klo, tabL, tabL, 2, qmax=10.)[1]/len(klo)
    eml = np.abs(oml - refml)/np.maximum(np.abs(refml), Plo[None,:])
    print('%-30s nodes %5d (ref %d) ml %d (ref %d) | err/floor by k:'%(label, n, n_ref, nml, n_refml))
    print('     IK   ', ' '.join('%.0e'%x for x in e.max(axis=0)[::3]))
    print('     1loop', ' '.join('%.0e'%x for x in e1.max(axis=0)[::3]))
    print('     ML   ', ' '.join('%.0e'%x for x in eml.max(axis=0)[::3]))
print('k      ', ' '.join('%.0e'%x for x in klo[::3]))
print('k1     ', ' '.join('%.0e'%x for x in k1lo[::3]))
