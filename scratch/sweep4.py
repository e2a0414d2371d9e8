import sys
sys.path.insert(0,'/root/repo/scratch')
from acc_all import *
def run1(label, lo, hi, **kw):
    cfg = dict(qmin=1e-8, n_outer=12, n_inner=8, ln_far=2.0, ln_mid=0.5, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05,
               n_inner_thin=8, ln_thin=0.0, ln_inner=0.5, mid_lo=1e-3, mid_hi=10.0); cfg.update(kw)
    E.L.emul_set_config(cfg['qmin'], cfg['n_outer'], cfg['n_inner'], cfg['ln_far'], cfg['ln_mid'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq'])
    E.L.emul_set_config_ext(cfg['n_inner_thin'], cfg['ln_thin'], cfg['ln_inner'], cfg['mid_lo'], cfg['mid_hi'])
    idx = np.where((k1 >= lo) & (k1 <= hi))[0][::3]
    kk = np.ascontiguousarray(k1[idx]); w1 = 0
    for j,(kid,jid) in enumerate(((0,0),(1,1),(5,4))):
        I, n1 = E.ik(kid, kk, tabL, tabL, 0)
        J = E.linear(0, jid, kk, 1e-5, 1e5, tabL)
        e = np.abs(2*I + 6*kk**2*J*P1[idx] - com[j][idx])/np.maximum(np.abs(com[j][idx]), P1[idx])
        w1 = max(w1, e.max())
    I23,_ = E.ik(11, kk, tabL, tabL, 0); I32,_ = E.ik(14, kk, tabL, tabL, 0); I33,_ = E.ik(15, kk, tabL, tabL, 0)
    e = np.abs(I23 + 2./3*I32 + 0.2*I33 - com[3][idx])/np.maximum(np.abs(com[3][idx]), P1[idx])
    print('%-30s [%g, %g] 1loop %.1e P22 %.1e nodes/k %d'%(label, lo, hi, w1, e.max(), n1//len(kk)), flush=True)
base = dict(n_inner_thin=5, ln_thin=0.2)
for lo, hi in ((0.5, 10.),):
    run1('default', lo, hi)
    run1('thin5 .2', lo, hi, **base)
    run1('no10', lo, hi, n_outer=10, **base)
    run1('no12 ln_mid .6', lo, hi, ln_mid=0.6, **base)
    run1('no14 ln_mid .6', lo, hi, n_outer=14, ln_mid=0.6, **base)
    run1('no14', lo, hi, n_outer=14, **base)
    run1('no16', lo, hi, n_outer=16, **base)
    run1('ni10', lo, hi, n_inner=10, **base)
    run1('ln_far 1', lo, hi, ln_far=1.0, **base)
    run1('bao_hi 1.5', lo, hi, bao_hi=1.5, **base)
    run1('qmin 1e-7', lo, hi, qmin=1e-7, **base)
