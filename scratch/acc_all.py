"""Accuracy of a quadrature configuration in CPU emulation vs the tight-eps reference fixture: I (16), K (13) at 19 k,
ML (7x3) at 7 k, one-loop tables (sampled)."""
import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
E.L.emul_set_config_ext.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
kt = t['k']; Pk = t['plin_k']
KSPEC = list(range(16, 29))
spec = [('vv','dd',29),('vv','dd',30),('dv','dv',25),('dv','dv',26),('vv','vv',11),('vv','vv',14),('vv','vv',15)]
P1 = PL(k1)
def run(label, full=True, **kw):
    cfg = dict(qmin=1e-8, n_outer=12, n_inner=8, ln_far=2.0, ln_mid=0.5, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05,
               n_inner_thin=8, ln_thin=0.0, ln_inner=0.5, mid_lo=1e-3, mid_hi=10.0, n_outer_hi=None, k_switch=0.3); cfg.update(kw)
    E.L.emul_set_config(cfg['qmin'], cfg['n_outer'], cfg['n_inner'], cfg['ln_far'], cfg['ln_mid'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq'])
    E.L.emul_set_config_ext(cfg['n_inner_thin'], cfg['ln_thin'], cfg['ln_inner'], cfg['mid_lo'], cfg['mid_hi'])
    E.L.emul_set_config_hi.argtypes = [C.c_int, C.c_double]
    if cfg['n_outer_hi']: E.L.emul_set_config_hi(cfg['n_outer_hi'], cfg['k_switch'])
    wI = 0; wK = 0
    for kid in range(16):
        out, nI = E.ik(kid, kt, tabL, tabL, 0)
        wI = max(wI, (np.abs(out - t['I'][kid])/np.maximum(np.abs(t['I'][kid]), Pk)).max())
    for j, kid in enumerate(KSPEC):
        out, _ = E.ik(kid, kt, tabL, tabL, 1)
        wK = max(wK, (np.abs(out - t['K'][j])/np.maximum(np.abs(t['K'][j]), Pk)).max())
    wM = 0
    for j,(p1,p2,kid) in enumerate(spec):
        lin, nM = E.ik(kid, kml, tabs['L'], tabs['L'], 2, qmax=10.)
        c1, _ = E.ik(kid, kml, tabs['L'], tabs[p2], 2, qmax=10.)
        c2, _ = E.ik(kid, kml, tabs[p1], tabs['L'], 2, qmax=10.)
        ol, _ = E.ik(kid, kml, tabs[p1], tabs[p2], 2, qmax=10.)
        if j == 5:
            lin, _ = E.ik(11, kml, tabs['L'], tabs['L'], 2, qmax=10.); ol, _ = E.ik(11, kml, tabs[p1], tabs[p2], 2, qmax=10.)
        for w, v in enumerate((lin, c1+c2, ol)):
            wM = max(wM, (np.abs(v-t['ML'][j,w])/np.maximum(np.abs(t['ML'][j,w]), floor)).max())
    msg = '%-28s I %.1e K %.1e ML %.1e nodes/k I %d ML %d'%(label, wI, wK, wM, nI//len(kt), nM//len(kml))
    if full:
        sel = slice(0, 1000, 7); kk = np.ascontiguousarray(k1[sel]); w1 = 0; w1lo = 0
        for j,(kid,jid) in enumerate(((0,0),(1,1),(5,4))):
            I, n1 = E.ik(kid, kk, tabL, tabL, 0)
            J = E.linear(0, jid, kk, 1e-5, 1e5, tabL)
            e = np.abs(2*I + 6*kk**2*J*P1[sel] - com[j][sel])/np.maximum(np.abs(com[j][sel]), P1[sel])
            w1 = max(w1, e.max()); w1lo = max(w1lo, e[kk < 0.5].max())
        I23,_ = E.ik(11, kk, tabL, tabL, 0); I32,_ = E.ik(14, kk, tabL, tabL, 0); I33,_ = E.ik(15, kk, tabL, tabL, 0)
        e = np.abs(I23 + 2./3*I32 + 0.2*I33 - com[3][sel])/np.maximum(np.abs(com[3][sel]), P1[sel])
        msg += ' | 1loop k<.5 %.1e all %.1e P22 %.1e nodes/k %d'%(w1lo, w1, e.max(), n1//len(kk))
    print(msg, flush=True)
if __name__ == '__main__':
    run('default')
    run('thin4 .05', n_inner_thin=4, ln_thin=0.05)
    run('thin4 .1', n_inner_thin=4, ln_thin=0.1)
    run('thin4 .2', n_inner_thin=4, ln_thin=0.2)
    run('thin5 .2', n_inner_thin=5, ln_thin=0.2)
    run('thin6 .3', n_inner_thin=6, ln_thin=0.3)
    run('thin3 .05', n_inner_thin=3, ln_thin=0.05)
