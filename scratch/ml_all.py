import sys
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
spec = [('vv','dd',29),('vv','dd',30),('dv','dv',25),('dv','dv',26),('vv','vv',11),('vv','vv',14),('vv','vv',15)]
def runall(label, **kw):
    cfg = dict(qmin=1e-8, n_outer=12, n_inner=8, ln_far=2.0, ln_mid=0.5, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05); cfg.update(kw)
    E.L.emul_set_config(cfg['qmin'], cfg['n_outer'], cfg['n_inner'], cfg['ln_far'], cfg['ln_mid'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq'])
    print(label)
    worst = 0
    for j,(p1,p2,kid) in enumerate(spec):
        lin, nn = E.ik(kid, kml, tabs['L'], tabs['L'], 2, qmax=10.)
        c1, _ = E.ik(kid, kml, tabs['L'], tabs[p2], 2, qmax=10.)
        c2, _ = E.ik(kid, kml, tabs[p1], tabs['L'], 2, qmax=10.)
        ol, _ = E.ik(kid, kml, tabs[p1], tabs[p2], 2, qmax=10.)
        if j == 5:   # reference quirk: f23 kernel for linear and one-loop of (3,2)
            lin, _ = E.ik(11, kml, tabs['L'], tabs['L'], 2, qmax=10.); ol, _ = E.ik(11, kml, tabs[p1], tabs[p2], 2, qmax=10.)
        for w, v in enumerate((lin, c1+c2, ol)):
            e = np.abs(v-t['ML'][j,w])/np.maximum(np.abs(t['ML'][j,w]), floor); worst = max(worst, e.max())
            print('  ML%d part %d'%(j, w), ' '.join('%.0e'%x for x in e))
    print(' worst %.1e nodes/k %d'%(worst, nn//len(kml)))
runall('default')
