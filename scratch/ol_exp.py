import sys
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
P1 = PL(k1)
def runol(label, sel=slice(0,1000,1), **kw):
    cfg = dict(qmin=1e-8, n_outer=12, n_inner=8, ln_far=2.0, ln_mid=0.5, bao_lo=0.01, bao_hi=0.7, bao_dq=0.05); cfg.update(kw)
    E.L.emul_set_config(cfg['qmin'], cfg['n_outer'], cfg['n_inner'], cfg['ln_far'], cfg['ln_mid'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq'])
    kk = np.ascontiguousarray(k1[sel])
    res = {}
    for j,(kid,jid) in enumerate(((0,0),(1,1),(5,4))):
        I, nn = E.ik(kid, kk, tabL, tabL, 0, qmax=1e5)
        J = E.linear(0, jid, kk, 1e-5, 1e5, tabL)
        tab = 2*I + 6*kk**2*J*P1[sel]
        e = np.abs(tab - com[j][sel])/np.maximum(np.abs(com[j][sel]), P1[sel])
        lo = kk < 0.5
        i = np.argmax(np.where(lo, e, 0))
        print(label, 'P%d'%j, 'max(k<0.5) %.1e at k=%.4g'%(e[lo].max(), kk[i]), 'max all %.1e'%e.max(), 'nodes/k', nn//len(kk))
        res[j] = e
    return res
if __name__ == '__main__':
    r = runol('default')
    e = r[1]; idx = np.where((e > 5e-6) & (k1 < 0.5))[0]
    print(idx, k1[idx], e[idx])
