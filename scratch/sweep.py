import sys, time
sys.path.insert(0, '/root/repo/scratch'); sys.path.insert(0,'/root/repo')
import numpy as np
from proto_quad import *
ref = np.load('scratch/ref_tight.npy', allow_pickle=True).item()
ks = np.array([1e-4, 1e-3, 5e-3, 0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 0.8, 1.0, 3.0, 10.0])
Pk = PL(ks)
tests = [('I00', f00, True, True), ('I13', f13, False, True), ('K00', k00, True, False), ('K11', k11, False, False), ('K01', k01, True, False)]
def run(cfg, label):
    worst = {}
    ns=None
    for name, kern, sym, isI in tests:
        errs=[]; ns=[]
        for i,k in enumerate(ks):
            val, n = Imn_fixed(k, kern, sym, cfg, V8=isI)
            errs.append(abs(val-ref[name][i])/max(abs(ref[name][i]), Pk[i])); ns.append(n if sym else n//2)
        worst[name]=errs
    print(label)
    for name in worst: print('  %-6s'%name, ' '.join('%.0e'%e for e in worst[name]))
    print('  nodes(half)', ns, 'mean', int(np.mean(ns)))
base = dict(qmin=1e-8, n_outer=8, n_inner=8, ln_outer=1.5, ln_inner=0.7, bao_lo=0.01, bao_hi=0.7, bao_dq=0.03, bao_dq_in=0.03)
for label, upd in [
    ('dq .06 n8', dict(bao_dq=0.06, bao_dq_in=0.06)),
    ('dq .06 n10', dict(bao_dq=0.06, bao_dq_in=0.06, n_outer=10, n_inner=10)),
    ('dq .05 n8 lnout2 lnin1', dict(bao_dq=0.05, bao_dq_in=0.05, ln_outer=2.0, ln_inner=1.0)),
    ('dq .04 n8 bao .015-.5', dict(bao_dq=0.04, bao_dq_in=0.04, bao_lo=0.015, bao_hi=0.5)),
    ('qmin 1e-7', dict(qmin=1e-7)),
]:
    cfg = dict(base); cfg.update(upd)
    run(cfg, label)
