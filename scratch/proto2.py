import sys, time
sys.path.insert(0, '/root/repo/scratch'); sys.path.insert(0,'/root/repo')
import numpy as np
import proto_quad as pq
from proto_quad import *

def width(q, cfg):
    w = cfg['ln_far']
    if cfg['mid_lo'] <= q <= cfg['mid_hi']: w = min(w, cfg['ln_mid'])
    if cfg['bao_lo'] <= q <= cfg['bao_hi']: w = min(w, cfg['bao_dq']/q)
    return w

def make_edges2(lo, hi, extra, cfg, scale=1.0):
    brk = sorted(set([lo, hi] + [e for e in extra if lo < e < hi] +
                     [e for e in (cfg['mid_lo'], cfg['mid_hi'], cfg['bao_lo'], cfg['bao_hi']) if lo < e < hi]))
    out = [brk[0]]
    for p0, p1 in zip(brk[:-1], brk[1:]):
        # march in ln q with local width, then rescale to fit exactly
        xs = [np.log(p0)]; x1 = np.log(p1)
        while xs[-1] < x1 - 1e-12:
            q = np.exp(xs[-1])
            xs.append(xs[-1] + scale*width(min(q*1.0000001, p1), cfg))
        n = len(xs) - 1
        # rescale marching positions proportionally
        xs = np.array(xs); xs = xs[0] + (xs - xs[0])*(x1 - xs[0])/(xs[-1] - xs[0])
        seg = np.exp(xs[1:]); seg[-1] = p1
        out.extend(seg.tolist())
    return np.array(out)

ZONES = [1e-5, 1.0457e-5, 10., 20.678]
def nodes_half2(k, cfg, qmax=1e5):
    xo, wo = leggauss(cfg['n_outer']); xi, wi = leggauss(cfg['n_inner'])
    A=[];B=[];U=[];Vp=[];W=[]
    edges = make_edges2(cfg['qmin'], qmax, [k/2, k, 2*k] + ZONES, cfg)
    for e0, e1 in zip(edges[:-1], edges[1:]):
        l0, l1 = np.log(e0), np.log(e1)
        xs = 0.5*(l1+l0) + 0.5*(l1-l0)*xo; ws = 0.5*(l1-l0)*wo
        for x, w in zip(xs, ws):
            a = np.exp(x)
            blo, bhi = (k - a, k + a) if a < k/2 else (a, a + k)
            be = make_edges2(blo, bhi, ZONES, cfg, scale=cfg.get('inner_scale',1.0))
            for b0, b1 in zip(be[:-1], be[1:]):
                bs = 0.5*(b1+b0) + 0.5*(b1-b0)*xi; wb = 0.5*(b1-b0)*wi
                vp = (bs - a)/k
                A.append(np.full_like(bs, a)); B.append(bs); Vp.append(vp); U.append(2*a/k + vp); W.append(w*wb/k)
    return map(np.concatenate, (A,B,U,Vp,W))
pq.nodes_half = nodes_half2

ref = np.load('scratch/ref_tight.npy', allow_pickle=True).item()
ks = np.array([1e-4, 1e-3, 5e-3, 0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 0.8, 1.0, 3.0, 10.0])
Pk = PL(ks)
tests = [('I00', f00, True, True), ('I13', f13, False, True), ('K00', k00, True, False), ('K11', k11, False, False), ('K01', k01, True, False)]
def run(cfg, label):
    worst = {}
    for name, kern, sym, isI in tests:
        errs=[]; ns=[]
        for i,k in enumerate(ks):
            val, n = Imn_fixed(k, kern, sym, cfg, V8=isI)
            errs.append(abs(val-ref[name][i])/max(abs(ref[name][i]), Pk[i])); ns.append(n if sym else n//2)
        worst[name]=errs
    print(label)
    for name in worst: print('  %-6s'%name, ' '.join('%.0e'%e for e in worst[name]))
    print('  nodes(half)', ns, 'mean', int(np.mean(ns)))
base = dict(qmin=1e-8, n_outer=8, n_inner=8, ln_far=1.5, ln_mid=0.7, mid_lo=1e-3, mid_hi=2.0, bao_lo=0.01, bao_hi=0.7, bao_dq=0.03)
if __name__ == '__main__':
    for label, upd in [
        ('base2', {}),
        ('ln_mid .5', dict(ln_mid=0.5)),
        ('dq .045', dict(bao_dq=0.045)),
        ('dq .06 n10', dict(bao_dq=0.06, n_outer=10, n_inner=10)),
    ]:
        cfg = dict(base); cfg.update(upd)
        run(cfg, label)
