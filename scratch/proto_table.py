import sys
sys.path.insert(0, '/root/repo/scratch'); sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import gcl_ref as R
g = np.load('tests/golden/golden_pt.npz')
cosmo = R.golden_cosmology(g); PL = R.RefLinearPS(cosmo)

def build_knots(zones):
    """zones: list of (lo, hi, h). returns x knots (ln q), zone start index, per-zone (x0, h, n)"""
    xs=[]; meta=[]
    for lo,hi,h in zones:
        n = int(np.ceil(np.log(hi/lo)/h))
        hh = np.log(hi/lo)/n
        start = len(xs)
        xs.extend((np.log(lo)+hh*np.arange(n+1)).tolist())   # each zone has its own endpoints (duplicate at boundary)
        meta.append((np.log(lo), hh, n, start))
    return np.array(xs), meta

def interp(xq, tab, meta, order=6):
    out = np.empty_like(xq)
    for i,x in enumerate(xq):
        for (x0,h,n,start) in meta:
            if x <= x0 + h*n + 1e-15: break
        c = int(np.floor((x-x0)/h)); c = min(max(c,0), n-1)
        j0 = min(max(c - (order//2-1), 0), n+1-order)
        s = (x - x0)/h - j0
        val=0.
        for l in range(order):
            L=1.
            for m in range(order):
                if m!=l: L *= (s-m)/(l-m)
            val += L*tab[start+j0+l]
        out[i]=val
    return out

for hb in [0.004, 0.006, 0.008]:
    zones = [(1e-8,1e-5,0.05),(1e-5,5e-3,0.04),(5e-3,1.0,hb),(1.0,10.,0.03),(10.,100.,0.04),(100.,1.2e5,0.06)]
    xs, meta = build_knots(zones)
    tab = PL(np.exp(xs))
    rng = np.random.default_rng(1)
    xq = rng.uniform(np.log(1e-8), np.log(1.2e5), 20000)
    ex = PL(np.exp(xq)); ap = interp(xq, tab, meta)
    rel = np.abs(ap-ex)/ex
    q = np.exp(xq)
    print('hb',hb,'M',len(xs), 'max rel err by zone:', [ '%.1e'%rel[(q>=lo)&(q<hi)].max() for lo,hi,h in zones])
print('order 4')
for hb, hs in [(0.004, 0.02), (0.003, 0.015), (0.005,0.02)]:
    zones = [(1e-8,1e-5,hs*1.5),(1e-5,5e-3,hs),(5e-3,1.0,hb),(1.0,10.,hs),(10.,100.,hs),(100.,1.2e5,hs*1.5)]
    xs, meta = build_knots(zones)
    tab = PL(np.exp(xs))
    rng = np.random.default_rng(1)
    xq = rng.uniform(np.log(1e-8), np.log(1.2e5), 20000)
    ex = PL(np.exp(xq)); ap = interp(xq, tab, meta, order=4)
    rel = np.abs(ap-ex)/ex
    q = np.exp(xq)
    print('hb',hb,'hs',hs,'M',len(xs), 'max rel err by zone:', [ '%.1e'%rel[(q>=lo)&(q<hi)].max() for lo,hi,h in zones])
