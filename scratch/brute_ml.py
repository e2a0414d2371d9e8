import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import gcl_ref as R
from tests.host_emul import emul_lib
t = np.load('tests/golden/oracle_tight.npz')
E = emul_lib.load()
k1 = R.logspace(1e-5, 10, 1000)
com = t['oneloop_common']
def P(j, q):
    out = np.zeros_like(q); s = (q >= 1e-5) & (q <= 10.)
    out[s] = E.spline(k1, com[j], q[s]); return out
def h01(u,v): return -((2*(-1 + u*u)*(-1 + v*v))/(u - v)**4)
def h02(u,v): return (6 - 8*u*v - 2*v*v + u*u*(-2 + 6*v*v))/(u - v)**4
gx, gw = np.polynomial.legendre.leggauss(12)
def brute(f, j1, j2, k, nq=600, nr=60, qmax=10.):
    # q panels in ln q with breaks at k (kink of |k-q|) and 2qmax-... 
    edges = np.unique(np.concatenate([np.linspace(np.log(1e-5), np.log(10.), nq), [np.log(k), np.log(k/2), np.log(qmax-k/2) if qmax>k/2 else 0.]]))
    tot = 0.0
    for e0, e1 in zip(edges[:-1], edges[1:]):
        lq = 0.5*(e0+e1) + 0.5*(e1-e0)*gx; wq = 0.5*(e1-e0)*gw
        q = np.exp(lq)
        Pq = P(j1, q)
        for qi, wi, pq in zip(q, wq, Pq):
            lo = max(abs(k-qi), 1e-5); hi = min(qi+k, 2*qmax-qi, 10.)
            if hi <= lo: continue
            re = np.linspace(lo, hi, nr)
            mid = 0.5*(re[:-1]+re[1:])[:,None]; hw = 0.5*(re[1:]-re[:-1])[:,None]
            r = (mid + hw*gx[None,:]).ravel(); wr = (hw*gw[None,:]).ravel()
            u = (qi+r)/k; v = (r-qi)/k
            tot += wi*qi* np.sum(wr*(2/k**2)*qi*r*pq*P(j2, r)*f(u,v))
    return k/(8*np.pi**2)*tot
kml = t['k_ml']
for ik in (5, 6):
    k = kml[ik]
    for name, f, jj in (('h01', h01, 0), ('h02', h02, 1)):
        b1 = brute(f, 2, 0, k, 300, 30); b2 = brute(f, 2, 0, k, 600, 60)
        print('k=%.3f %s brute %.10e %.10e  tight %.10e   rel(b2,tight) %.1e  rel(b1,b2) %.1e'%(k, name, b1, b2, t['ML'][jj,2,ik], abs(b2/t['ML'][jj,2,ik]-1), abs(b1/b2-1)))
