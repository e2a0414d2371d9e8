"""Prototype (numpy) of the fixed-order (ln a, v') quadrature for I_mn/K_mn vs the tight oracle."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from numpy.polynomial.legendre import leggauss
from oracle import gcl_ref as R

g = np.load('tests/golden/golden_pt.npz')
cosmo = R.golden_cosmology(g)
PL = R.RefLinearPS(cosmo)

# ---- kernels (u, v)
def f00(u,v): return 4*(4+3*v*v+u*u*(3-10*v*v))**2/(49*(u*u-v*v)**4)
def f22(u,v): return (4*(-1+u*v)*(-8+v*v+u*u*(1+6*v*v)))/(7*(u-v)**4*(u+v)**2)
def f13(u,v):
    u2,v2=u*u,v*v
    return (8*(v2+u2*(1-2*v2)+u**3*v*(-2+3*v2)+u*(v-2*v**3)))/((u-v)**4*(u+v)**2)
def f03(u,v): return (8*(-1+u*u)*(-1+3*u*v)*(-1+v*v))/((u-v)**4*(u+v)**2)
def f33(u,v): return (16*(3+2*v*v+3*v**4+u*u*(2+12*v*v-30*v**4)+u**4*(3-30*v*v+35*v**4)))/(u*u-v*v)**4
def F2(u,v): return (8+6*v*v+u*u*(6-20*v*v))/(7*(u*u-v*v)**2)
def S2(u,v):
    u2,v2=u*u,v*v
    return (2*(6+u2*u2-6*v2+v2*v2+u2*(-6+4*v2)))/(3*(u2-v2)**2)
def k00(u,v): return F2(u,v)
def k11(u,v): return (2-2*u*v)/(u-v)**2
def h04(u,v): return 2*(1+v*v+u*u*(1-3*v*v))/(u*u-v*v)**2
def k20s_b(u,v): return S2(u,v)*h04(u,v)
def k01(u,v): return np.ones_like(u)
def k00s(u,v): return F2(u,v)*S2(u,v)

QMIN_TAB = 1e-8
def make_edges(lo, hi, extra, max_ln, bao_lo, bao_hi, bao_dq):
    pts = [lo, hi] + [e for e in extra if lo < e < hi] + [e for e in (bao_lo, bao_hi) if lo < e < hi]
    pts = sorted(set(pts))
    out = [pts[0]]
    for p0, p1 in zip(pts[:-1], pts[1:]):
        mid = np.sqrt(p0*p1)
        if bao_lo <= mid <= bao_hi:
            n = max(1, int(np.ceil((p1-p0)/bao_dq - 1e-9)))
            seg = p0 + (p1-p0)*np.arange(1, n+1)/n
        else:
            n = max(1, int(np.ceil(np.log(p1/p0)/max_ln - 1e-9)))
            seg = p0*np.exp(np.log(p1/p0)*np.arange(1, n+1)/n)
        seg[-1] = p1
        out.extend(seg.tolist())
    return np.array(out)

def nodes_half(k, cfg, qmax=1e5):
    """nodes for the half r>=q: returns a, b, u, vprime, weight (weight includes (2/k) a^2 b jacobian... no: plain dx dv')"""
    xo, wo = leggauss(cfg['n_outer'])
    xi, wi = leggauss(cfg['n_inner'])
    A=[];B=[];U=[];Vp=[];W=[]
    amax = qmax
    edges = make_edges(cfg['qmin'], amax, [k/2, k, 2*k, 1e-5, 1.0457e-5, 10., 20.678], cfg['ln_outer'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq'])
    for e0, e1 in zip(edges[:-1], edges[1:]):
        l0, l1 = np.log(e0), np.log(e1)
        xs = 0.5*(l1+l0) + 0.5*(l1-l0)*xo
        ws = 0.5*(l1-l0)*wo
        for x, w in zip(xs, ws):
            a = np.exp(x)
            if a < k/2:
                blo, bhi = k - a, k + a
            else:
                blo, bhi = a, a + k
            be = make_edges(blo, bhi, [1e-5, 1.0457e-5, 10., 20.678], cfg['ln_inner'], cfg['bao_lo'], cfg['bao_hi'], cfg['bao_dq_in'])
            for b0, b1 in zip(be[:-1], be[1:]):
                bs = 0.5*(b1+b0) + 0.5*(b1-b0)*xi
                wb = 0.5*(b1-b0)*wi          # db
                vp = (bs - a)/k             # v' = (b-a)/k ; dv' = db/k
                A.append(np.full_like(bs, a)); B.append(bs); Vp.append(vp)
                U.append(2*a/k + vp); W.append(w*wb/k)
    return map(np.concatenate, (A,B,U,Vp,W))

def integrate(k, kern, sign, cfg, Pfun):
    a,b,u,vp,w = nodes_half(k, cfg)
    mask = u <= 2*1e5/k
    Pa = Pfun(a); Pb = Pfun(b)
    val = np.sum(w*mask*(2/k)*a*a*b*Pa*Pb*kern(u, sign*vp))
    return val, len(a)

def Imn_fixed(k, kern, symmetric, cfg, V8=True):
    V = k/(8*np.pi**2) if V8 else k/(4*np.pi**2)
    pos, n1 = integrate(k, kern, +1, cfg, PL)
    if V8:
        if symmetric: return 2*V*pos, n1
        neg, n2 = integrate(k, kern, -1, cfg, PL)
        return V*(pos+neg), n1+n2
    else:
        if symmetric: return V*pos, n1
        neg, n2 = integrate(k, kern, -1, cfg, PL)
        return 0.5*V*(pos+neg), n1+n2

if __name__ == '__main__':
    cfg = dict(qmin=1e-8, n_outer=8, n_inner=8, ln_outer=1.5, ln_inner=0.7, bao_lo=0.01, bao_hi=0.7, bao_dq=0.03, bao_dq_in=0.03)
    ks = np.array([1e-4, 1e-3, 5e-3, 0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 0.8, 1.0, 3.0, 10.0])
    tests = [('I00', f00, True, True, (0,0)), ('I22', f22, False, True, (2,2)), ('I13', f13, False, True, (1,3)), ('I33', f33, True, True,(3,3)),
             ('K00', k00, True, False, (0,0,False,0)), ('K11', k11, False, False, (1,1,False,0)), ('K20s_b', k20s_b, True, False, (2,0,True,1)), ('K01', k01, True, False,(0,1,False,0))]
    t=time.time()
    ref = {}
    for name, kern, sym, isI, idx in tests:
        if isI: ref[name] = R.Imn(PL, ks, idx[0], idx[1], epsrel=1e-9)
        else: ref[name] = R.Kmn(PL, ks, idx[0], idx[1], idx[2], idx[3], epsrel=1e-9)
    print('oracle time', time.time()-t)
    Pk = PL(ks)
    np.save('scratch/ref_tight.npy', ref, allow_pickle=True)
    for name, kern, sym, isI, idx in tests:
        errs=[]; ns=[]
        for i,k in enumerate(ks):
            val, n = Imn_fixed(k, kern, sym, cfg, V8=isI)
            errs.append(abs(val-ref[name][i])/max(abs(ref[name][i]), Pk[i])); ns.append(n)
        print(name, ' '.join('%.1e'%e for e in errs))
    print('nodes', ns)
