import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
E.L.emul_set_tier.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
KSPEC = list(range(16, 29))
spec = [('vv','dd',29),('vv','dd',30),('dv','dv',25),('dv','dv',26),('vv','vv',11),('vv','vv',14),('vv','vv',15)]
kint = np.logspace(np.log10(5e-6), 0, 250)
def allk(kk):
    res = []
    for kid in range(16): res.append(E.ik(kid, kk, tabL, tabL, 0)[0])
    for kid in KSPEC: res.append(E.ik(kid, kk, tabL, tabL, 1)[0])
    return np.array(res)
def mlk(kk):
    res = []
    for j,(p1,p2,kid) in enumerate(spec):
        res.append(E.ik(kid, kk, tabs['L'], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs['L'], tabs[p2], 2, qmax=10.)[0] + E.ik(kid, kk, tabs[p1], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs[p1], tabs[p2], 2, qmax=10.)[0])
    return np.array(res)
def ref_cfg():
    E.L.emul_set_config_default()
    for i in range(4): E.L.emul_set_tier(i, [0.02,0.08,0.3,1e30][i], 14, 10, 10)
tier = int(sys.argv[1]); lo, hi = [(5e-6,0.02),(0.02,0.08),(0.08,0.3),(0.3,10.)][tier]
sel = (kint >= lo) & (kint < hi); kk = np.ascontiguousarray(kint[sel][::max(1, sel.sum()//12)])
k1s = np.ascontiguousarray(k1[(k1 >= max(lo,1e-5)) & (k1 < hi)][::max(1, ((k1>=lo)&(k1<hi)).sum()//12)])
fl = PL(kk); fl1 = PL(k1s)
ref_cfg(); rI = allk(kk); r1 = allk(k1s)[[0,1,5,11,14,15]]
E.L.emul_use_ml_variant()
for i in range(4): E.L.emul_set_tier(i, [0.02,0.06,0.3,1e30][i], 14, 10, 10)
rM = mlk(kk) if hi <= 1.0 else None
cands = eval(sys.argv[2])
for (no, ni, nt) in cands:
    E.L.emul_set_config_default(); E.L.emul_set_tier(tier, [0.02,0.08,0.3,1e30][tier], no, ni, nt)
    oI = allk(kk); n = E.ik(0, kk, tabL, tabL, 0)[1]/len(kk)
    o1 = allk(k1s)[[0,1,5,11,14,15]]
    eI = (np.abs(oI-rI)/np.maximum(np.abs(rI), fl[None,:])).max(); e1 = (np.abs(o1-r1)/np.maximum(np.abs(r1), fl1[None,:])).max()
    # 1-loop combination amplifies I errors: 2I vs result ~ (2I+6k^2JP); report error of 2*I00 relative to max(|P1loop|, P_L) roughly = 2*e
    msg = '(%d,%d,%d) nodes/k %5d  IK %.1e  1loopI %.1e' % (no, ni, nt, n, eI, e1)
    if rM is not None:
        E.L.emul_use_ml_variant(); E.L.emul_set_tier(tier, [0.02,0.06,0.3,1e30][tier], no, ni, nt)
        oM = mlk(kk); nm = E.ik(29, kk, tabs['L'], tabs['L'], 2, qmax=10.)[1]/len(kk)
        msg += '  ML %.1e (nodes %d)' % ((np.abs(oM-rM)/np.maximum(np.abs(rM), fl[None,:])).max(), nm)
    print(msg, flush=True)
