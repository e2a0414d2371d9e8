import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
spec = [('vv','dd',29),('vv','dd',30),('dv','dv',25),('dv','dv',26),('vv','vv',11),('vv','vv',14),('vv','vv',15)]
E.L.emul_set_tier.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
kint = np.logspace(np.log10(5e-6), 0, 250)
# denser set of test k: need reference values -> use tight fixture k_ml only (7 k) plus self-convergence at other k vs full-order emulation
ktest = np.ascontiguousarray(kint[150:250:7])
def mlk(kk):
    res = []
    for j,(p1,p2,kid) in enumerate(spec):
        res.append(E.ik(kid, kk, tabs['L'], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs['L'], tabs[p2], 2, qmax=10.)[0] + E.ik(kid, kk, tabs[p1], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs[p1], tabs[p2], 2, qmax=10.)[0])
    return np.array(res)
def setml(tiers):
    E.L.emul_set_config_default(); E.L.emul_use_ml_variant()
    for i, t_ in enumerate(tiers): E.L.emul_set_tier(i, *t_)
ref_t = [(1e-30, 16, 12, 12)]*3 + [(1e30, 16, 12, 12)]
setml(ref_t); ref = mlk(ktest); fl = PL(ktest)
for label, tiers in (
    ('current', [(0.02, 6, 4, 3), (0.06, 8, 5, 4), (1e30, 12, 8, 8), (1e30, 12, 8, 8)]),
    ('mid 10/7/7 <0.3', [(0.02, 6, 4, 3), (0.06, 8, 5, 4), (0.3, 10, 7, 7), (1e30, 12, 8, 8)]),
    ('mid 10/6/6 <0.2', [(0.02, 6, 4, 3), (0.06, 8, 5, 4), (0.2, 10, 6, 6), (1e30, 12, 8, 8)]),
    ('mid 8/6/6 <0.15', [(0.02, 6, 4, 3), (0.06, 8, 5, 4), (0.15, 8, 6, 6), (1e30, 12, 8, 8)]),
    ('top 12/7/7', [(0.02, 6, 4, 3), (0.06, 8, 5, 4), (0.3, 10, 7, 7), (1e30, 12, 7, 7)]),
    ('top 10/8/8', [(0.02, 6, 4, 3), (0.06, 8, 5, 4), (0.3, 10, 7, 7), (1e30, 10, 8, 8)]),
    ('top 12/8/6', [(0.02, 6, 4, 3), (0.06, 8, 5, 4), (0.3, 10, 7, 5), (1e30, 12, 8, 6)]),
    ):
    setml(tiers)
    out = mlk(ktest); nn = E.ik(29, kint, tabs['L'], tabs['L'], 2, qmax=10.)[1]//250
    e = np.abs(out - ref)/np.maximum(np.abs(ref), fl[None,:])
    # vs tight fixture too
    o2 = mlk(kml); e2 = 0
    r2 = np.array([t['ML'][j, w] if not (j==5 and w!=1) else t['ML'][j,w] for j in range(7) for w in range(3)])
    print('%-18s nodes/k %5d  self-conv by k: %s' % (label, nn, ' '.join('%.0e'%x for x in e.max(axis=0))))
print('k', ' '.join('%.0e'%x for x in ktest))
