import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
spec = [('vv','dd',29),('vv','dd',30),('dv','dv',25),('dv','dv',26),('vv','vv',11),('vv','vv',14),('vv','vv',15)]
E.L.emul_set_tier.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
kint = np.logspace(np.log10(5e-6), 0, 250)
ktest = np.ascontiguousarray(kint[192:250:5])
def mlk(kk):
    res = []
    for j,(p1,p2,kid) in enumerate(spec):
        res.append(E.ik(kid, kk, tabs['L'], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs['L'], tabs[p2], 2, qmax=10.)[0] + E.ik(kid, kk, tabs[p1], tabs['L'], 2, qmax=10.)[0])
        res.append(E.ik(kid, kk, tabs[p1], tabs[p2], 2, qmax=10.)[0])
    return np.array(res)
def setml(t2, t3):
    E.L.emul_set_config_default(); E.L.emul_use_ml_variant()
    E.L.emul_set_tier(2, 0.3, *t2); E.L.emul_set_tier(3, 1e30, *t3)
setml((16,12,12),(16,12,12)); ref = mlk(ktest); fl = PL(ktest)
for t2, t3 in (((10,7,7),(10,8,8)), ((10,7,6),(10,8,7)), ((10,7,5),(10,8,6)), ((9,7,7),(9,8,8)), ((8,7,7),(8,8,8)), ((10,6,6),(10,7,7)), ((9,6,6),(9,7,7))):
    setml(t2, t3)
    out = mlk(ktest); nn = E.ik(29, kint, tabs['L'], tabs['L'], 2, qmax=10.)[1]//250
    e = np.abs(out - ref)/np.maximum(np.abs(ref), fl[None,:])
    o2 = mlk(kml); e2 = 0
    for j in range(7):
        for w in range(3):
            e2 = max(e2, (np.abs(o2[3*j+w] - t['ML'][j,w])/np.maximum(np.abs(t['ML'][j,w]), floor)).max())
    print(t2, t3, 'nodes/k %d  self-conv max %.1e  vs tight %.1e   by k: %s' % (nn, e.max(), e2, ' '.join('%.0e'%x for x in e.max(axis=0))))
