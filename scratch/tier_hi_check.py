import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
E.L.emul_set_tier.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
ids = [0, 1, 5, 11, 14, 15]
def allk(kk): return np.array([E.ik(kid, kk, tabL, tabL, 0)[0] for kid in ids])
sel = k1 >= 0.3
kk = np.ascontiguousarray(k1[sel][::9])
fl = PL(kk)
E.L.emul_set_config_default()
for i in range(4): E.L.emul_set_tier(i, [0.02,0.08,0.3,1e30][i], 14, 10, 10)
ref = allk(kk)
E.L.emul_set_config_default()
o = allk(kk)
e = np.abs(o - ref)/np.maximum(np.abs(ref), fl[None, :])
print('default config, k in [0.3, 10]: max err per kernel', ' '.join('%.1e' % x for x in e.max(axis=1)), ' worst k', kk[e.max(axis=0).argmax()])
tot = sum(E.ik(0, np.array([k]), tabL, tabL, 0)[1] for k in k1[::5])*5
kint = np.logspace(np.log10(5e-6), 0, 250)
tik = sum(E.ik(0, np.array([k]), tabL, tabL, 0)[1] for k in kint)
print('one-loop nodes ~%.2f M, I/K nodes %.3f M' % (tot/1e6, tik/1e6))
