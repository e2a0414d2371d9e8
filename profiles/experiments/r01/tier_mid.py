"""band 0.3 <= k < 0.5 of the one-loop / I/K operators: what lower orders would cost in accuracy (for the next round)"""
import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
E.L.emul_set_tier.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
ids = [0, 1, 5, 11, 14, 15]
def allk(kk): return np.array([E.ik(kid, kk, tabL, tabL, 0)[0] for kid in ids])
sel = (k1 >= 0.3) & (k1 < 0.5)
kk = np.ascontiguousarray(k1[sel][::3])
fl = PL(kk)
E.L.emul_set_config_default()
for i in range(4): E.L.emul_set_tier(i, [0.02,0.08,0.3,1e30][i], 14, 10, 10)
ref = allk(kk)
for (no, ni, nt) in ((12,7,5), (10,7,5), (10,6,5), (8,6,5)):
    E.L.emul_set_config_default(); E.L.emul_set_tier(3, 0.5, no, ni, nt)
    o = allk(kk); n = E.ik(0, kk, tabL, tabL, 0)[1]/len(kk)
    e = np.abs(o - ref)/np.maximum(np.abs(ref), fl[None, :])
    print('(%d,%d,%d) nodes/k %5d  max err I00 %.1e I01 %.1e I11 %.1e I23 %.1e I32 %.1e I33 %.1e' % ((no, ni, nt, n) + tuple(e.max(axis=1))), flush=True)
