"""one-loop operator at k > 0.5: error of reduced Gauss-Legendre orders against a converged rule (14, 10, 10),
relative to max(|I|, P_L(k)) -- how much of the 45 % of one-loop nodes that sit above k = 0.5 could go"""
import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
E.L.emul_set_tier.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
ids = [0, 1, 5, 11, 14, 15]            # I00 I01 I11 I23 I32 I33
def allk(kk):
    return np.array([E.ik(kid, kk, tabL, tabL, 0)[0] for kid in ids])
for lo, hi in ((0.5, 1.0), (1.0, 3.0), (3.0, 10.0)):
    sel = (k1 >= lo) & (k1 < hi)
    kk = np.ascontiguousarray(k1[sel][::max(1, sel.sum()//8)])
    fl = PL(kk)
    E.L.emul_set_config_default()
    for i in range(4): E.L.emul_set_tier(i, [0.02,0.08,0.3,1e30][i], 14, 10, 10)
    ref = allk(kk)
    for (no, ni, nt) in ((12,7,5), (10,6,5), (8,6,5), (8,5,4), (6,5,4)):
        E.L.emul_set_config_default(); E.L.emul_set_tier(3, 1e30, no, ni, nt)
        o = allk(kk); n = E.ik(0, kk, tabL, tabL, 0)[1]/len(kk)
        e = np.abs(o - ref)/np.maximum(np.abs(ref), fl[None, :])
        # the one-loop spectra are 2 I + 6 k^2 J P_L: their error is 2x the I error; P22bar = I23 + 2/3 I32 + 1/5 I33
        print('k in [%.1f, %.1f): (%d,%d,%d) nodes/k %5d  max err I00 %.1e I01 %.1e I11 %.1e I23 %.1e I32 %.1e I33 %.1e' % ((lo, hi, no, ni, nt, n) + tuple(e.max(axis=1))), flush=True)
