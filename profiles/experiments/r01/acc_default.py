"""Accuracy of the library's DEFAULT quadrature configuration (tiers) in CPU emulation vs the tight fixture."""
import sys, ctypes as C
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
kt = t['k']; Pk = t['plin_k']
KSPEC = list(range(16, 29))
spec = [('vv','dd',29),('vv','dd',30),('dv','dv',25),('dv','dv',26),('vv','vv',11),('vv','vv',14),('vv','vv',15)]
P1 = PL(k1)
E.L.emul_set_config_default()
wI = np.zeros(len(kt)); wK = np.zeros(len(kt))
for kid in range(16):
    out, nI = E.ik(kid, kt, tabL, tabL, 0)
    wI = np.maximum(wI, np.abs(out - t['I'][kid])/np.maximum(np.abs(t['I'][kid]), Pk))
for j, kid in enumerate(KSPEC):
    out, _ = E.ik(kid, kt, tabL, tabL, 1)
    wK = np.maximum(wK, np.abs(out - t['K'][j])/np.maximum(np.abs(t['K'][j]), Pk))
print('k ', ' '.join('%.0e'%x for x in kt)); print('I ', ' '.join('%.0e'%x for x in wI)); print('K ', ' '.join('%.0e'%x for x in wK))
kint = np.logspace(np.log10(5e-6), 0, 250)
print('nodes: ik %d per k_interp'%(E.ik(0, kint, tabL, tabL, 0)[1]//250), ' 1loop %d per k'%(E.ik(0, np.ascontiguousarray(k1[::4]), tabL, tabL, 0)[1]//250))
sel = slice(0, 1000, 5); kk = np.ascontiguousarray(k1[sel]); w1 = np.zeros(len(kk))
for j,(kid,jid) in enumerate(((0,0),(1,1),(5,4))):
    I, n1 = E.ik(kid, kk, tabL, tabL, 0)
    J = E.linear(0, jid, kk, 1e-5, 1e5, tabL)
    w1 = np.maximum(w1, np.abs(2*I + 6*kk**2*J*P1[sel] - com[j][sel])/np.maximum(np.abs(com[j][sel]), P1[sel]))
I23,_ = E.ik(11, kk, tabL, tabL, 0); I32,_ = E.ik(14, kk, tabL, tabL, 0); I33,_ = E.ik(15, kk, tabL, tabL, 0)
e22 = np.abs(I23 + 2./3*I32 + 0.2*I33 - com[3][sel])/np.maximum(np.abs(com[3][sel]), P1[sel])
print('1loop: k<0.5 %.1e  all %.1e  P22 %.1e'%(w1[kk<0.5].max(), w1.max(), e22.max()))
for lo, hi in ((0,0.02),(0.02,0.08),(0.08,0.3),(0.3,0.5),(0.5,10.1)):
    s_ = (kk>=lo)&(kk<hi); print('   [%g,%g) 1loop %.1e P22 %.1e'%(lo,hi,w1[s_].max(), e22[s_].max()))
E.L.emul_use_ml_variant()
wM = np.zeros(len(kml))
for j,(p1,p2,kid) in enumerate(spec):
    lin, nM = E.ik(kid, kml, tabs['L'], tabs['L'], 2, qmax=10.)
    c1, _ = E.ik(kid, kml, tabs['L'], tabs[p2], 2, qmax=10.)
    c2, _ = E.ik(kid, kml, tabs[p1], tabs['L'], 2, qmax=10.)
    ol, _ = E.ik(kid, kml, tabs[p1], tabs[p2], 2, qmax=10.)
    if j == 5:
        lin, _ = E.ik(11, kml, tabs['L'], tabs['L'], 2, qmax=10.); ol, _ = E.ik(11, kml, tabs[p1], tabs[p2], 2, qmax=10.)
    for w, v in enumerate((lin, c1+c2, ol)):
        wM = np.maximum(wM, np.abs(v-t['ML'][j,w])/np.maximum(np.abs(t['ML'][j,w]), floor))
print('kml', ' '.join('%.0e'%x for x in kml)); print('ML ', ' '.join('%.0e'%x for x in wM))
print('nodes: ml %d per k_interp'%(E.ik(29, kint, tabs['L'], tabs['L'], 2, qmax=10.)[1]//250))
