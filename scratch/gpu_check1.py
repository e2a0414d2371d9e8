import sys, time, ctypes as C, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0, '.')
import numpy as np
from pyrsd_b200 import _native as N
from oracle import gcl_ref as R
g = np.load('tests/golden/golden_pt.npz')
cosmo = R.golden_cosmology(g); PL = R.RefLinearPS(cosmo)
k_i, T_i, s8, ns, h, Om, Ob, Tc = cosmo.args
L = N.lib()
ctx = N.Context(0)
def mk_spectra(nc, amps=None):
    k = np.tile(k_i, (nc,1)); T = np.tile(T_i, (nc,1))
    pars = np.zeros((nc,8)); pars[:,0]=ns; pars[:,1]=cosmo.delta_H()**2 if amps is None else amps; pars[:,2]=h; pars[:,3]=Om; pars[:,4]=Ob; pars[:,5]=Tc
    p = C.c_void_p(); N.check(L.prsd_spectra_create(ctx.ptr, nc, len(k_i), k, T, pars, C.byref(p))); return p
def section(name):
    print('=== '+name, flush=True)
try:
    section('spectra')
    sp = mk_spectra(1)
    q = np.exp(np.random.default_rng(0).uniform(np.log(1e-8), np.log(1e5), 4000))
    out = np.empty((1,len(q))); N.check(L.prsd_spectra_plin(sp, len(q), q, out))
    print('plin vs oracle max rel', np.max(np.abs(out[0]/PL(q)-1)))
    M = len(N.knot_grid()); tab = np.empty((1,M)); N.check(L.prsd_spectra_table(sp, 0, tab))
    print('table vs oracle', np.max(np.abs(tab[0]/PL(np.exp(N.knot_grid()))-1)))
except Exception: traceback.print_exc()
try:
    section('Imn/Kmn one-off vs tight oracle')
    ref = np.load('scratch/ref_tight.npy', allow_pickle=True).item()
    ks = np.array([1e-4, 1e-3, 5e-3, 0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 0.8, 1.0, 3.0, 10.0]); Pk = PL(ks)
    for name,(kind,args) in {'I00':('I',(0,0)),'I22':('I',(2,2)),'I13':('I',(1,3)),'I33':('I',(3,3)),'K00':('K',(0,0,0,0)),'K11':('K',(1,1,0,0)),'K20s_b':('K',(2,0,1,1)),'K01':('K',(0,1,0,0))}.items():
        out = np.empty((1,len(ks))); t=time.time()
        if kind=='I': N.check(L.prsd_eval_imn(sp, args[0], args[1], len(ks), ks, None, out))
        else: N.check(L.prsd_eval_kmn(sp, args[0], args[1], args[2], args[3], len(ks), ks, None, out))
        err = np.abs(out[0]-ref[name])/np.maximum(np.abs(ref[name]), Pk)
        print('%-7s'%name, ' '.join('%.0e'%e for e in err), '%.3fs'%(time.time()-t))
except Exception: traceback.print_exc()
try:
    section('Jmn + ps integrals')
    ks2 = np.array([1e-3, 0.02, 0.1, 0.5, 1.0])
    for (m,n) in [(0,0),(0,2),(1,0),(2,0)]:
        out = np.empty((1,len(ks2))); N.check(L.prsd_eval_jmn(sp, m, n, len(ks2), ks2, out))
        print('J%d%d'%(m,n), np.max(np.abs(out[0]/R.Jmn(PL, ks2, m, n, epsrel=1e-10)-1)))
    out = np.empty((1,1)); N.check(L.prsd_eval_ps_integral(sp, 0, 1, np.zeros(1), 0.5, out)); print('sigv2', out[0,0], PL.sigv_tol())
    rs = np.array([0.01, 1., 20., 105., 300., 1000., 1e4, 1e5]); ox = np.empty((1,len(rs))); oy = np.empty((1,len(rs)))
    t=time.time(); N.check(L.prsd_eval_ps_integral(sp, 2, len(rs), rs, 0.5, ox)); N.check(L.prsd_eval_ps_integral(sp, 3, len(rs), rs, 0.5, oy)); print('X/Y build+apply %.3fs'%(time.time()-t))
    print('X abs err', np.abs(ox[0]-PL.X_Zel(rs,1e-9,1e-13))); print('Y abs err', np.abs(oy[0]-PL.Y_Zel(rs,1e-9,1e-13)))
    out = np.empty((1,2)); N.check(L.prsd_eval_ps_integral(sp, 5, 2, np.array([8., 1.]), 0.5, out)); print('Sigma(8,1)', out, PL.Sigma_tol([8.,1.]))
    out = np.empty((1,1)); N.check(L.prsd_eval_ps_integral(sp, 4, 1, np.array([1.0]), 0.5, out)); print('Q3(1)', out[0,0]/PL.Q3_tol([1.0])[0]-1)
except Exception: traceback.print_exc()
try:
    section('plan (250 k_interp)')
    kint = g['k_interp']
    pl = C.c_void_p(); t=time.time(); N.check(L.prsd_plan_create(ctx.ptr, len(kint), kint, None, 3, C.byref(pl))); print('plan build %.2fs'%(time.time()-t))
    info = np.empty(8); N.check(L.prsd_plan_info(pl, info)); print('info', info)
    res = C.c_void_p()
    for rep in range(3):
        t=time.time(); N.check(L.prsd_plan_run(pl, sp, C.byref(res))); print('run ncosmo=1: %.2f ms'%(1e3*(time.time()-t)))
    def get(what, shape):
        o = np.empty(shape); N.check(L.prsd_result_get(res, what, o, o.size)); return o
    nk=len(kint)
    I = get(0,(1,16,nk)); J=get(1,(1,6,nk)); K=get(2,(1,13,nk)); ML=get(3,(1,7,3,nk)); OL=get(4,(1,4,1000)); SK=get(5,(1,nk)); SC=get(6,(1,4)); X0=get(7,(1,1024)); YY=get(8,(1,1024))
    Pk = PL(kint)
    def cmp(name, mine, gold):
        err = np.abs(mine-gold)/np.maximum(np.abs(gold), Pk if len(gold)==nk else 1e-300)
        print('  %-10s vs golden(default eps): max %.1e  median %.1e'%(name, err.max(), np.median(err)))
    for nm, idx in [('I02',2),('I03',3),('I12',6),('I20',8),('I22',10),('I23',11),('I31',13)]: cmp(nm, I[0,idx], g['pt_'+nm])
    for nm, idx in [('J02',2),('J10',3),('J20',5)]: cmp(nm, J[0,idx], g['pt_'+nm])
    for nm, idx in [('K00',0),('K00s',2),('K10',5),('K10s',7),('K11',6),('K11s',8),('K20_a',9),('K20_b',10),('K20s_a',11),('K20s_b',12)]: cmp(nm, K[0,idx], g['pt_'+nm])
    for nm, idx in [('Ivvdd_h01',0),('Ivvdd_h02',1),('Idvdv_h03',2),('Idvdv_h04',3)]:
        for j,part in enumerate(['lin','cross','1loop']): cmp(nm+'_'+part, ML[0,idx,j], g['pt_'+nm][j])
    for j,nm in enumerate(['_Pdd_0','_Pdv_0','_Pvv_0','_P22bar_0']):
        gold=g['oneloop'+nm]; err=np.abs(OL[0,j]-gold)/np.maximum(np.abs(gold), PL(R.logspace(1e-5,10,1000)) ); print('  oneloop%s max %.1e median %.1e'%(nm, err.max(), np.median(err)))
    cmp('sigmasq_k', SK[0], g['pt_sigmasq_k'])
    print('  sigma_v^2', SC[0,0], 'golden zel_sigmasq', float(g['zel_sigmasq']), 'kurtosis', SC[0,1], float(g['velocity_kurtosis_unnormed']), 'Q3/k^4', SC[0,2], PL.Q3_tol([1.0])[0])
    print('  X0 abs diff vs golden max', np.max(np.abs(X0[0]-g['zel_X0'])), ' YY', np.max(np.abs(YY[0]-g['zel_YY'])))
    np.savez('gpurun_out/plan_result_golden.npz', I=I, J=J, K=K, ML=ML, OL=OL, SK=SK, SC=SC, X0=X0, YY=YY)
    section('plan batch timing')
    for nc in [8, 64]:
        spb = mk_spectra(nc, amps=cosmo.delta_H()**2*np.linspace(0.5,1.5,nc))
        resb = C.c_void_p()
        for rep in range(3):
            t=time.time(); N.check(L.prsd_plan_run(pl, spb, C.byref(resb))); dt=time.time()-t
        print('run ncosmo=%d: %.2f ms  (%.3f ms / cosmology)'%(nc, 1e3*dt, 1e3*dt/nc))
        o = np.empty((nc,16,nk)); N.check(L.prsd_result_get(resb, 0, o, o.size))
        amp = np.linspace(0.5,1.5,nc)
        print('  scaling check I00 ~ amp^2: max rel', np.max(np.abs(o[:,0,:]/ (amp[:,None]**2*I[0,0][None,:]) - 1)))
except Exception: traceback.print_exc()
print('launches', ctx.launches())
