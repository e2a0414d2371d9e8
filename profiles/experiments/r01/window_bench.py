"""reference-mode survey likelihood: ONE tabulated cosmology, nw walkers; P(k, mu) on the window grid (1024 x 40)
-> multipoles -> window convolution (6 FFTLogs of N = 4096 per walker) -> spline onto the data k -> chi^2, all on
the device (PTEngine.galaxy_window_chi2)"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np
import torch
import bench
from pyrsd_b200 import rsd, transfers
from window_inputs import synthetic_window
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
data = bench.synthetic_batch(nw, 0)
eng = rsd.PTEngine()
sp = eng.spectra(data['k'][:1], data['T'][:1], data['pars'][0, 0], data['pars'][0, 1], data['pars'][0, 2], data['pars'][0, 3], data['pars'][0, 4], data['pars'][0, 5])
res = eng.run(sp)
kmin, kmax = 1e-4, 0.7
kmodel = np.logspace(np.log10(kmin), np.log10(kmax), 200)
gal = eng.galaxy(kmodel, rsd.galaxy_config())
tr = transfers.WindowFunctionTransfer(synthetic_window(), [0, 2, 4], kmin=kmin, kmax=kmax)
cmap = np.zeros(nw, dtype=np.int32)
k_out = np.linspace(0.01, 0.4, 40)
n = 3*len(k_out)
rng = np.random.default_rng(1)
wpar = data['wpar'].copy(); wpar[:, 2:5] = 1.0            # no AP remapping beyond the model grid
chi2, poles = eng.galaxy_window_chi2(gal, sp, res, wpar, data['stoch'], tr, k_out, np.zeros(n), np.eye(n), cmap=cmap, return_poles=True)
d = poles[0].cpu().numpy().ravel()
cinv = np.diag(1./(0.05*np.abs(d) + 1.)**2)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    chi2 = eng.galaxy_window_chi2(gal, sp, res, wpar, data['stoch'], tr, k_out, d, cinv, cmap=cmap)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print('step %d: windowed chi2 of %d walkers %.2f ms = %.2f us per walker (finite: %s)' % (it, nw, 1e3*(t1 - t0), 1e6*(t1 - t0)/nw, np.isfinite(chi2).all()), flush=True)
# stages
P = torch.rand((nw, 1024, 40), dtype=torch.float64, device='cuda')
from pyrsd_b200.transfers import GriddedMultipoleTransfer
for name, fn in (('grid multipoles + window + spline', lambda: tr(P, k_out=k_out)),
                 ('grid multipoles + window', lambda: tr(P)),
                 ('grid multipoles + zero padding', lambda: tr(P, no_convolution=True)),
                 ('grid multipoles only', lambda: GriddedMultipoleTransfer.__call__(tr, P))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print('%s: %.2f ms per step' % (name, 1e3*(time.perf_counter() - t0)/5))

# kernel-family breakdown of one likelihood step (CUDA events inside the library)
import ctypes as C
from pyrsd_b200 import _native as N
L = N.lib()
N.check(L.prsd_ctx_profile(eng.ctx.ptr, 1))
chi2 = eng.galaxy_window_chi2(gal, sp, res, wpar, data['stoch'], tr, k_out, d, cinv, cmap=cmap)
ms, cnt = np.zeros(8), np.zeros(8, dtype=np.int64)
L.prsd_ctx_profile_read.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.float64), np.ctypeslib.ndpointer(dtype=np.int64)]
N.check(L.prsd_ctx_profile_read(eng.ctx.ptr, ms, cnt))
N.check(L.prsd_ctx_profile(eng.ctx.ptr, 0))
print('families [IK, linear, zel, galaxy(+transfers), splines, fftlog, ML, diag] ms:', np.round(ms, 3), cnt)
