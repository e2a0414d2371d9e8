import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyrsd_b200 import pygcl, rsd
from tests.gpu_common import golden_pygcl_cosmology, b2_00, b2_01, stochasticity
g = np.load(os.path.join(ROOT, 'tests/golden/golden_pt.npz'))
ref = np.load(os.path.join(ROOT, 'tests/golden/galaxy_ref.npz'))
st = np.load(os.path.join(ROOT, 'profiles/experiments/zel_state_tight.npy'))
cosmo = golden_pygcl_cosmology(g)
eng = rsd.PTEngine()
model = rsd.GalaxySpectrum(params=cosmo, z=0., b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity, engine=eng)
dm = model.dm_spectra()
k = model.k
r = ref['z0_default_dm']
sel = (k > 3.5e-3) & (k < 9e-3)
print('k        ', np.array2string(k[sel], precision=5))
print('P00 err  ', np.array2string(dm['P00'][sel]/r[1][sel] - 1, precision=2))
print('P01 err  ', np.array2string(dm['P01'][sel]/r[2][sel] - 1, precision=2))
print('Plin err ', np.array2string(dm['P_lin'][sel]/r[0][sel] - 1, precision=2))
X0, YY, sc = model.pt_table('X0'), model.pt_table('YY'), model.pt_table('scalars')
ssq, Xt, Yt = st[0], st[1:1025], st[1025:]
print('sigma^2 rel', sc[0]/ssq - 1, ' X0 max abs/ssq %.2e  YY max abs/ssq %.2e' % (np.max(np.abs(X0 - Xt))/ssq, np.max(np.abs(YY - Yt))/ssq))
# one-off Zel'dovich with both states at the model k
s8 = cosmo.sigma8()
for tag, (a, b, c) in (('plan state', (sc[0], X0, YY)), ('tight state', (ssq, Xt, Yt))):
    z = pygcl.ZeldovichP00(pygcl.ZeldovichPS(cosmo, True, s8, 5e-3, a, b, b + 2*a, c))
    print(tag, np.array2string(z(k[sel]), precision=8))
from oracle import gcl_ref as R
rc = R.golden_cosmology(g)
zr = R.RefZeldovich(rc, True, (s8, 5e-3, ssq, Xt, Xt + 2*ssq, Yt))
print('oracle tight', np.array2string(zr('P00', k[sel], 1, 5e-3), precision=8))
print('ref P00     ', np.array2string(r[1][sel], precision=8))
print('gpu P00     ', np.array2string(dm['P00'][sel], precision=8))
rr = 10.**(1.5 + (np.arange(1, 1025) - 512.5)*(7./1024))
print('r, (X0 - tight)/ssq, (YY - tight)/ssq, (X0+YY)/ssq:')
for i in range(0, 1024, 32):
    print('%10.3e  %10.2e  %10.2e  %10.3e' % (rr[i], (X0[i] - Xt[i])/ssq, (YY[i] - Yt[i])/ssq, (Xt[i] + Yt[i])/ssq))
i = np.argmax(np.abs(X0 - Xt)); j = np.argmax(np.abs(YY - Yt))
print('worst X0 at r=%.3e: %.2e ; worst YY at r=%.3e: %.2e' % (rr[i], (X0[i]-Xt[i])/ssq, rr[j], (YY[j]-Yt[j])/ssq))
