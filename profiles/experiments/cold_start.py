"""cold start: where do the seconds of a first model evaluation go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
t0 = time.perf_counter()
import numpy as np
from pyrsd_b200 import pygcl, rsd, _native as N
from pyrsd_b200.synthetic import base_cosmology, b2_00, b2_01, stochasticity
t1 = time.perf_counter()
ctx = N.default_context(0)
t2 = time.perf_counter()
eng = rsd.PTEngine(context=ctx)
t3 = time.perf_counter()
print('import %.2f s, context %.2f s, PTEngine %.2f s' % (t1 - t0, t2 - t1, t3 - t2), eng.build_profile(), flush=True)
c = base_cosmology()
params = dict(h=c['h'], Omega_b=c['Omega0_b'], Omega_cdm=c['Omega0_m'] - c['Omega0_b'], n_s=c['n_s'], T_cmb=c['Tcmb'],
              growth={0.0: 1.0, 0.55: 0.7486}, f={0.0: 0.5, 0.55: 0.7602}, H={0.0: 100.*c['h'], 0.55: 133.86*c['h']})
cosmo = pygcl.Cosmology(params, 0, c['sigma8'], c['k'], c['T'])
t4 = time.perf_counter()
model = rsd.GalaxySpectrum(params=cosmo, z=0.55, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity, engine=eng)
P = model.power(np.logspace(-2, -0.4, 100), np.linspace(0, 1, 5))
t5 = time.perf_counter()
P = model.power(np.logspace(-2, -0.4, 100), np.linspace(0, 1, 5))
t6 = time.perf_counter()
print('Cosmology %.3f s, first GalaxySpectrum.power %.3f s, second %.4f s' % (t4 - t3, t5 - t4, t6 - t5), flush=True)
t7 = time.perf_counter()
z = pygcl.ZeldovichP00(cosmo, 0.55)(np.logspace(-2, -0.4, 50))
t8 = time.perf_counter()
print('first pygcl.ZeldovichP00(cosmo, z)(k): %.3f s' % (t8 - t7))
