"""which public names of the reference's model classes (reference_public_names.json: the def / class-level names of
GalaxySpectrum, BiasedSpectrum, DarkMatterSpectrum, listed by ast in the build container) does an instance of ours answer?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from pyrsd_b200 import pygcl, rsd, dm
from pyrsd_b200.synthetic import base_cosmology, b2_00, b2_01, stochasticity
names = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_public_names.json')))
c = base_cosmology()
params = dict(h=c['h'], Omega_b=c['Omega0_b'], Omega_cdm=c['Omega0_m'] - c['Omega0_b'], n_s=c['n_s'], T_cmb=c['Tcmb'],
              growth={0.0: 1.0, 0.55: 0.7486}, f={0.0: 0.5, 0.55: 0.7602}, H={0.0: 100.*c['h'], 0.55: 133.86*c['h']})
cosmo = pygcl.Cosmology(params, 0, c['sigma8'], c['k'], c['T'])
eng = rsd.PTEngine()
kw = dict(params=cosmo, z=0.55, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity, engine=eng)
objs = {'GalaxySpectrum': (rsd.GalaxySpectrum(**kw), ['GalaxySpectrum', 'BiasedSpectrum', 'DarkMatterSpectrum']),
        'BiasedSpectrum': (dm.BiasedSpectrum(interpolate=False, **kw), ['BiasedSpectrum', 'DarkMatterSpectrum']),
        'DarkMatterSpectrum': (dm.DarkMatterSpectrum(params=cosmo, z=0.55, engine=eng, interpolate=False), ['DarkMatterSpectrum'])}
for cls, (obj, srcs) in objs.items():
    want = sorted(set(n for s in srcs for n in names[s]))
    missing = []
    for n in want:
        try:
            ok = hasattr(obj, n)
        except Exception as e:
            ok = True       # the attribute exists and raised for another reason
        if not ok:
            missing.append(n)
    print(cls, 'answers %d of %d;' % (len(want) - len(missing), len(want)), 'missing:', missing, flush=True)
