"""Dense GPU parity: the batched plan (the path bench.py times) against the reference C++ at tight tolerance on ALL
250 ``k_interp`` points -- 16 I_mn, 13 K_mn, 6 J_mn, sigma_v^2(k), the four one-loop tables on their whole 1000-point
grid and the 7 x 3 ImnOneLoop parts -- for the golden Planck15 spectrum and for cosmologies drawn from bench.py's
own synthetic batch (tests/golden/make_tight_dense.py -> tight_dense_<name>.npz).  The 250 points straddle every
hard switch of the fixed rule (k = 0.02, 0.06, 0.08, 0.3) that the sparse fixtures of round 1 could not show.

Metric: |mine - ref| / max(|ref|, P_L(k)) for I, K and the one-loop tables (the reference stops on
epsabs = eps P_L / V, Imn.cpp:129), plain relative for J and sigma_v^2(k); for ImnOneLoop the floor is the scale of
the integral as the model reads it (gpu_common.ml_scale_floor: the three parts are only used summed and splined in k,
and the f32 / f33 parts change sign between tabulated points); the strict max(|ref|, P_L) figure is recorded too.  Bound: 1e-5 everywhere the
model can read (k <= 1 = INTERP_KMAX); the one-loop tables for 1 < k <= 10 (integrands of ImnOneLoop only) are held
to 5e-5 of the table value (measured <= 3.1e-5): 2 I + 6 k^2 J P_L cancels 70-140 x there, i.e. the TERMS are matched
to < 4e-7
(measured budget in profiles/r02_parity.md).

The worst error per family and cosmology is written to gpurun_out/r02_parity.json."""
import json
import os

import numpy as np
import pytest

from tests.gpu_common import relerr_floor, ml_scale_floor

pytestmark = pytest.mark.gpu
TOL = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ['golden', 'bench0', 'bench1', 'benchmax', 'benchmin']
MLSPEC = [('vv', 'dd', 0, 1), ('vv', 'dd', 0, 2), ('dv', None, 0, 3), ('dv', None, 0, 4),
          ('vv', None, 2, 3), ('vv', None, 3, 2), ('vv', None, 3, 3)]


def _load(name):
    path = os.path.join(HERE, 'golden', 'tight_dense_%s.npz' % name)
    if not os.path.exists(path):
        pytest.skip("tests/golden/tight_dense_%s.npz not generated" % name)
    return np.load(path)


def dense_cosmology(golden, d):
    """pygcl.Cosmology of a dense fixture: golden T(k) tilted by (k / 0.05)^(delta / 2), normalised to sigma8"""
    from pyrsd_b200 import pygcl
    params = {k[len('cosmo_'):]: float(golden[k]) for k in golden.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['T_cmb'] = 2.7255
    k_i = golden['cosmo_k']
    T_i = golden['cosmo_T']*(k_i/0.05)**(0.5*float(d['delta']))
    return pygcl.Cosmology(params, 0, float(golden['cosmo_sigma8']), k_i, T_i)


def _record(name, worst):
    out = os.path.join(os.path.dirname(HERE), 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, 'r02_parity.json')
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur.setdefault(name, {}).update(worst)
        json.dump(cur, open(path, 'w'), indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope='module')
def engine():
    from pyrsd_b200 import rsd
    return rsd.PTEngine()


@pytest.mark.parametrize('name', NAMES)
def test_plan_vs_dense_tight_reference(engine, golden, name):
    from pyrsd_b200 import pygcl
    d = _load(name)
    cosmo = dense_cosmology(golden, d)
    assert abs(cosmo.delta_H()/float(d['delta_H']) - 1) < 1e-6
    eng = engine
    assert np.max(np.abs(eng.k_interp/d['k'] - 1)) < 1e-13
    sp = eng.spectra(cosmo.GetDiscreteK(), cosmo.GetDiscreteTk(), cosmo.n_s(), cosmo.delta_H()**2, cosmo.h(),
                     cosmo.Omega0_m(), cosmo.Omega0_b(), cosmo.Tcmb())
    res = eng.run(sp)
    k, floor = d['k'], d['plin_k']
    worst = {}

    def check(tag, err, tol=TOL):
        i = int(np.argmax(err))
        worst[tag] = [float(err[i]), float(k[i]) if len(err) == len(k) else float(i)]
        assert err[i] < tol, (name, tag, float(err[i]), worst[tag][1])

    I, J, K = eng.get(res, 'I')[0], eng.get(res, 'J')[0], eng.get(res, 'K')[0]
    for j in range(16):
        check('I%d%d' % (j//4, j % 4), relerr_floor(I[j], d['I'][j], floor))
    for j in range(13):
        check('K[%d]' % j, relerr_floor(K[j], d['K'][j], floor))
    for j, mn in enumerate(('00', '01', '02', '10', '11', '20')):
        check('J' + mn, np.abs(J[j]/d['J'][j] - 1))
    # sigma_v^2(k) = integral over [1e-5, k/2].  The synthetic tilt (k / 0.05)^(delta / 2) only multiplies the TABULATED
    # T(k); below its first sample k_i[0] = 1.046e-5 the Eisenstein-Hu extrapolation is untilted, so the tilted spectra
    # jump by up to 40 % there: a k whose upper limit k/2 lies within one table cell of the jump is left out (3 of 250)
    esig = np.abs(eng.get(res, 'sigmasq_k')[0]/d['sigmasq_k'] - 1)
    if name != 'golden':
        esig = np.where((0.5*k > 1e-5) & (0.5*k < 1.25e-5), 0.0, esig)
    check('sigmasq_k', esig)
    sc = eng.get(res, 'scalars')[0]
    assert abs(sc[0]/float(d['sigma_v2']) - 1) < 1e-6
    # one-loop tables, whole 1000-point grid
    ol = eng.get(res, 'oneloop')[0]
    k1 = np.exp(np.log(1e-5) + np.arange(1000)*np.log(1e6)/999)
    for j, w in enumerate(('dd', 'dv', 'vv', '22bar')):
        err = relerr_floor(ol[j], d['oneloop_common'][j], d['oneloop_plin'])
        lo = k1 <= 1.0
        worst['oneloop_%s(k<=1)' % w] = [float(err[lo].max()), float(k1[lo][np.argmax(err[lo])])]
        worst['oneloop_%s(1<k<=10)' % w] = [float(err[~lo].max()), float(k1[~lo][np.argmax(err[~lo])])]
        _record(name, worst)
        assert err[lo].max() < TOL, (name, w, err[lo].max())
        assert err[~lo].max() < 5e-5, (name, w, err[~lo].max())
    assert abs(sc[1]/float(d['kurtosis_common_tables']) - 1) < TOL
    # ImnOneLoop of the plan: built on the plan's OWN one-loop tables (the full pipeline), against the reference's
    # integrals over the tight common tables
    ML = eng.get(res, 'ML')[0]
    for j, (p1, p2, m, n) in enumerate(MLSPEC):
        scale = ml_scale_floor(d['ML'][j], floor)
        for w, part in enumerate(('lin', 'cross', '1loop')):
            strict = relerr_floor(ML[j, w], d['ML'][j, w], floor)
            worst['strict_ML%d%d_%s' % (m, n, part)] = [float(strict.max()), float(k[np.argmax(strict)])]
            try:
                check('ML%d%d_%s' % (m, n, part), np.abs(ML[j, w] - d['ML'][j, w])/scale)
            finally:
                _record(name, worst)
    _record(name, worst)


@pytest.mark.parametrize('name', ['golden', 'benchmax'])
def test_imn_oneloop_stage_on_common_tables_dense(golden, name):
    """the ImnOneLoop stage in isolation on all 250 k: both sides read the SAME tight one-loop tables"""
    from pyrsd_b200 import pygcl
    d = _load(name)
    cosmo = dense_cosmology(golden, d)
    plin = pygcl.LinearPS(cosmo, 0.)
    common = d['oneloop_common']
    ol = {w: cls(plin, 1e-7, common[j]) for j, (w, cls) in
          enumerate((('dd', pygcl.OneLoopPdd), ('dv', pygcl.OneLoopPdv), ('vv', pygcl.OneLoopPvv)))}
    k, floor = d['k'], d['plin_k']
    worst = {}
    for j, (p1, p2, m, n) in enumerate(MLSPEC):
        Iml = pygcl.ImnOneLoop(ol[p1]) if p2 is None else pygcl.ImnOneLoop(ol[p1], ol[p2], 1e-4)
        scale = ml_scale_floor(d['ML'][j], floor)
        for w, fn in enumerate((Iml.EvaluateLinear, Iml.EvaluateCross, Iml.EvaluateOneLoop)):
            mine = fn(k, m, n)
            strict = relerr_floor(mine, d['ML'][j, w], floor)
            worst['strict_stage_ML%d%d_%d' % (m, n, w)] = [float(strict.max()), float(k[np.argmax(strict)])]
            err = np.abs(mine - d['ML'][j, w])/scale
            worst['stage_ML%d%d_%d' % (m, n, w)] = [float(err.max()), float(k[np.argmax(err)])]
            _record(name, worst)
            assert err.max() < TOL, (name, p1, p2, m, n, w, err.max(), k[np.argmax(err)])
