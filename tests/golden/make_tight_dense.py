"""Generate ``tests/golden/tight_dense_<name>.npz``: the reference C++ (oracle/_ref) at TIGHT ``epsrel`` on ALL 250
``k_interp`` points for the 16 I_mn, 13 K_mn, 6 J_mn and the 7 x 3 ImnOneLoop parts, on

  golden   the reference's shipped Planck15 transfer function (the spectrum the fixed rule was tuned on)
  bench<i> cosmology i of rank 0's synthetic batch of ``bench.py`` (same seed, same tilt recipe): i = 0, 1 and the
           two extreme tilts of the 256 (largest positive / negative delta)

Same tolerances and same recipe for the common one-loop tables as ``make_tight.py`` (Imn 1e-9 / 1e-8 for the tables,
Kmn 1e-9, Jmn 1e-10, ImnOneLoop 1e-8).  The 250 points straddle every hard switch of the fixed rule (k = 0.02, 0.06,
0.08, 0.3, 0.5), which the 19-point fixture of round 1 could not show.

Run in the build container only (8 cores, about a quarter of an hour per cosmology):
    python tests/golden/make_tight_dense.py [golden bench0 ...]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import gcl_ref as R  # noqa: E402

SEED = 20261019           # bench.py:SEED
NCHUNK = 10               # k points are dealt round-robin into NCHUNK tasks per kernel (load balance)

KSPEC = [(0, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 1, 1, 0), (0, 2, 1, 0), (1, 0, 0, 0), (1, 1, 0, 0),
         (1, 0, 1, 0), (1, 1, 1, 0), (2, 0, 0, 0), (2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 1, 1)]
JSPEC = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]
MLSPEC = [('vv', 'dd', 0, 1), ('vv', 'dd', 0, 2), ('dv', None, 0, 3), ('dv', None, 0, 4),
          ('vv', None, 2, 3), ('vv', None, 3, 2), ('vv', None, 3, 3)]


def bench_deltas():
    """the tilts of rank 0's batch of 256 (bench.py:synthetic_batch draws delta first)"""
    return np.random.default_rng(SEED).normal(0, 0.02, 256)


def spec_of(name):
    """(delta, n_s scale) of a named cosmology"""
    if name == 'golden':
        return 0.0
    d = bench_deltas()
    if name == 'benchmax':
        return float(d[np.argmax(d)])
    if name == 'benchmin':
        return float(d[np.argmin(d)])
    return float(d[int(name[len('bench'):])])


def cosmology(name):
    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    delta = spec_of(name)
    k = np.asarray(g['cosmo_k'], dtype=float)
    T = np.asarray(g['cosmo_T'], dtype=float)*(k/0.05)**(0.5*delta)
    h = float(g['cosmo_h'])
    Om = float(g['cosmo_Omega_b']) + float(g['cosmo_Omega_cdm']) + float(g['cosmo_m_ncdm'])/93.14/h**2
    return R.RefCosmology(k, T, float(g['cosmo_sigma8']), float(g['cosmo_n_s']), h, Om, float(g['cosmo_Omega_b']), 2.7255)


_state = {}


def _pl(name):
    if name not in _state:
        c = cosmology(name)
        _state[name] = (c, R.RefLinearPS(c))
    return _state[name][1]


def _ol(name, common):
    key = name + ':ol'
    if key not in _state:
        PL = _pl(name)
        _state[key] = {w: R.RefOneLoop(PL, w, 1e-7, common[i]) for i, w in enumerate(('dd', 'dv', 'vv', '22bar'))}
    return _state[key]


def work(task):
    name, kind, args, k = task[:4]
    PL = _pl(name)
    if kind == 'I':
        return R.Imn(PL, k, args[0], args[1], epsrel=args[2])
    if kind == 'J':
        return R.Jmn(PL, k, args[0], args[1], epsrel=1e-10)
    if kind == 'K':
        m, n, tid, part = args
        return R.Kmn(PL, k, m, n, bool(tid), part, epsrel=1e-9)
    if kind == 'S':
        return np.array([PL.sigv_tol(0.5*kk) if 0.5*kk >= 1e-5 else PL.VelocityDispersion(np.array([kk]), 0.5)[0] for kk in k])
    if kind == 'ML':
        ol = _ol(name, task[4])
        p1, p2, m, n, which = args
        return R.ImnOneLoop(ol[p1], None if p2 is None else ol[p2], k, m, n, which, epsrel=1e-8)
    raise ValueError(kind)


def scatter(pool, name, jobs, k, extra=()):
    """jobs: list of (kind, args); returns array [len(jobs), len(k)]"""
    tasks, where = [], []
    for j, (kind, args) in enumerate(jobs):
        for c in range(NCHUNK):
            idx = np.arange(c, len(k), NCHUNK)
            tasks.append((name, kind, args, np.ascontiguousarray(k[idx])) + tuple(extra))
            where.append((j, idx))
    out = np.empty((len(jobs), len(k)))
    for (j, idx), val in zip(where, pool.map(work, tasks, chunksize=1)):
        out[j, idx] = val
    return out


def run(name, pool):
    t0 = time.time()
    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    kint = np.ascontiguousarray(g['k_interp'])
    c = cosmology(name)
    PL = R.RefLinearPS(c)
    out = {'name': name, 'delta': spec_of(name), 'n_s': float(g['cosmo_n_s']), 'delta_H': c.delta_H(), 'k': kint,
           'plin_k': PL(kint), 'sigma_v2': PL.sigv_tol()}
    jobs = [('I', (m, n, 1e-9)) for m in range(4) for n in range(4)] + [('J', mn) for mn in JSPEC] \
        + [('K', s) for s in KSPEC] + [('S', ())]
    res = scatter(pool, name, jobs, kint)
    out['I'], out['J'], out['K'], out['sigmasq_k'] = res[:16], res[16:22], res[22:35], res[35]
    print(name, 'I/J/K done %.0fs' % (time.time() - t0), flush=True)
    # common tight one-loop tables (OneLoopPS.cpp:56-185 from Imn at 1e-8 and Jmn at 1e-10, see make_tight.py)
    k1 = R.logspace(1e-5, 10, 1000)
    P1 = PL(k1)
    jobs = [('I', (0, 0, 1e-8)), ('I', (0, 1, 1e-8)), ('I', (1, 1, 1e-8)), ('I', (2, 3, 1e-8)), ('I', (3, 2, 1e-8)),
            ('I', (3, 3, 1e-8)), ('J', (0, 0)), ('J', (0, 1)), ('J', (1, 1))]
    t = scatter(pool, name, jobs, k1)
    common = np.array([2*t[0] + 6*k1*k1*t[6]*P1, 2*t[1] + 6*k1*k1*t[7]*P1, 2*t[2] + 6*k1*k1*t[8]*P1,
                       t[3] + (2./3.)*t[4] + (1./5.)*t[5]])
    out['oneloop_common'], out['oneloop_plin'] = common, P1
    out['kurtosis_common_tables'] = R.RefOneLoop(PL, '22bar', 1e-7, common[3]).VelocityKurtosis(1e-10)
    print(name, 'one-loop tables done %.0fs' % (time.time() - t0), flush=True)
    jobs = [('ML', (p1, p2, m, n, which)) for (p1, p2, m, n) in MLSPEC for which in ('linear', 'cross', 'oneloop')]
    out['ML'] = scatter(pool, name, jobs, kint, extra=(common,)).reshape(7, 3, len(kint))
    print(name, 'ImnOneLoop done %.0fs' % (time.time() - t0), flush=True)
    path = os.path.join(HERE, 'tight_dense_%s.npz' % name)
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f kB' % (os.path.getsize(path)/1e3), flush=True)


if __name__ == '__main__':
    names = sys.argv[1:] or ['golden', 'bench0', 'bench1', 'benchmax', 'benchmin']
    with mp.get_context('fork').Pool(int(os.environ.get('NPROC', '8'))) as pool:
        for nm in names:
            run(nm, pool)
