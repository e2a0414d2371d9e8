"""Generate ``tests/golden/oracle_tight_variant.npz``: the reference C++ (oracle/_ref) at TIGHT ``epsrel`` on a
transfer function that is NOT the one the fixed quadrature rule was tuned on: the golden Planck15 T(k) tilted by
(k / 0.05)^(delta / 2) with delta = +0.05 (2.5 sigma of the synthetic batches of bench.py) and n_s = 0.93.  Guards
against a rule over-fitted to one spectrum (tests/test_gpu_pt.py::test_variant_cosmology_vs_tight_reference).

Run in the build container only:   python tests/golden/make_tight_variant.py   (a few minutes)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import gcl_ref as R  # noqa: E402

DELTA, NS = 0.05, 0.93


def variant_inputs(g):
    k = np.asarray(g['cosmo_k'], dtype=float)
    T = np.asarray(g['cosmo_T'], dtype=float)*(k/0.05)**(0.5*DELTA)
    h = float(g['cosmo_h'])
    Om = float(g['cosmo_Omega_b']) + float(g['cosmo_Omega_cdm']) + float(g['cosmo_m_ncdm'])/93.14/h**2
    return k, T, float(g['cosmo_sigma8']), NS, h, Om, float(g['cosmo_Omega_b']), 2.7255


def main():
    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    k_i, T_i, s8, ns, h, Om, Ob, Tcmb = variant_inputs(g)
    cosmo = R.RefCosmology(k_i, T_i, s8, ns, h, Om, Ob, Tcmb)
    PL = R.RefLinearPS(cosmo)
    kint = g['k_interp']
    k = np.concatenate([kint[[20, 80, 120, 150, 175, 195, 210, 225, 240, 249]], [2.5]])
    out = {'delta': DELTA, 'n_s': NS, 'delta_H': cosmo.delta_H(), 'k': k, 'plin_k': PL(k)}
    t0 = time.time()
    I = np.empty((16, len(k)))
    for m in range(4):
        for n in range(4):
            I[4*m + n] = R.Imn(PL, k, m, n, epsrel=1e-9)
    out['I'] = I
    print('I done %.0fs' % (time.time() - t0), flush=True)
    J = np.empty((6, len(k)))
    for j, (m, n) in enumerate([(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]):
        J[j] = R.Jmn(PL, k, m, n, epsrel=1e-10)
    out['J'] = J
    Kspec = [(0, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 1, 1, 0), (0, 2, 1, 0), (1, 0, 0, 0), (1, 1, 0, 0),
             (1, 0, 1, 0), (1, 1, 1, 0), (2, 0, 0, 0), (2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 1, 1)]
    K = np.empty((13, len(k)))
    for j, (m, n, tid, part) in enumerate(Kspec):
        K[j] = R.Kmn(PL, k, m, n, bool(tid), part, epsrel=1e-9)
    out['K'] = K
    print('K done %.0fs' % (time.time() - t0), flush=True)
    # one-loop spectra (OneLoopPS.cpp:56-185) at points of their own 1000-point grid
    k1 = R.logspace(1e-5, 10, 1000)
    idx = np.array([120, 400, 560, 640, 700, 740, 770, 800, 830, 870, 930])
    kk = k1[idx]
    P1 = PL(kk)
    ol = np.empty((4, len(idx)))
    for w, (m, n) in enumerate([(0, 0), (0, 1), (1, 1)]):
        ol[w] = 2*R.Imn(PL, kk, m, n, epsrel=1e-8) + 6*kk*kk*R.Jmn(PL, kk, m, n, epsrel=1e-10)*P1
    ol[3] = R.Imn(PL, kk, 2, 3, epsrel=1e-8) + (2./3.)*R.Imn(PL, kk, 3, 2, epsrel=1e-8) + (1./5.)*R.Imn(PL, kk, 3, 3, epsrel=1e-8)
    out['oneloop_idx'], out['oneloop'], out['oneloop_plin'] = idx, ol, P1
    print('one-loop done %.0fs' % (time.time() - t0), flush=True)
    out['sigma_v2'] = PL.sigv_tol()
    # ImnOneLoop on common tight one-loop tables (same recipe as make_tight.py)
    sys.path.insert(0, HERE)
    from make_tight import ml_part
    ml_part(out, R, PL, kint, t0)
    path = os.path.join(HERE, 'oracle_tight_variant.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f kB' % (os.path.getsize(path)/1e3))


if __name__ == '__main__':
    main()
