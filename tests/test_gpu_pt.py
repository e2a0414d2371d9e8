"""GPU parity of the PT integrals: CUDA path (through the C-ABI) vs the reference C++ at tight
tolerance (tests/golden/oracle_tight.npz) and vs the reference's shipped default-eps tables.

Tolerance (BASELINE.json north_star): 1e-5 relative on every term; for I / K the reference stops on
an ABSOLUTE criterion eps*P_L(k)/V (Imn.cpp:129), so the error is measured relative to
max(|value|, P_L(k))."""
import numpy as np
import pytest

from tests.gpu_common import golden_pygcl_cosmology, relerr_floor

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope='module')
def cosmo(golden):
    return golden_pygcl_cosmology(golden)


@pytest.fixture(scope='module')
def plin(cosmo):
    from pyrsd_b200 import pygcl
    return pygcl.LinearPS(cosmo, 0.)


def test_delta_H_and_linear_power(cosmo, plin, tight, golden):
    assert abs(cosmo.delta_H()/float(tight['delta_H']) - 1) < 1e-6
    out = plin(tight['k'])
    assert np.max(np.abs(out/tight['plin_k'] - 1)) < 1e-6       # amplitude carries delta_H^2
    assert isinstance(plin(0.1), float)
    assert plin(np.array([])).shape == (0,)


def test_imn_all_sixteen(plin, tight):
    from pyrsd_b200 import pygcl
    I = pygcl.Imn(plin)
    for m in range(4):
        for n in range(4):
            err = relerr_floor(I(tight['k'], m, n), tight['I'][4*m + n], tight['plin_k'])
            assert err.max() < TOL, (m, n, err.max())


def test_kmn_all_thirteen(plin, tight):
    from pyrsd_b200 import pygcl
    K = pygcl.Kmn(plin)
    spec = [(0, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 1, 1, 0), (0, 2, 1, 0), (1, 0, 0, 0), (1, 1, 0, 0),
            (1, 0, 1, 0), (1, 1, 1, 0), (2, 0, 0, 0), (2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 1, 1)]
    for j, (m, n, t, p) in enumerate(spec):
        err = relerr_floor(K(tight['k'], m, n, bool(t), p), tight['K'][j], tight['plin_k'])
        assert err.max() < TOL, (m, n, t, p, err.max())


def test_jmn_all_six(plin, tight):
    from pyrsd_b200 import pygcl
    J = pygcl.Jmn(plin)
    sel = tight['k'] <= 1.0      # k_interp range (J02/J10/J20 are only tabulated there)
    for j, (m, n) in enumerate([(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]):
        k = tight['k'] if (m, n) in ((0, 0), (0, 1), (1, 1)) else tight['k'][sel]
        ref = tight['J'][j] if len(k) == len(tight['k']) else tight['J'][j][sel]
        assert np.max(np.abs(J(k, m, n)/ref - 1)) < TOL, (m, n)


def test_scalar_zero_and_invalid_arguments(plin):
    from pyrsd_b200 import pygcl, _native
    I = pygcl.Imn(plin)
    assert I(0.0, 0, 0) == 0.0 and I(-1.0, 2, 2) == 0.0          # Imn.cpp:122: k <= 0 -> 0
    assert I(np.array([]), 0, 0).shape == (0,)
    with pytest.raises(_native.NativeError):
        I(0.1, 4, 0)                                           # the reference abort()s here (Imn.cpp:186)
    with pytest.raises(_native.NativeError):
        pygcl.Jmn(plin)(0.1, 1, 2)
    with pytest.raises(_native.NativeError):
        pygcl.Kmn(plin)(0.1, 2, 1)


def test_one_dimensional_integrals(plin, tight):
    assert abs(plin.VelocityDispersion()/float(tight['sigma_v2']) - 1) < 1e-6
    assert np.max(np.abs(plin.VelocityDispersion(tight['k'], 0.5)/tight['sigmasq_k'] - 1)) < TOL
    scale = 2*float(tight['sigma_v2'])
    assert np.max(np.abs(plin.X_Zel(tight['r']) - tight['X_Zel'])) < 1e-6*scale
    assert np.max(np.abs(plin.Y_Zel(tight['r']) - tight['Y_Zel'])) < 1e-6*scale
    assert abs(plin.Q3_Zel(1.0)/float(tight['Q3_at_k1']) - 1) < TOL
    assert np.max(np.abs(plin.Sigma(tight['Sigma_R'])/tight['Sigma'] - 1)) < TOL


def test_oneloop_tables_vs_tight(plin, tight, golden):
    from pyrsd_b200 import pygcl
    objs = [pygcl.OneLoopPdd(plin), pygcl.OneLoopPdv(plin), pygcl.OneLoopPvv(plin), pygcl.OneLoopP22Bar(plin)]
    idx = tight['oneloop_idx']
    for j, o in enumerate(objs):
        tab = o.GetOneLoopPower()
        assert tab.shape == (1000,)
        err = relerr_floor(tab[idx], tight['oneloop'][j], tight['oneloop_plin'])
        # 2 I + 6 k^2 J P_L cancels to ~7% of each term at k > 1 h/Mpc: 1e-5 is asked of the terms,
        # which leaves ~1.5e-5 x 15 on the difference there
        lowk = idx < 780                       # k < 0.5 h/Mpc
        assert err[lowk].max() < TOL, (j, err[lowk].max())
        assert err.max() < 2e-4, (j, err.max())
    # full 1000-point tables against the reference's constructors at epsrel = 1e-7
    k1 = np.exp(np.log(1e-5) + np.arange(1000)*np.log(1e6)/999)
    P1 = plin(k1)
    for j, o in enumerate(objs):
        err = relerr_floor(o.GetOneLoopPower(), tight['oneloop_common'][j], P1)
        assert err[k1 < 0.5].max() < TOL, (j, err[k1 < 0.5].max())
        assert err.max() < 2e-4, (j, err.max())
    # and against the reference's shipped default-eps tables (only as accurate as epsrel = 1e-4 allows)
    k1 = np.exp(np.log(1e-5) + np.arange(1000)*np.log(1e6)/999)
    for j, name in enumerate(['_Pdd_0', '_Pdv_0', '_Pvv_0', '_P22bar_0']):
        gold = golden['oneloop' + name]
        err = relerr_floor(objs[j].GetOneLoopPower(), gold, plin(k1))
        assert np.median(err) < 1e-6 and np.percentile(err, 90) < 5e-3


def test_oneloop_evaluate_semantics(plin, golden):
    from pyrsd_b200 import pygcl
    P = pygcl.OneLoopPdd(plin, 1e-4, golden['oneloop_Pdd_0'])
    k = np.array([0.5e-5, 1e-5, 0.01, 0.3, 10.0, 11.0])
    ev = P(k)
    assert ev[0] == 0.0 and ev[-1] == 0.0                      # OneLoopPS.cpp:28-33
    assert abs(ev[1]/golden['oneloop_Pdd_0'][0] - 1) < 1e-12 and abs(ev[4]/golden['oneloop_Pdd_0'][-1] - 1) < 1e-12
    full = P.EvaluateFull(k)
    assert full[-1] == 0.0 and abs(full[0]/plin(k[0]) - 1) < 1e-12       # OneLoopPS.cpp:21-26
    assert np.allclose(full[1:5], ev[1:5] + plin(k[1:5]), rtol=1e-13)
    state = P.__getstate__()['args']
    assert state[1] == 1e-4 and np.array_equal(state[2], golden['oneloop_Pdd_0'])


def test_imn_oneloop_vs_tight_with_common_tables(plin, tight, golden):
    """ImnOneLoop stage in isolation: both sides are fed the SAME one-loop tables (the reference's own
    OneLoopP* constructors at epsrel = 1e-7, committed in the fixture)"""
    from pyrsd_b200 import pygcl
    common = tight['oneloop_common']
    ol = {w: cls(plin, 1e-7, common[j]) for j, (w, cls) in
          enumerate((('dd', pygcl.OneLoopPdd), ('dv', pygcl.OneLoopPdv), ('vv', pygcl.OneLoopPvv)))}
    P22 = pygcl.OneLoopP22Bar(plin, 1e-7, common[3])
    assert abs(P22.VelocityKurtosis()/float(tight['kurtosis_common_tables']) - 1) < TOL
    spec = [('vv', 'dd', 0, 1), ('vv', 'dd', 0, 2), ('dv', None, 0, 3), ('dv', None, 0, 4),
            ('vv', None, 2, 3), ('vv', None, 3, 2), ('vv', None, 3, 3)]
    k = tight['k_ml']
    floor = plin(k)
    for j, (p1, p2, m, n) in enumerate(spec):
        I = pygcl.ImnOneLoop(ol[p1]) if p2 is None else pygcl.ImnOneLoop(ol[p1], ol[p2], 1e-4)
        for w, fn in enumerate((I.EvaluateLinear, I.EvaluateCross, I.EvaluateOneLoop)):
            err = relerr_floor(fn(k, m, n), tight['ML'][j, w], floor)
            assert err.max() < TOL, (p1, p2, m, n, w, err.max())
    # the reference's shipped default-eps tables carry ~1e-4 point-to-point noise that its 1000-point spline
    # interpolates; the tabulated read follows that noise only to ~3e-5
    noisy = {w: cls(plin, 1e-4, golden['oneloop_P%s_0' % w]) for w, cls in (('dd', pygcl.OneLoopPdd), ('vv', pygcl.OneLoopPvv))}
    out = pygcl.ImnOneLoop(noisy['vv'], noisy['dd'], 1e-4).EvaluateLinear(k, 0, 1)
    assert relerr_floor(out, golden['pt_Ivvdd_h01'][0][[60, 130, 170, 195, 215, 235, 249]], floor).max() < 1e-4


def test_plan_matches_one_off_calls_and_golden(plin, cosmo, golden, tight):
    """the batched plan (what a fit uses) against the one-off pygcl calls and the shipped tables"""
    from pyrsd_b200 import pygcl, rsd
    eng = rsd.PTEngine()
    sp = eng.spectra(cosmo.GetDiscreteK(), cosmo.GetDiscreteTk(), cosmo.n_s(), cosmo.delta_H()**2, cosmo.h(),
                     cosmo.Omega0_m(), cosmo.Omega0_b(), cosmo.Tcmb())
    res = eng.run(sp)
    I, K, J = eng.get(res, 'I')[0], eng.get(res, 'K')[0], eng.get(res, 'J')[0]
    k = eng.k_interp
    sel = np.arange(30, 250, 31)
    Pk = plin(k)
    # below k = 0.01 the plan integrates every 6th point of a long k list and fills the others by 8-point Lagrange
    # interpolation in ln k (plan.cu, PTPlan ctor); a one-off call with a short list integrates every k it is given
    direct = k[sel] >= 0.01
    for tab, ref in ((I[10], pygcl.Imn(plin)(k[sel], 2, 2)), (K[6], pygcl.Kmn(plin)(k[sel], 1, 1))):
        # same nodes and tables; the operator entries are rounded to binary32 per column, and the plan's column of a
        # non-symmetric kernel is f(u, v) + f(u, -v) where the one-off call sums two separately rounded operators
        assert direct.sum() >= 3 and np.max(np.abs(tab[sel] - ref)[direct]) <= 1e-7*np.max(np.abs(tab))
        assert relerr_floor(tab[sel], ref, Pk[sel]).max() < 1e-6
    # ImnOneLoop comes from the one-lane-per-node kernel in the plan and from the DMMA operator in the one-off pygcl
    # calls: same nodes, different kernels and summation orders
    ML = eng.get(res, 'ML')[0]
    ol = {w: cls(plin) for w, cls in (('dd', pygcl.OneLoopPdd), ('dv', pygcl.OneLoopPdv), ('vv', pygcl.OneLoopPvv))}
    spec = [('vv', 'dd', 0, 1), ('vv', 'dd', 0, 2), ('dv', None, 0, 3), ('dv', None, 0, 4),
            ('vv', None, 2, 3), ('vv', None, 3, 2), ('vv', None, 3, 3)]
    sel_ml = np.array([60, 130, 170, 195, 215, 235, 249])
    for j, (p1, p2, m, n) in enumerate(spec):
        Iml = pygcl.ImnOneLoop(ol[p1]) if p2 is None else pygcl.ImnOneLoop(ol[p1], ol[p2], 1e-4)
        for w, fn in enumerate((Iml.EvaluateLinear, Iml.EvaluateCross, Iml.EvaluateOneLoop)):
            err = relerr_floor(ML[j, w][sel_ml], fn(k[sel_ml], m, n), Pk[sel_ml])
            assert err.max() < 1e-7, (j, w, err.max())      # the nodal operator streams its entries as binary32 (bilinear.cuh)
    # shipped tables were computed at the reference's default epsrel: medians agree to ~1e-9, the
    # worst points differ by what epsrel = 1e-4 / 1e-3 allows
    for name, arr in [('I02', I[2]), ('I22', I[10]), ('I31', I[13]), ('K00', K[0]), ('K11', K[6]), ('K20s_b', K[12]),
                      ('J02', J[2]), ('J10', J[3])]:
        err = relerr_floor(arr, golden['pt_' + name], Pk)
        assert np.median(err) < 1e-7, name
        assert err.max() < 5e-2, name
    sc = eng.get(res, 'scalars')[0]
    assert abs(sc[0]/float(tight['sigma_v2']) - 1) < 1e-6
    assert abs(np.sqrt(sc[0])/float(golden['sigma_lin_unnormed']) - 1) < 1e-6
    assert abs(sc[1]/float(golden['velocity_kurtosis_unnormed']) - 1) < 1e-3     # reference: epsrel 1e-4 tables + 1e-4 integral
    X0, YY = eng.get(res, 'X0')[0], eng.get(res, 'YY')[0]
    assert np.max(np.abs(X0 - golden['zel_X0'])) < 1e-3 and np.max(np.abs(YY - golden['zel_YY'])) < 2e-3


def test_batch_consistency_and_scaling_property(cosmo, golden):
    """size-independent properties: I, K ~ amplitude^2, J ~ amplitude, ragged batch (9 = 8 + 1)"""
    from pyrsd_b200 import rsd
    k_interp = rsd.k_interp_default()[::10]
    eng = rsd.PTEngine(k_interp=k_interp)
    nc = 9
    amp = cosmo.delta_H()**2*np.linspace(0.6, 1.4, nc)
    sp = eng.spectra(cosmo.GetDiscreteK(), np.tile(cosmo.GetDiscreteTk(), (nc, 1)), cosmo.n_s(), amp, cosmo.h(),
                     cosmo.Omega0_m(), cosmo.Omega0_b(), cosmo.Tcmb())
    res = eng.run(sp)
    I, J, ML = eng.get(res, 'I'), eng.get(res, 'J'), eng.get(res, 'ML')
    r = amp/amp[4]
    # measured against max(|value|, P_L(k)) as everywhere for I: at the lowest k the kernels cancel by
    # ~1e7 between the two halves of the domain, so the bare ratio only scales to ~1e-9
    from pyrsd_b200 import pygcl
    floor = pygcl.LinearPS(cosmo, 0.)(k_interp)[None, None, :]
    assert relerr_floor(I, r[:, None, None]**2*I[4][None], floor).max() < 1e-12
    assert np.max(np.abs(J/(r[:, None, None]*J[4][None]) - 1)) < 1e-12
    # ImnOneLoop: linear ~ A^2, cross ~ A^3, one-loop ~ A^4
    for part, power in ((0, 2), (1, 3), (2, 4)):
        err = relerr_floor(ML[:, :, part, :], r[:, None, None]**power*ML[4][None, :, part, :], floor)
        assert err.max() < 1e-9


def test_variant_cosmology_vs_tight_reference(golden):
    """the fixed rule was tuned on the golden Planck15 spectrum: the same parity on a tilted one (delta = +0.05,
    n_s = 0.93; tests/golden/make_tight_variant.py)"""
    import os
    from pyrsd_b200 import pygcl
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'oracle_tight_variant.npz')
    if not os.path.exists(path):
        pytest.skip("tests/golden/oracle_tight_variant.npz not generated")
    v = np.load(path)
    params = {k[len('cosmo_'):]: float(golden[k]) for k in golden.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['T_cmb'] = 2.7255
    params['n_s'] = float(v['n_s'])
    k_i = golden['cosmo_k']
    T_i = golden['cosmo_T']*(k_i/0.05)**(0.5*float(v['delta']))
    cosmo = pygcl.Cosmology(params, 0, float(golden['cosmo_sigma8']), k_i, T_i)
    assert abs(cosmo.delta_H()/float(v['delta_H']) - 1) < 1e-6
    plin = pygcl.LinearPS(cosmo, 0.)
    k, floor = v['k'], v['plin_k']
    assert np.max(np.abs(plin(k)/floor - 1)) < 1e-6
    I, K, J = pygcl.Imn(plin), pygcl.Kmn(plin), pygcl.Jmn(plin)
    for m in range(4):
        for n in range(4):
            err = relerr_floor(I(k, m, n), v['I'][4*m + n], floor)
            assert err.max() < TOL, ('I', m, n, err.max())
    spec = [(0, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 1, 1, 0), (0, 2, 1, 0), (1, 0, 0, 0), (1, 1, 0, 0),
            (1, 0, 1, 0), (1, 1, 1, 0), (2, 0, 0, 0), (2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 1, 1)]
    for j, (m, n, t, p) in enumerate(spec):
        err = relerr_floor(K(k, m, n, bool(t), p), v['K'][j], floor)
        assert err.max() < TOL, ('K', m, n, t, p, err.max())
    sel = k <= 1.0
    for j, (m, n) in enumerate([(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]):
        kk = k if (m, n) in ((0, 0), (0, 1), (1, 1)) else k[sel]
        ref = v['J'][j] if len(kk) == len(k) else v['J'][j][sel]
        assert np.max(np.abs(J(kk, m, n)/ref - 1)) < TOL, ('J', m, n)
    idx = v['oneloop_idx']
    for j, o in enumerate([pygcl.OneLoopPdd(plin), pygcl.OneLoopPdv(plin), pygcl.OneLoopPvv(plin), pygcl.OneLoopP22Bar(plin)]):
        err = relerr_floor(o.GetOneLoopPower()[idx], v['oneloop'][j], v['oneloop_plin'])
        lowk = idx < 780
        assert err[lowk].max() < TOL, ('one-loop', j, err[lowk].max())
        assert err.max() < 2e-4, ('one-loop', j, err.max())
    assert abs(plin.VelocityDispersion()/float(v['sigma_v2']) - 1) < TOL       # sigma_v^2 (PowerSpectrum.cpp:56-58)
    if 'ML' in v.files:
        common = v['oneloop_common']
        ol = {w: cls(plin, 1e-7, common[j]) for j, (w, cls) in
              enumerate((('dd', pygcl.OneLoopPdd), ('dv', pygcl.OneLoopPdv), ('vv', pygcl.OneLoopPvv)))}
        P22 = pygcl.OneLoopP22Bar(plin, 1e-7, common[3])
        assert abs(P22.VelocityKurtosis()/float(v['kurtosis_common_tables']) - 1) < TOL
        mspec = [('vv', 'dd', 0, 1), ('vv', 'dd', 0, 2), ('dv', None, 0, 3), ('dv', None, 0, 4),
                 ('vv', None, 2, 3), ('vv', None, 3, 2), ('vv', None, 3, 3)]
        kml = v['k_ml']
        fl = plin(kml)
        for j, (p1, p2, m, n) in enumerate(mspec):
            Iml = pygcl.ImnOneLoop(ol[p1]) if p2 is None else pygcl.ImnOneLoop(ol[p1], ol[p2], 1e-4)
            for w, fn in enumerate((Iml.EvaluateLinear, Iml.EvaluateCross, Iml.EvaluateOneLoop)):
                err = relerr_floor(fn(kml, m, n), v['ML'][j, w], fl)
                assert err.max() < TOL, ('ML', p1, p2, m, n, w, err.max())
        # the computed one-loop tables against the tight common ones, whole grid
        k1 = np.exp(np.log(1e-5) + np.arange(1000)*np.log(1e6)/999)
        P1 = plin(k1)
        for j, o in enumerate([pygcl.OneLoopPdd(plin), pygcl.OneLoopPdv(plin), pygcl.OneLoopPvv(plin), pygcl.OneLoopP22Bar(plin)]):
            err = relerr_floor(o.GetOneLoopPower(), common[j], P1)
            assert err[k1 < 0.5].max() < TOL, ('one-loop table', j, err[k1 < 0.5].max())


def test_pygcl_small_surface(golden, tmp_path):
    """Cosmology.from_power / clone(tf=EH_NoWiggle) / NormalizeTransferFunction, PowerSpectrum.Save, ImnOneLoop getters"""
    from pyrsd_b200 import pygcl
    from tests.gpu_common import golden_pygcl_cosmology
    cosmo = golden_pygcl_cosmology(golden)
    P = pygcl.LinearPS(cosmo, 0.)
    # a cosmology read off its own linear spectrum reproduces it (the transfer function is defined up to its amplitude)
    k = cosmo.GetDiscreteK()
    c2 = pygcl.Cosmology.from_power(cosmo.GetParams(), k, P(k), sigma8=cosmo.sigma8())
    assert abs(c2.GetDiscreteTk().max() - 1e7) < 1e-3
    kq = np.logspace(-3, 0, 30)
    assert np.max(np.abs(pygcl.LinearPS(c2, 0.)(kq)/P(kq) - 1)) < 1e-9
    with pytest.raises(ValueError):
        pygcl.Cosmology.from_power(cosmo.GetParams(), k, P(k))
    # no-wiggle clone: same sigma8, smooth spectrum within the BAO amplitude of the full one
    nw = cosmo.clone(tf=pygcl.transfers.EH_NoWiggle)
    Pnw = pygcl.LinearPS(nw, 0.)
    assert abs(Pnw.Sigma(8.)/cosmo.sigma8() - 1) < 1e-6
    assert np.max(np.abs(Pnw(kq)/P(kq) - 1)) < 0.15
    assert np.allclose(nw.GetDiscreteTk(), cosmo.GetNoWiggleTransfer(k), rtol=1e-15)
    c3 = cosmo.clone()
    c3.NormalizeTransferFunction(0.5)
    assert abs(pygcl.LinearPS(c3, 0.).Sigma(8.) - 0.5) < 1e-6
    path = str(tmp_path / 'plin.dat')
    P.Save(path, 1e-3, 1., 11, True)
    tab = np.loadtxt(path)
    assert tab.shape == (11, 2) and np.allclose(tab[:, 1], P(tab[:, 0]), rtol=1e-5)
    Pdd, Pvv = pygcl.OneLoopPdd(P), pygcl.OneLoopPvv(P)
    I = pygcl.ImnOneLoop(Pvv, Pdd, 1e-4)
    assert I.GetOneLoopPS1() is Pvv and I.GetOneLoopPS2() is Pdd and not I.GetEqual()
    assert pygcl.ImnOneLoop(Pdd).GetEqual() and Pdd.GetCosmology() is cosmo
    assert pygcl.CubicSpline(k[:50], P(k[:50])).Evaluate(k[10]) == P(k[10])
