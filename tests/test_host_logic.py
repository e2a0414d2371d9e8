"""Host logic shared with the CUDA kernels (knot grid, panel generator, kernel formulae, table
reads, product-integration weights) exercised on the CPU through tests/host_emul and checked
against the reference C++ at tight tolerance (tests/golden/oracle_tight.npz).  CPU only."""
import numpy as np
import pytest

KID = dict(I=lambda m, n: 4*m + n)
K_IDS = [16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28]


@pytest.fixture(scope='module')
def table(emul, golden, tight):
    h = float(golden['cosmo_h'])
    args = (float(golden['cosmo_n_s']), float(tight['delta_H'])**2, h, float(tight['Omega0_m']),
            float(golden['cosmo_Omega_b']), float(tight['Tcmb']))
    return emul.plin_table(golden['cosmo_k'], golden['cosmo_T'], *args), args


def test_knot_grid(emul):
    x = emul.knots()
    assert len(x) == 2457
    d = np.diff(x)
    assert np.all(d >= 0) and np.sum(d == 0) == 5          # zone edges are duplicated, nothing else
    for edge in (1e-5, 10., 100.):
        assert np.min(np.abs(x - np.log(edge))) < 1e-12


def test_plin_matches_reference_formula(emul, ref_plin, golden, table):
    _, args = table
    q = np.exp(np.random.default_rng(0).uniform(np.log(1e-8), np.log(1e5), 3000))
    out = emul.plin_eval(golden['cosmo_k'], golden['cosmo_T'], *args, q)
    assert np.max(np.abs(out/ref_plin(q) - 1)) < 1e-13


def test_table_read_accuracy(emul, ref_plin, table):
    tab, _ = table
    q = np.exp(np.random.default_rng(1).uniform(np.log(1e-7), np.log(15.), 4000))
    err = np.abs(emul.table_read(tab, q)/ref_plin(q) - 1)
    assert err.max() < 2e-6 and np.median(err) < 5e-8


def test_netlib_spline_matches_reference(emul, oracle):
    x = np.exp(np.linspace(np.log(1e-5), np.log(10.), 1000))
    y = np.sin(3*np.log(x))*x**-0.7
    xq = np.exp(np.random.default_rng(2).uniform(np.log(1e-5), np.log(10.), 2000))
    assert np.max(np.abs(emul.spline(x, y, xq) - oracle.CubicSpline(x, y, xq))) < 1e-11*np.max(np.abs(y))
    assert emul.spline(x, y, np.array([0.5e-5, 11.]))[0] == 0.0      # zero outside the domain (Spline.cpp:362-367)


@pytest.mark.parametrize('idx', [0, 5, 7, 10, 12, 15])
def test_imn_fixed_rule_vs_tight_reference(emul, tight, table, idx):
    tab, _ = table
    out, nodes = emul.ik(idx, tight['k'], tab, tab, 0)
    err = np.abs(out - tight['I'][idx])/np.maximum(np.abs(tight['I'][idx]), tight['plin_k'])
    assert err.max() < 1e-5, err
    assert nodes/len(tight['k']) < 2e4


@pytest.mark.parametrize('j', [0, 2, 6, 8, 10, 12])
def test_kmn_fixed_rule_vs_tight_reference(emul, tight, table, j):
    tab, _ = table
    out, _ = emul.ik(K_IDS[j], tight['k'], tab, tab, 1)
    err = np.abs(out - tight['K'][j])/np.maximum(np.abs(tight['K'][j]), tight['plin_k'])
    assert err.max() < 1e-5, err


def test_jmn_product_integration_vs_tight_reference(emul, tight, table):
    tab, _ = table
    sel = tight['k'] <= 1.0
    for j, sub in enumerate([0, 1, 2, 3, 4, 6]):
        out = emul.linear(0, sub, tight['k'][sel], 1e-5, 1e5, tab)
        assert np.max(np.abs(out/tight['J'][j][sel] - 1)) < 1e-5


def test_one_dimensional_integrals_vs_tight_reference(emul, tight, table):
    tab, _ = table
    assert abs(emul.linear(1, 0, [0.], 1e-5, 100., tab)[0]/float(tight['sigma_v2']) - 1) < 1e-7
    r = tight['r'][::3]
    X = emul.linear(2, 0, r, 1e-5, 100., tab)
    Y = emul.linear(3, 0, r, 1e-5, 100., tab)
    scale = 2*float(tight['sigma_v2'])
    assert np.max(np.abs(X - tight['X_Zel'][::3])) < 1e-7*scale
    assert np.max(np.abs(Y - tight['Y_Zel'][::3])) < 1e-7*scale
    q3 = emul.linear(4, 0, [0.], 1e-5, 100., tab*tab, scale=1/(10*np.pi**2))[0]
    assert abs(q3/float(tight['Q3_at_k1']) - 1) < 1e-7
    sig = np.sqrt(emul.linear(5, 0, tight['Sigma_R'], 1e-5, 100., tab))
    assert np.max(np.abs(sig/tight['Sigma'] - 1)) < 1e-7


def test_not_a_knot_spline_matches_scipy():
    import ctypes as C
    from scipy.interpolate import InterpolatedUnivariateSpline
    from tests.host_emul import emul_lib
    L = emul_lib.load().L
    dp = emul_lib.dp
    L.emul_nak_spline.argtypes = [C.c_int, dp, dp, C.c_int, dp, dp]
    for n in (20, 200):
        x = np.logspace(-3, np.log10(0.5), n)
        y = np.cos(25*x)*x**-1.1 + 3
        xq = np.exp(np.random.default_rng(3).uniform(np.log(x[0]), np.log(x[-1]), 1000))
        out = np.empty_like(xq)
        L.emul_nak_spline(n, x, y, len(xq), xq, out)
        assert np.max(np.abs(out/InterpolatedUnivariateSpline(x, y)(xq) - 1)) < 1e-12


def test_banded_spline_operator_claim():
    """on FIXED abscissae the second derivatives of an interpolating cubic spline are a linear map
    of the ordinates whose entries decay like 0.268^|i - j|, so a band of 64 columns is exact to rounding.  Checked
    here with SciPy's not-a-knot spline on the model grid and on the 250-point PT grid (k_pt_resample, k_gal_moments,
    k_spline_build_band use the same construction with the library's own spline builders)."""
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(11)
    for x in (np.logspace(-3, np.log10(0.5), 200), np.logspace(np.log10(5e-6), 0., 250)):
        n = len(x)
        op = np.empty((n, n))
        for j in range(n):
            e = np.zeros(n)
            e[j] = 1.
            op[:, j] = CubicSpline(x, e, bc_type='not-a-knot')(x, 2)
        band = 64
        lo = np.clip(np.arange(n) - band//2, 0, n - band)
        y = rng.normal(size=n)*np.exp(rng.normal(size=n))
        full = CubicSpline(x, y, bc_type='not-a-knot')(x, 2)
        banded = np.array([op[i, lo[i]:lo[i] + band] @ y[lo[i]:lo[i] + band] for i in range(n)])
        scale = np.max(np.abs(op), axis=1)*np.max(np.abs(y))
        assert np.max(np.abs(banded - full)/scale) < 1e-13
        # the decay itself: 40 knots away the influence is below 1e-17 of the diagonal
        i = n//2
        assert abs(op[i, i + 40]) < 1e-17*abs(op[i, i])


def test_thinned_k_lists_and_fill_maps(emul, tight):
    """the thinned k lists of PTPlan and their 8-point Lagrange fill maps (csrc/quad.cpp: thin_k_list, thin_graded,
    lagrange_fill_map): kept points pass through exactly, the weights reproduce polynomials of degree 7 in ln k, the Lebesgue
    constant stays small (it was 21 at the junction of an ungraded list: the point-to-point quadrature noise of the computed
    values, 4e-7, became 9e-6), and smooth mode-coupling-like functions are filled to < 1e-6"""
    k_interp = np.logspace(np.log10(5e-6), 0., 250)
    k_1loop = np.exp(np.linspace(np.log(1e-5), np.log(10.), 1000))
    for which, k, expect in ((0, k_interp, (110, 130)), (1, k_1loop, (400, 440)), (2, k_1loop, (520, 580))):
        keep, idx, w, nsub = emul.fill_map(which, k)
        assert keep.sum() == nsub and expect[0] <= nsub <= expect[1], (which, nsub)
        assert keep[0] and keep[-1]
        if which == 0:
            assert keep[k >= 0.01].all() and not keep[k < 0.01].all()
        else:
            assert keep[(k >= 0.01) & (k <= 1.2)].all()
        pos = np.where(keep)[0]
        # kept points: identity
        assert np.array_equal(idx[keep][:, 0], np.arange(nsub)) and np.all(w[keep][:, 0] == 1.) and np.all(w[keep][:, 1:] == 0.)
        x = np.log(k)
        xs = x[pos]
        leb = np.abs(w).sum(axis=1)
        inner = (k > 1e-4) & (k < 9.)
        assert leb[inner].max() < 4.0, (which, leb[inner].max())
        assert leb.max() < 8.0
        assert np.allclose(w.sum(axis=1), 1., atol=1e-11)
        for deg in (1, 3, 7):
            f = (xs/5.)**deg
            assert np.max(np.abs((w*f[idx]).sum(axis=1) - (x/5.)**deg)) < 1e-9, (which, deg)
        # a smooth function with the shape of a low-k mode-coupling integral (power law x slow logarithmic running)
        f = lambda q: q**3.2*(1. + 0.3*np.log(q) + 0.02*np.log(q)**2)/(1. + (q/0.05)**2)
        filled = (w*f(k[pos])[idx]).sum(axis=1)
        scale = np.maximum(np.abs(f(k)), 1e-3*k**0.96)
        assert np.max(np.abs(filled - f(k))/scale) < 1e-6, which
    # short, unsorted or non-uniform lists are not thinned
    for kk in (np.logspace(-5, 0, 40), np.logspace(-5, 0, 200)[::-1].copy(), np.sort(np.random.default_rng(1).uniform(1e-5, 1., 200))):
        keep, idx, w, nsub = emul.fill_map(0, kk)
        assert nsub == len(kk) and keep.all()


def test_eh_nowiggle_host_formula_vs_reference_extrapolation(ref_cosmo, golden):
    """``pygcl.eh_nowiggle_transfer`` (the table of ``Cosmology.clone(tf=EH_NoWiggle)``): outside the tabulated k the
    reference's transfer function is its no-wiggle fit scaled to the table's end (Cosmology.cpp:187-199), so ratios of the
    reference transfer there are ratios of the fit"""
    from pyrsd_b200 import pygcl
    a = ref_cosmo.args
    h, Om, Ob, Tcmb = a[4], a[5], a[6], a[7]
    for ks in (np.array([25., 40., 90., 300.]), np.array([1e-6, 3e-6, 8e-6])):
        ref = ref_cosmo.transfer(ks)
        mine = pygcl.eh_nowiggle_transfer(ks, h, Om, Ob, Tcmb)
        assert np.max(np.abs((mine/mine[0])/(ref/ref[0]) - 1)) < 1e-13


def test_quasar_growth_function_einstein_de_sitter():
    """qso.growth_function (power_qso.py:32-46): E(a) int_0^a da' / (a' E)^3 = a / 2.5 for Omega_m = 1"""
    from pyrsd_b200.qso import growth_function
    eds = {'Omega0_m': 1.0, 'Omega0_lambda': 0.0, 'Omega0_k': 0.0}
    for z in (0., 1., 100.):
        assert abs(growth_function(eds, z)*2.5*(1 + z) - 1) < 1e-8
