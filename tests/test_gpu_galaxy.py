"""GPU parity of the GalaxySpectrum P(k, mu) assembly against the UNMODIFIED reference Python run on
the reference C++ at tight tolerances (tests/golden/galaxy_ref.npz, tests/golden/make_galaxy_ref.py),
plus size-independent properties at the BASELINE C2 grid (1000 k x 40 mu)."""
import os

import numpy as np
import pytest

from tests.gpu_common import golden_pygcl_cosmology, b2_00, b2_01, stochasticity

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, 'golden', 'galaxy_ref.npz')
TOL = 1e-5

CASES = {
    'z0_default': (0.0, dict(), dict()),
    'z055_ap': (0.55, dict(), dict()),
    'z055_spt': (0.55, dict(use_P00_model=False, use_P01_model=False, use_P11_model=False, use_Pdv_model=False,
                            use_Pvv_model=False, use_Phm_model=False), dict()),
    'z055_so_gauss': (0.55, dict(use_so_correction=True, fog_model='gaussian', max_mu=6, use_tidal_bias=True), dict()),
    # the reference's default: Zel'dovich terms from the 100 x 300 sigma8_z x k tables (rsd/hzpt/interpolated.py); "_edge":
    # sigma8_z / D lies outside the table and that one call falls back to the direct evaluation
    'z055_interp': (0.55, dict(interpolate=True), dict()),
    'z055_interp_edge': (0.55, dict(interpolate=True), dict()),
    # simulation-calibrated options of HaloSpectrum (power_biased.py:26-52, 219-286, 462-520): the fits behind them are
    # the smooth stand-ins of pyrsd_b200/synthetic.py on both sides
    'z055_simopts': (0.55, dict(use_mean_bias=True, vel_disp_from_sims=True, correct_mu2=True, correct_mu4=True), dict()),
    'z055_veldisp': (0.55, dict(vel_disp_from_sims=True, correct_mu4=True), dict()),
    # user-loaded dark-matter terms (SimLoaderMixin._load, rsd/sim_loader.py:56-66)
    'z055_loaded': (0.55, dict(), dict(load=True)),
}
INTERP_CASES = ('z055_interp', 'z055_interp_edge')
PI_ONLY = INTERP_CASES + ('z055_simopts', 'z055_veldisp', 'z055_loaded')


@pytest.fixture(scope='module')
def cosmo(golden):
    return golden_pygcl_cosmology(golden)


@pytest.fixture(scope='module')
def engine():
    from pyrsd_b200 import rsd
    return rsd.PTEngine()


def make_model(cosmo, engine, z, **kw):
    from pyrsd_b200 import rsd
    kw.setdefault('interpolate', False)          # the fixtures' cases state interpolate=False (make_galaxy_ref.py) unless a test says otherwise
    from pyrsd_b200 import synthetic as S
    return rsd.GalaxySpectrum(params=cosmo, z=z, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity, engine=engine,
                              vel_disp_fitter=S.vel_disp, Pmu2_correction=S.Pmu2_corr, Pmu4_correction=S.Pmu4_corr, **kw)


# k0_low <= k' < 0.0075: the Zel'dovich transform amplifies errors of 1e-9 sigma_v^2 in Y(r > 1e4 Mpc/h) to 1e-4 in
# P00.  The reference's own adaptive X_Zel / Y_Zel do not converge there even at epsrel = 1e-10 (Y off by 1.5e-4 at
# r = 3.7e4 against brute-force quadrature, which the library's product integration matches to 1e-13 sigma_v^2;
# tests/golden/make_galaxy_ref.py ZEL_SOURCE, profiles/r02_parity.md).  So: against the reference fed with ITS OWN
# state the bound in that band is the reference's noise (2.5e-4), 1e-5 elsewhere; against the reference fed with the
# converged state ("_pi" cases) it is 1e-5 everywhere.
ZEL_BAND = (4.5e-3, 8e-3)       # in OBSERVED k: the AP cases map k to k' = k / alpha (3 %), and the model's spline rings one knot below k0_low
VARIANTS = [(c, '') for c in sorted(CASES) if c not in PI_ONLY] + [(c, '_pi') for c in sorted(CASES) if c != 'z055_spt']


def _record(tag, value):
    try:
        import json
        path = os.path.join(os.path.dirname(HERE), 'gpurun_out', 'r02_parity.json')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur.setdefault('galaxy', {})[tag] = value
        json.dump(cur, open(path, 'w'), indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/galaxy_ref.npz not generated")
@pytest.mark.parametrize('case,variant', VARIANTS)
def test_power_vs_reference_python(cosmo, engine, case, variant):
    ref = np.load(REF)
    z, kw, extra = CASES[case]
    model = make_model(cosmo, engine, z, **kw)
    if extra.get('load'):
        from pyrsd_b200 import synthetic as S
        kl = np.logspace(-3.2, 0., 400)
        for term, Pl in S.loaded_terms(kl).items():
            model._load(term, kl, Pl)
    names = [str(n) for n in ref['par_names']]
    key = case + variant
    if key + '_P' not in ref.files:
        pytest.skip("fixture has no case %s" % key)
    pars = dict(zip(names, ref[key + '_pars']))
    pars.pop('D')
    model.update(**pars)
    assert np.allclose(model.k, ref[key + '_modelk'], rtol=1e-14)
    assert abs(model.sigma_lin/float(ref[key + '_sigma_lin']) - 1) < TOL
    # the benchmark's own output grid: every second k of bench.py's 1000 in [1.12e-3, 0.49] x 40 mu
    k, mu = ref[key + '_k'], ref['mu']
    assert len(mu) == 40 and k.min() < 1.2e-3 and k.max() > 0.475
    P = model.power(k, mu)
    assert P.shape == ref[key + '_P'].shape
    err = np.abs(P/ref[key + '_P'] - 1)
    band = (k >= ZEL_BAND[0]) & (k < ZEL_BAND[1]) if (variant == '' and case != 'z055_spt') else np.zeros(len(k), dtype=bool)
    e_out = err[~band]
    i, j = np.unravel_index(np.argmax(np.where(band[:, None], 0, err)), err.shape)
    _record('power/' + key, dict(max_err=float(e_out.max()), at_k=float(k[i]), at_mu=float(mu[j]), grid=list(P.shape),
                                 max_err_in_reference_noise_band=float(err[band].max()) if band.any() else None))
    assert e_out.max() < TOL, (key, e_out.max(), k[i], mu[j])
    if band.any():
        assert err[band].max() < 2.5e-4, (key, err[band].max())


def test_power_properties_at_baseline_grid(cosmo, engine):
    """BASELINE config C2: 1000 k x 40 mu.  No oracle at this size: identities the model must obey."""
    model = make_model(cosmo, engine, 0.55)
    k = np.logspace(-2.9, np.log10(0.49), 1000)
    mu = np.linspace(0., 1., 40)
    P = model.power(k, mu)
    assert P.shape == (1000, 40) and np.all(np.isfinite(P)) and np.all(P > 0)
    # flatten = column-major ravel (power_dm.py:1011-1012)
    assert np.array_equal(model.power(k, mu, flatten=True), np.ravel(P, order='F'))
    # paired (k, mu) evaluation equals the grid
    assert np.array_equal(model.power(k[::50], np.full(20, mu[7])), P[::50, 7])
    # isotropic AP: P_obs(k, mu) = P(k/alpha, mu)/alpha^3 (rsd/tools.py:54-75, 211)
    a = 1.02
    model.update(alpha_par=a, alpha_perp=a)
    Pa = model.power(k[100:900:40], mu)
    model.update(alpha_par=1., alpha_perp=1.)
    assert np.max(np.abs(Pa*a**3/model.power(k[100:900:40]/a, mu) - 1)) < 1e-12
    # alpha_drag rescales by its cube (tools.py:214)
    model.update(alpha_drag=1.01)
    assert np.max(np.abs(model.power(k[::100], mu)/(1.01**3*P[::100]) - 1)) < 1e-13
    model.update(alpha_drag=1.)
    # no satellites: only the centrals' terms survive; mu = 0 is undamped
    model.update(fs=0., fcB=0.)
    P0 = model.power(k[::100], np.array([0.]))
    model.update(sigma_c=7.)
    assert np.max(np.abs(model.power(k[::100], np.array([0.]))/P0 - 1)) < 1e-14
    # out-of-domain k raises like the reference's InterpolationDomainError
    with pytest.raises(ValueError):
        model.power(np.array([0.9]), np.array([0.5]))


def test_moment_splines_reproduce_knots(cosmo, engine):
    """the not-a-knot splines of P[mu^n] interpolate their knots (RSDSpline on self.k)"""
    model = make_model(cosmo, engine, 0.)
    kmodel = model.k
    P = model.power(kmodel, np.array([0., 0.5, 1.0]))
    gal = engine.galaxy(kmodel, model._config())
    mom = engine.galaxy_moments(gal, 1)[0]           # [pair][moment][nkm]
    assert mom.shape == (10, 4, len(kmodel)) and np.all(np.isfinite(mom))
    # with fs = fcB = 0 and mu = 0 the model is the cAcA P[mu^0] knot values themselves
    model.update(fs=0., fcB=0.)
    P0 = model.power(kmodel, np.array([0.]))
    assert np.max(np.abs(P0/engine.galaxy_moments(gal, 1)[0][0, 0] - 1)) < 1e-13


def test_walker_batch_matches_single_evaluations(cosmo, engine):
    """many walkers on one cosmology (the reference's fitting mode) == one at a time"""
    from pyrsd_b200 import rsd
    model = make_model(cosmo, engine, 0.55)
    k = np.logspace(-2, np.log10(0.4), 30)
    mu = np.linspace(0, 1, 5)
    kk, mm = [a.ravel() for a in np.meshgrid(k, mu, indexing='ij')]
    rng = np.random.default_rng(7)
    pars, stochs, singles = [], [], []
    for i in range(5):
        model.update(sigma8_z=rng.uniform(0.5, 0.8), f=rng.uniform(0.6, 0.9), b1_cA=rng.uniform(1.5, 2.2),
                     alpha_par=rng.uniform(0.97, 1.03), alpha_perp=rng.uniform(0.97, 1.03), fs=rng.uniform(0.05, 0.2))
        pars.append(model._walker())
        stochs.append(model._stoch(model.k))
        singles.append(model.power(kk, mm))
    model._ensure_tables()
    gal = engine.galaxy(model.k, model._config())
    out = engine.galaxy_power(gal, model._spectra, model._result, np.array(pars), np.array(stochs), kk, mm,
                              cmap=np.zeros(5, dtype=np.int32))
    assert np.array_equal(out, np.array(singles))


def test_poles_match_simpson_of_the_power_grid(cosmo, engine, oracle):
    """GalaxySpectrum.poles == MultipoleTransfer(k, ells, Nmu)(power(k, mu)) (rsd/transfers/poles.py:8-69): the fused
    device kernel against the reference's SimpsIntegrate applied to this library's own P(k, mu) grid"""
    from pyrsd_b200 import rsd
    from scipy.special import legendre
    model = rsd.GalaxySpectrum(params=cosmo, z=0.55, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity, engine=engine, interpolate=False)
    model.update(alpha_par=1.02, alpha_perp=0.985, f=0.78)
    k = np.logspace(-2, np.log10(0.4), 30)
    for Nmu in (40, 41, 11):
        n = Nmu + 1 if Nmu % 2 == 0 else Nmu
        mu = np.linspace(0., 1., n)
        P = model.power(k, mu)
        poles = model.poles(k, [0, 2, 4], Nmu=Nmu)
        assert len(poles) == 3
        for ell, got in zip((0, 2, 4), poles):
            kern = (2*ell + 1.)*legendre(ell)(mu)
            ref = np.array([oracle.SimpsIntegrate(mu, kern*row) for row in P])
            assert np.max(np.abs(got - ref)) < 1e-12*np.max(np.abs(P)), (Nmu, ell)
    assert np.array_equal(model.poles(k, 2, Nmu=11), poles[1])           # scalar ell -> one array
    with pytest.raises(Exception):
        model.poles(np.array([5.0]), [0])           # AP-distorted k outside the model's spline domain


def test_chi2_of_an_ensemble(cosmo, engine):
    """chi^2 = (M - D)^T C^-1 (M - D) per walker (rsdfit/driver.py:780-802), fused behind the multipole kernel"""
    from pyrsd_b200 import rsd
    rng = np.random.default_rng(21)
    sp = engine.spectra(cosmo.GetDiscreteK(), cosmo.GetDiscreteTk(), cosmo.n_s(), cosmo.delta_H()**2, cosmo.h(),
                        cosmo.Omega0_m(), cosmo.Omega0_b(), cosmo.Tcmb())
    res = engine.run(sp, result=rsd._new_result())
    kmodel = np.logspace(-3, np.log10(0.5), 200)
    gal = engine.galaxy(kmodel, rsd.galaxy_config())
    ks = rsd.stochasticity_grid(kmodel)
    nw = 37
    wpar = np.zeros((nw, rsd.NPAR)); stoch = np.zeros((nw, rsd.NPAIR, rsd.NSTOCH))
    for i in range(nw):
        b1 = [1.8 + 0.01*i, 2.7, 2.5, 3.4]
        b00, b01 = [b2_00(b) for b in b1], [b2_01(b) for b in b1]
        wpar[i] = rsd.pack_walker(0.6 + 0.002*i, 0.75, 1.0, 1.0, 1.0, b1, 0.10, 0.08, 0.40, 1.0, 4.2, 6.0, 3e4, 9e4, 0., 0., None,
                                  0.7486, cosmo.sigma8(), [b00, b00, b00, b00, b01, b01])
        for p, (a, b) in enumerate(rsd.PAIRS):
            lo, hi = sorted([b1[a], b1[b]])
            stoch[i, p] = stochasticity(ks, wpar[i, 0], lo, hi)
    kd = np.linspace(0.02, 0.38, 19)
    ells = [0, 2, 4]
    cmap = np.zeros(nw, dtype=np.int32)
    poles = engine.galaxy_poles(gal, sp, res, wpar, stoch, kd, ells, 40, cmap=cmap)
    n = 3*len(kd)
    data = poles[5].ravel()*(1 + 0.01*rng.normal(size=n))
    A = rng.normal(size=(n, n))
    cinv = A @ A.T/n + np.eye(n)
    cinv /= np.outer(np.abs(data), np.abs(data))*1e-4
    chi2, M = engine.galaxy_chi2(gal, sp, res, wpar, stoch, kd, ells, data, cinv, 40, cmap=cmap, return_poles=True)
    assert np.array_equal(M, poles)
    d = poles.reshape(nw, n) - data[None, :]
    ref = np.einsum('wi,ij,wj->w', d, cinv, d)
    assert np.allclose(chi2, ref, rtol=1e-12)
    assert np.argmin(chi2) == 5


def test_dm_spectra_vs_shipped_model(cosmo, engine, golden):
    """Pdd, Pdv, Pvv on the model grid against the arrays cached in the reference's shipped model
    (docs/data/galaxy_power.npy: z = 0, default tolerances, Zel'dovich terms read from its interpolation tables):
    the shipped numbers are only as good as epsrel = 1e-4 and a 100 x 300 table allow"""
    from pyrsd_b200 import rsd
    mk = golden['model_k']
    m = rsd.GalaxySpectrum(params=cosmo, z=0., kmin=mk[0], kmax=mk[-1], Nk=len(mk), b2_00=b2_00, b2_01=b2_01,
                           stochasticity=stochasticity, engine=engine)
    m.update(f=float(golden['par_f']))
    assert np.max(np.abs(m.k/mk - 1)) < 1e-12
    dm = m.dm_spectra()
    assert set(dm) == {'P_lin', 'P00', 'P01', 'P11_mu4', 'Pdd', 'Pdv', 'Pvv'}
    for name in ('Pdd', 'Pdv', 'Pvv'):
        err = np.abs(dm[name]/golden['model_' + name] - 1)
        # Pdv / Pvv go through the shipped 100 x 300 Zel'dovich interpolation table: 1.6e-3 at worst, 1e-4 typically
        assert err.max() < 3e-3 and np.median(err) < 3e-4, (name, err.max(), np.median(err))
    # internal consistency: linear theory at low k
    low = mk < 3e-3
    assert np.max(np.abs(dm['P00'][low]/dm['P_lin'][low] - 1)) < 2e-2
    assert np.max(np.abs(dm['P01'][low]/(2*m.f*dm['P_lin'][low]) - 1)) < 5e-2


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/galaxy_ref.npz not generated")
@pytest.mark.parametrize('case,variant', [('z0_default', ''), ('z055_ap', ''), ('z055_spt', ''), ('z0_default', '_pi'), ('z055_ap', '_pi')])
def test_dm_blocks_vs_reference_python(cosmo, engine, case, variant):
    """the dark-matter building blocks on the model grid (P_lin, P00, P01, P11[mu4], Pdv, Pvv, Pdd) against the
    reference's own model objects (power_dm.py:797-884, dm/P00.py .. P11.py), all 200 model k"""
    ref = np.load(REF)
    key = case + variant
    if key + '_dm' not in ref.files:
        pytest.skip("fixture has no dm blocks")
    z, kw, extra = CASES[case]
    model = make_model(cosmo, engine, z, **kw)
    if extra.get('load'):
        from pyrsd_b200 import synthetic as S
        kl = np.logspace(-3.2, 0., 400)
        for term, Pl in S.loaded_terms(kl).items():
            model._load(term, kl, Pl)
    names = [str(n) for n in ref['par_names']]
    pars = dict(zip(names, ref[key + '_pars']))
    pars.pop('D')
    model.update(**pars)
    dm = model.dm_spectra()
    k = model.k
    band = (k >= ZEL_BAND[0]) & (k < ZEL_BAND[1]) if (variant == '' and case != 'z055_spt') else np.zeros(len(k), dtype=bool)
    worst = {}
    for row, name in enumerate(('P_lin', 'P00', 'P01', 'P11_mu4', 'Pdv', 'Pvv', 'Pdd')):
        r = ref[key + '_dm'][row]
        err = np.abs(dm[name] - r)/np.maximum(np.abs(r), 1e-3*ref[key + '_dm'][0])
        i = int(np.argmax(np.where(band, 0, err)))
        worst[name] = [float(err[i]), float(k[i])]
        _record('dm/' + key, worst)
        assert err[i] < TOL, (key, name, err[i], k[i])
        if band.any():
            assert err[band].max() < 2.5e-4, (key, name, err[band].max())


def test_simulation_pdv_identity_and_default_fits(cosmo, engine):
    """Pdv_model_type='sims' (power_dm.py:869-870, rsd/simulation.py:550-660): with the model cosmology standing in for the
    simulation cosmology, at a snapshot redshift and the fiducial f, sigma8_z the normalisation cancels and the model
    returns the measurement itself; the shipped Gaussian processes drive vel_disp_from_sims / correct_mu2 / correct_mu4
    when no callable is given."""
    from pyrsd_b200 import rsd, simulation as sim
    d = sim.dm_sims()
    zs = 0.509
    c = cosmo.clone()
    c._params['growth'] = {0.0: 1.0, 0.509: 0.77, 0.989: 0.61}
    c._params['f'] = {0.0: 0.52, 0.509: 0.75, 0.989: 0.86}
    m = sim.SimulationPdv(c, zs, c.Sigma8_z(zs), c.f_z(zs), sim_cosmo=c)
    k = d['k'][200:480:7]
    assert np.max(np.abs(m(k)/d['Pdv_mu0'][1][200:480:7] - 1)) < 1e-12
    # default normalisation (shipped linear spectrum of the simulations + LCDM growth): Pdv ~ -f P_lin at low k
    m0 = sim.SimulationPdv(cosmo, 0.55, cosmo.Sigma8_z(0.55), cosmo.f_z(0.55))
    kl = np.array([2e-3, 5e-3])
    from pyrsd_b200 import pygcl
    PL = pygcl.LinearPS(cosmo, 0.55)(kl)
    assert np.max(np.abs(m0(kl)/(-cosmo.f_z(0.55)*PL) - 1)) < 0.02
    # through the model: the 'sims' Pdv replaces the Jennings fit on the model grid
    model = rsd.GalaxySpectrum(params=cosmo, z=0.55, b2_00=b2_00, b2_01=b2_01, stochasticity=stochasticity, engine=engine,
                               interpolate=False, Pdv_model_type='sims')
    dm = model.dm_spectra()
    assert np.max(np.abs(dm['Pdv']/m0(model.k) - 1)) < 1e-13
    with model.load_dm_sims('teppei_midz'):
        dm2 = model.dm_spectra()
    from scipy.interpolate import InterpolatedUnivariateSpline as spline
    assert np.max(np.abs(dm2['P00']/spline(d['k'], d['P00_mu0'][1])(model.k) - 1)) < 1e-13
    assert np.max(np.abs(model.dm_spectra()['P00']/dm['P00'] - 1)) < 1e-14          # unloaded again
    # shipped Gaussian processes as the defaults of the simulation options
    g = rsd.GalaxySpectrum(params=cosmo, z=0.55, engine=engine, interpolate=False, vel_disp_from_sims=True, correct_mu2=True,
                           correct_mu4=True, use_mean_bias=True)
    kk, mu = np.logspace(-2, np.log10(0.4), 30), np.linspace(0, 1, 5)
    P1 = g.power(kk, mu)
    g2 = rsd.GalaxySpectrum(params=cosmo, z=0.55, engine=engine, interpolate=False, vel_disp_from_sims=True, correct_mu2=True,
                            correct_mu4=True, use_mean_bias=True, vel_disp_fitter=sim.VelocityDispersionFits(),
                            Pmu2_correction=sim.Pmu2ResidualCorrection(), Pmu4_correction=sim.Pmu4ResidualCorrection())
    assert np.all(np.isfinite(P1)) and np.array_equal(P1, g2.power(kk, mu))
    g.update(correct_mu2=False, correct_mu4=False, vel_disp_from_sims=False, use_mean_bias=False)
    P0 = g.power(kk, mu)
    assert np.max(np.abs(P1/P0 - 1)) > 1e-3          # the options do act


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/galaxy_ref.npz not generated")
@pytest.mark.parametrize('case', ['z055_ap', 'z055_so_gauss'])
def test_sub_spectra_vs_reference_python(cosmo, engine, case):
    """Pgal_cAcA ... Pgal_sBsB and the three groups (power_gal.py:292-399) against the UNMODIFIED reference methods; the
    ten recombine to ``power`` with the galaxy fractions"""
    ref = np.load(REF)
    key = case + '_pi'
    if key + '_sub' not in ref.files:
        pytest.skip("fixture has no sub-spectra for %s" % key)
    z, kw, _ = CASES[case]
    model = make_model(cosmo, engine, z, **kw)
    pars = dict(zip([str(n) for n in ref['par_names']], ref[key + '_pars']))
    pars.pop('D')
    model.update(**pars)
    ks, mus = ref[key + '_sub_k'], ref[key + '_sub_mu']
    K, M = [a.ravel() for a in np.meshgrid(ks, mus, indexing='ij')]
    worst = {}
    got = {}
    for name, r in zip([str(n) for n in ref['sub_names']], ref[key + '_sub']):
        got[name] = np.asarray(getattr(model, 'Pgal_' + name)(K, M)).reshape(len(ks), len(mus))
        worst[name] = float(np.max(np.abs(got[name]/r - 1)))
    _record('sub_spectra/' + key, worst)
    assert max(worst.values()) < TOL, worst
    if not model.use_so_correction:
        fs, fcB, fsB = model.fs, model.fcB, model.fsB
        cc = (1 - fcB)**2*got['cAcA'] + 2*fcB*(1 - fcB)*got['cAcB'] + fcB**2*got['cBcB']
        cs = (1 - fcB)*(1 - fsB)*got['cAsA'] + (1 - fcB)*fsB*got['cAsB'] + fcB*(1 - fsB)*got['cBsA'] + fcB*fsB*got['cBsB']
        ss = (1 - fsB)**2*got['sAsA'] + 2*fsB*(1 - fsB)*got['sAsB'] + fsB**2*got['sBsB']
        total = (1 - fs)**2*cc + 2*fs*(1 - fs)*cs + fs**2*ss
        assert np.max(np.abs(total/model.power(K, M).reshape(total.shape) - 1)) < 1e-12
        assert np.max(np.abs(cc/got['cc'] - 1)) < 1e-12 and np.max(np.abs(ss/got['ss'] - 1)) < 1e-12
    # FOG kernels (gal/fog_kernels.py:40-84)
    x = 0.2*0.5*4.0
    expect = {'modified_lorentzian': 1/(1 + 0.5*x*x)**2, 'lorentzian': 1/(1 + 0.5*x*x), 'gaussian': np.exp(-0.5*x*x)}
    assert abs(model.FOG(0.2, 0.5, 4.0) - expect[model.fog_model]) < 1e-15
    # the halo-pair accessors of the model itself (a GalaxySpectrum IS a BiasedSpectrum in the reference): the moments of the
    # pair (b1, b1_bar) against the fixture's, the terms behind them, and pickling with the table caches they leave behind
    import pickle
    model.b1, model.b1_bar = model.b1_cB, model.b1_sA
    km = model.k
    mom = np.array([model.P_mu0(km), model.P_mu2(km), model.P_mu4(km), model.P_mu6(km)])
    rm = ref[key + '_moments_cBsA']
    # (a moment is a sum of terms of either sign -- P[mu^2] crosses zero near k = 0.3 -- and enters P(k, mu) as
    # mu^(2i) P[mu^(2i)]: where it is small it is measured against 10 % of P[mu^0], the rule of tests/test_gpu_dm.py)
    e_rows = np.max(np.abs(mom - rm)/np.maximum(np.abs(rm), 0.1*np.abs(rm[0])), axis=1)
    _record('pair_moments/' + key, [float(e) for e in e_rows])
    assert e_rows.max() < TOL, e_rows
    assert np.allclose(model.P01_ss.mu2(km) + model.P11_ss.mu2(km) + model.P02_ss.mu2(km), mom[1], rtol=1e-9, atol=1e-9*np.abs(mom[0]).max())
    assert model.bs == (-2./7*(model.b1 - 1.) if model.use_tidal_bias else 0.) and model.sigmav_halo == model.sigma_v
    clone = pickle.loads(pickle.dumps(model))
    assert clone.b1 == model.b1 and clone._result is None and '_dm_tabs' not in clone.__dict__
    model.redshift_params = ['f']
    model.z = 0.55
    assert model.f == cosmo.f_z(0.55)
