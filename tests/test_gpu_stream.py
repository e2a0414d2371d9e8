"""The streaming (asynchronous) path of the C-ABI -- prsd_spectra_reload, prsd_plan_run_async,
prsd_galaxy_power_async, prsd_galaxy_wait_copy over page-locked buffers (rsd.PTStream) -- returns bit for bit what
the blocking calls return, batch after batch, including when slots are recycled and when a result is read late."""
import numpy as np
import pytest

from tests.gpu_common import b2_00, b2_01, stochasticity

pytestmark = pytest.mark.gpu


def _batch(golden, tight, n, seed):
    from pyrsd_b200 import rsd
    rng = np.random.default_rng(seed)
    k_i, T_i = golden['cosmo_k'], golden['cosmo_T']
    delta = rng.normal(0, 0.02, n)
    T = T_i[None, :]*(k_i[None, :]/0.05)**(delta[:, None]/2)
    k = np.repeat(k_i[None, :], n, axis=0)
    pars = np.zeros((n, 8))
    pars[:, :6] = [float(golden['cosmo_n_s']), float(tight['delta_H'])**2, float(golden['cosmo_h']), float(tight['Omega0_m']),
                   float(golden['cosmo_Omega_b']), float(tight['Tcmb'])]
    kmodel = np.logspace(-3, np.log10(0.5), 200)
    ks = rsd.stochasticity_grid(kmodel)
    wpar = np.zeros((n, rsd.NPAR))
    stoch = np.zeros((n, rsd.NPAIR, rsd.NSTOCH))
    for i in range(n):
        b1c = rng.uniform(1.2, 2.5)
        b1 = [b1c, 1.5*b1c, 1.4*b1c, 1.9*b1c]
        b00, b01 = [b2_00(b) for b in b1], [b2_01(b) for b in b1]
        wpar[i] = rsd.pack_walker(rng.uniform(0.45, 0.95), rng.uniform(0.4, 1.0), 1.0, 1.0, 1.0, b1, 0.10, 0.08, 0.40, 1.0, 4.2,
                                  6.0, 3e4, 9e4, 0., 0., None, 0.7486, float(golden['cosmo_sigma8']), [b00, b00, b00, b00, b01, b01])
        for p, (a, b) in enumerate(rsd.PAIRS):
            lo, hi = sorted([b1[a], b1[b]])
            stoch[i, p] = stochasticity(ks, wpar[i, 0], lo, hi)
    return dict(k=k, T=T, pars=pars, wpar=wpar, stoch=stoch, kmodel=kmodel)


def test_stream_equals_blocking_calls(golden, tight):
    from pyrsd_b200 import rsd
    eng = rsd.PTEngine()
    n = 6
    kout = np.logspace(-2.9, np.log10(0.45), 37)
    mu = np.linspace(0, 1, 5)
    kk, mm = [a.ravel() for a in np.meshgrid(kout, mu, indexing='ij')]
    batches = [_batch(golden, tight, n, seed) for seed in (1, 2, 3, 4, 5)]
    cfg = rsd.galaxy_config()
    # blocking reference
    gal = eng.galaxy(batches[0]['kmodel'], cfg)
    blocking = []
    for b in batches:
        sp = eng.spectra(b['k'], b['T'], b['pars'][:, 0], b['pars'][:, 1], b['pars'][:, 2], b['pars'][:, 3], b['pars'][:, 4],
                         b['pars'][:, 5])
        res = eng.run(sp)
        blocking.append(eng.galaxy_power(gal, sp, res, b['wpar'], b['stoch'], kk, mm))
    # streaming, depth 2: results read one batch late, slots recycled twice
    st = rsd.PTStream(eng, batches[0]['kmodel'], cfg, n, batches[0]['k'].shape[1], kk, mm, depth=2)
    tickets, got = [], {}
    for i, b in enumerate(batches):
        tickets.append(st.submit(b['k'], b['T'], b['pars'], b['wpar'], b['stoch']))
        if i >= 1:
            got[i - 1] = st.result(tickets[i - 1]).copy()
    got[len(batches) - 1] = st.result(tickets[-1]).copy()
    st.drain()
    for i in range(len(batches)):
        assert np.array_equal(got[i], blocking[i]), i
    with pytest.raises(ValueError):
        st.result(tickets[0])                    # its slot has been recycled
    h2d, d2h = st.bytes_per_step()
    assert d2h == n*len(kk)*8 and h2d > 0


def test_result_of_another_batch_is_refused(golden, tight):
    """a PT result must belong to the spectra batch it is used with"""
    from pyrsd_b200 import rsd, _native
    eng = rsd.PTEngine()
    a, b = _batch(golden, tight, 2, 11), _batch(golden, tight, 2, 12)
    spa = eng.spectra(a['k'], a['T'], a['pars'][:, 0], a['pars'][:, 1], a['pars'][:, 2], a['pars'][:, 3], a['pars'][:, 4], a['pars'][:, 5])
    spb = eng.spectra(b['k'], b['T'], b['pars'][:, 0], b['pars'][:, 1], b['pars'][:, 2], b['pars'][:, 3], b['pars'][:, 4], b['pars'][:, 5])
    ra = eng.run(spa)
    rb = eng.run(spb)
    assert ra is not rb                          # a fresh result per run: ra still holds batch a
    gal = eng.galaxy(a['kmodel'], rsd.galaxy_config())
    k, mu = np.array([0.1, 0.2]), np.array([0.3, 0.6])
    Pa = eng.galaxy_power(gal, spa, ra, a['wpar'], a['stoch'], k, mu)
    Pb = eng.galaxy_power(gal, spb, rb, b['wpar'], b['stoch'], k, mu)
    assert np.all(np.isfinite(Pa)) and not np.array_equal(Pa, Pb)
    with pytest.raises(_native.NativeError):
        eng.galaxy_power(gal, spb, ra, b['wpar'], b['stoch'], k, mu)      # result of batch a with spectra b
