"""Generate ``tests/golden/dm_ref.npz``: the term accessors of the reference's ``DarkMatterSpectrum`` and ``HaloSpectrum``
(``pyRSD.rsd`` imported UNMODIFIED from /root/reference; same stand-ins, tight tables and caches as make_galaxy_ref.py).

Run in the build container only:  python tests/golden/make_dm_ref.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_galaxy_ref as G  # noqa: E402

DM_TERMS = [('P00.mu0', lambda m: m.P00.mu0), ('P01.mu2', lambda m: m.P01.mu2), ('P11.mu2', lambda m: m.P11.mu2),
            ('P11.mu4', lambda m: m.P11.mu4), ('P11.mu4.scalar', lambda m: m.P11.mu4.scalar),
            ('P02.mu2', lambda m: m.P02.mu2), ('P02.mu2.no_velocity', lambda m: m.P02.mu2.no_velocity),
            ('P02.mu2.with_velocity', lambda m: m.P02.mu2.with_velocity), ('P02.mu4', lambda m: m.P02.mu4),
            ('P12.mu4', lambda m: m.P12.mu4), ('P12.mu6', lambda m: m.P12.mu6), ('P22.mu4', lambda m: m.P22.mu4),
            ('P22.mu6', lambda m: m.P22.mu6), ('P03.mu4', lambda m: m.P03.mu4), ('P13.mu4', lambda m: m.P13.mu4),
            ('P13.mu6', lambda m: m.P13.mu6), ('P04.mu4', lambda m: m.P04.mu4), ('P04.mu6', lambda m: m.P04.mu6),
            ('P_mu0', lambda m: m.P_mu0), ('P_mu2', lambda m: m.P_mu2), ('P_mu4', lambda m: m.P_mu4), ('P_mu6', lambda m: m.P_mu6),
            ('Pdd', lambda m: m.Pdd), ('Pdv', lambda m: m.Pdv), ('Pvv', lambda m: m.Pvv)]

DM_CASES = {
    'dm_2loop_spt': dict(include_2loop=True, max_mu=6, interpolate=False, use_P00_model=False, use_P01_model=False,
                         use_P11_model=False, use_Pdv_model=False, use_Pvv_model=False),
    'dm_1loop_hzpt': dict(include_2loop=False, max_mu=4, interpolate=False),
}
DM_UPDATE = dict(sigma8_z=0.61, f=0.77, sigma_v2=198., sigma_bv2=356., sigma_bv4=382., alpha_par=1.02, alpha_perp=0.99)


HALO_TERMS = ['P00_ss.mu0', 'P00_ss_no_stoch.mu0', 'P01_ss.mu2', 'P11_ss.mu2', 'P11_ss.mu4', 'P02_ss.mu2', 'P02_ss.mu4',
              'P12_ss.mu4', 'P12_ss.mu6', 'P22_ss.mu4', 'P22_ss.mu6', 'P03_ss.mu4', 'P13_ss.mu4', 'P13_ss.mu6', 'P04_ss.mu4',
              'P04_ss.mu6']


def main():
    t0 = time.time()
    G.ZEL_SOURCE = 'pi'
    cosmo_mod = G.install_stubs()
    G.install_tight()
    print('dense fixture seeded:', G.seed_from_dense(), flush=True)
    from pyRSD import pygcl
    from pyRSD.rsd import DarkMatterSpectrum, HaloSpectrum

    g = np.load(os.path.join(HERE, 'golden_pt.npz'))
    params = {k[len('cosmo_'):]: float(g[k]) for k in g.files
              if k.startswith('cosmo_') and k not in ('cosmo_k', 'cosmo_T', 'cosmo_sigma8', 'cosmo_tf')}
    params['_growth'] = {0.0: 1.0, 0.55: 0.7486}
    params['_f'] = {0.0: float(g['par_f']), 0.55: 0.7602}
    params['_H'] = {0.0: 100.*params['h'], 0.55: 100.*params['h']*1.3386}
    cosmo = pygcl.Cosmology(pygcl.ClassParams.from_dict(params), 0, float(g['cosmo_sigma8']), g['cosmo_k'], g['cosmo_T'])
    out = {'mu': np.linspace(0., 1., 7)}
    kq = np.logspace(np.log10(0.012), np.log10(0.45), 40)
    out['kq'] = kq
    for name, kw in DM_CASES.items():
        m = DarkMatterSpectrum(params=cosmo_mod.Cosmology(cosmo), z=0.55, **kw)
        m.update(**DM_UPDATE)
        km = np.asarray(m.k)
        out[name + '_modelk'] = km
        rows = []
        for tname, fn in DM_TERMS:
            try:
                rows.append(np.asarray(fn(m)(km), dtype=float)*np.ones_like(km))
            except AttributeError as e:
                # dm/P13.py:41 calls self.normed_power_lin instead of self.m.normed_power_lin: the 1-loop P13[mu^6] (and
                # with it P_mu6) raises in the reference; recorded as NaN
                print('   reference raises for', tname, ':', e)
                rows.append(np.full_like(km, np.nan))
        out[name + '_terms'] = np.array(rows)
        out[name + '_power'] = np.asarray(m.power(kq, out['mu']))
        out[name + '_I02_kq'] = np.asarray(m.I02(kq))
        print(name, 'done %.0fs' % (time.time() - t0), flush=True)
        G.save_cache()
    out['term_names'] = np.array([n for n, _ in DM_TERMS])
    out['halo_term_names'] = np.array(HALO_TERMS)
    # BiasedSpectrum (b1, b1_bar free) and HaloSpectrum (b1_bar = b1) with the synthetic stand-ins of make_galaxy_ref.build_model
    from pyRSD.rsd import BiasedSpectrum

    def fitter(b1, select=None):
        return G.b2_00(b1) if select.startswith('b2_00') else G.b2_01(b1)
    for tag, cls, upd in (('biased', BiasedSpectrum, dict(b1=2.1, b1_bar=3.2)), ('halo', HaloSpectrum, dict(b1=2.4))):
        h = cls(params=cosmo_mod.Cosmology(cosmo), z=0.55, interpolate=False, max_mu=6)
        h._cache['nonlinear_bias_fitter'] = fitter
        h._cache['auto_stochasticity_fits'] = lambda b1, sigma8_z, k: G.stochasticity(k, sigma8_z, b1, b1)
        h._cache['cross_stochasticity_fits'] = lambda b1_1, b1_2, sigma8_z, k: G.stochasticity(k, sigma8_z, b1_1, b1_2)
        h.update(sigma8_z=0.61, f=0.77, **upd)
        km = np.asarray(h.k)
        out[tag + '_modelk'] = km
        out[tag + '_moments'] = np.array([h.P_mu0(km), h.P_mu2(km), h.P_mu4(km), h.P_mu6(km)])
        out[tag + '_moments_kq'] = np.array([h.P_mu0(kq), h.P_mu2(kq), h.P_mu4(kq), h.P_mu6(kq)])
        out[tag + '_power'] = np.asarray(h.power(kq, out['mu']))
        out[tag + '_Phm'] = np.asarray(h.Phm(km))
        out[tag + '_Phm_bar'] = np.asarray(h.Phm_bar(km))
        out[tag + '_stochasticity'] = np.asarray(h.stochasticity(km))
        # the nine halo-halo terms, piece by piece (biased/P00.py ... P22.py)
        rows = []
        for tname in HALO_TERMS:
            term, piece = tname.split('.')
            rows.append(np.asarray(getattr(getattr(h, term), piece)(km), dtype=float)*np.ones_like(km))
        out[tag + '_halo_terms'] = np.array(rows)
        out[tag + '_scalars'] = np.array([h.bs, h.bs_bar, h.sigmav_halo, h.sigmav_halo_bar])
        print(tag, 'done %.0fs' % (time.time() - t0), flush=True)
    path = os.path.join(HERE, 'dm_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    main()
