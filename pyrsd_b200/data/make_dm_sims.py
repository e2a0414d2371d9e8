"""Convert the reference's dark-matter simulation measurements (pyRSD/data/dark_matter/pkmu_*_z_0.???.dat, read by
``pyRSD.data.P00_mu0_z_0_000()`` ... and used by ``DarkMatterSpectrum.load_dm_sims`` / ``SimulationPdv`` /
``SimulationP11``, rsd/power/dm/power_dm.py:250-279, rsd/simulation.py:433-660) into ``pyrsd_b200/data/dm_sims.npz``.

Stored: the common k grid [500], the four loadable terms at the three snapshot redshifts (0, 0.509, 0.989), and the linear
spectrum of the simulation cosmology at z = 0 the reference ships beside them (Plin_z_0.000.dat) together with the
parameters of data/params/teppei_sims.ini.  The reference normalises the measurements by a CLASS run of that parameter
file; CLASS is not part of this library, so the shipped linear spectrum is the default normalisation and a
``pygcl.Cosmology`` of the simulation can be passed instead (simulation.SimulationPdv, ``sim_cosmo``).

Run in the build container only:   python pyrsd_b200/data/make_dm_sims.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/pyRSD/data/dark_matter'
TAGS = ['0.000', '0.509', '0.989']


def main():
    out = {'z': np.array([float(t) for t in TAGS])}
    for term in ('Pdv_mu0', 'P00_mu0', 'P01_mu2', 'P11_mu4'):
        rows = [np.loadtxt(os.path.join(SRC, 'pkmu_%s_z_%s.dat' % (term, t))) for t in TAGS]
        for r in rows:
            assert np.array_equal(r[:, 0], rows[0][:, 0])
        out['k'] = rows[0][:, 0]
        out[term] = np.array([r[:, 1] for r in rows])
    pl = np.loadtxt(os.path.join(SRC, 'Plin_z_0.000.dat'))
    out['plin_k'], out['plin_P'] = pl[:, 0], pl[:, 1]
    # data/params/teppei_sims.ini
    out['sim_h'], out['sim_Omega_b'], out['sim_Omega_cdm'], out['sim_n_s'] = 0.701, 0.046035, 0.232965, 0.96
    np.savez_compressed(os.path.join(HERE, 'dm_sims.npz'), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
