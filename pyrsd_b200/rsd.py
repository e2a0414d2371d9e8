"""Drop-in for the P(k, mu) evaluation path of ``pyRSD.rsd``.

``GalaxySpectrum`` mirrors the reference class (``pyRSD/rsd/power/gal/power_gal.py:10-425``) for
the call contract the fitting layer uses -- ``model.update(**params)`` followed by
``model.power(k, mu)`` (``pyRSD/rsdfit/theory/base.py:612-699``) -- with the same attribute names and
defaults.  ``PTEngine`` is the batched form: many cosmologies / MCMC walkers per call, which is what
the GPU is for.

Inputs that the reference obtains from Gaussian-process fits to simulations (needing ``george`` and
pandas pickles) are HOST INPUTS here: ``b2_00``, ``b2_01`` (callables of the linear bias) and
``stochasticity`` (callable ``(k, sigma8_z, b1, b1_bar)``); see DESIGN.md.
"""
import ctypes as C

import numpy as np

from . import _native as N
from . import pygcl

INTERP_KMIN = 5e-6      # pyRSD/rsd/__init__.py:3-4
INTERP_KMAX = 1.0
NPAR = 56
NPAIROVR = 8
LOADABLE = ('Pdv', 'P00_mu0', 'P01_mu2', 'P11_mu4')     # rows of prsd_galaxy_set_loaded
NPAIR = 10
NSTOCH = 20

_vp = C.c_void_p
_dp = N._dp
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS')

FOG_MODELS = {'modified_lorentzian': 0, 'lorentzian': 1, 'gaussian': 2}
# (bias a, bias b) of the ten sub-spectra: cAcA cAcB cBcB cAsA cAsB cBsA cBsB sAsA sAsB sBsB
PAIRS = [(0, 0), (0, 1), (1, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3)]


def _L():
    L = pygcl._L()
    if not getattr(L, '_rsd_sigs', False):
        i, d = C.c_int, C.c_double
        L.prsd_galaxy_config_default.argtypes = [_dp]
        L.prsd_galaxy_create.argtypes = [_vp, _vp, i, _dp, i, _dp, C.POINTER(_vp)]
        L.prsd_galaxy_destroy.argtypes = [_vp]
        L.prsd_galaxy_power.argtypes = [_vp, _vp, _vp, i, _vp, _dp, _dp, i, _dp, _dp, _dp]
        L.prsd_galaxy_power_device.argtypes = [_vp, _vp, _vp, i, _vp, _dp, _dp, i, _vp, _vp, _vp]
        L.prsd_galaxy_moments.argtypes = [_vp, i, _dp]
        L.prsd_galaxy_poles.argtypes = [_vp, _vp, _vp, i, _vp, _dp, _dp, i, _dp, i, _ip, i, _dp]
        L.prsd_galaxy_poles_device.argtypes = [_vp, _vp, _vp, i, _vp, _dp, _dp, i, _vp, i, _ip, i, _vp]
        L.prsd_galaxy_chi2.argtypes = [_vp, _vp, _vp, i, _vp, _dp, _dp, i, _dp, i, _ip, i, _dp, _dp, _dp, _vp]
        L.prsd_chi2.argtypes = [_vp, i, i, _vp, _dp, _dp, _dp, i]
        L.prsd_chi2.restype = i
        L.prsd_galaxy_set_mu_corrections.argtypes = [_vp, i, _vp]
        L.prsd_galaxy_set_pair_overrides.argtypes = [_vp, i, _vp]
        L.prsd_galaxy_set_loaded.argtypes = [_vp, C.c_uint, _vp]
        for name in ('prsd_galaxy_set_mu_corrections', 'prsd_galaxy_set_pair_overrides', 'prsd_galaxy_set_loaded'):
            getattr(L, name).restype = i
        for name in ('prsd_galaxy_config_default', 'prsd_galaxy_create', 'prsd_galaxy_destroy', 'prsd_galaxy_power',
                     'prsd_galaxy_power_device', 'prsd_galaxy_moments', 'prsd_galaxy_poles', 'prsd_galaxy_poles_device',
                     'prsd_galaxy_chi2'):
            getattr(L, name).restype = i
        L._rsd_sigs = True
    return L


def k_interp_default():
    """the PT interpolation grid of the reference (power_dm.py:22)"""
    return np.logspace(np.log10(INTERP_KMIN), np.log10(INTERP_KMAX), 250)


class _Plan(N.Handle):
    _destroy = 'prsd_plan_destroy'


class _Result(N.Handle):
    _destroy = 'prsd_result_destroy'


class _Galaxy(N.Handle):
    _destroy = 'prsd_galaxy_destroy'


class PTEngine(object):
    """Batched PT + galaxy-assembly engine: one quadrature plan (cosmology independent) reused for
    any number of cosmologies / walkers."""

    RESULT_SHAPES = {'I': (0, lambda c, nk: (c, 16, nk)), 'J': (1, lambda c, nk: (c, 6, nk)), 'K': (2, lambda c, nk: (c, 13, nk)),
                     'ML': (3, lambda c, nk: (c, 7, 3, nk)), 'oneloop': (4, lambda c, nk: (c, 4, 1000)),
                     'sigmasq_k': (5, lambda c, nk: (c, nk)), 'scalars': (6, lambda c, nk: (c, 4)),
                     'X0': (7, lambda c, nk: (c, 1024)), 'YY': (8, lambda c, nk: (c, 1024))}

    def __init__(self, k_interp=None, quad_config=None, context=None, with_ml=True, with_zel=True):
        self.ctx = context if context is not None else N.default_context()
        self.k_interp = N.as_f64(k_interp_default() if k_interp is None else k_interp)
        keep, cfgp = N.cfg_ptr(quad_config)
        p = _vp()
        flags = (1 if with_ml else 0) | (2 if with_zel else 0)
        N.check(_L().prsd_plan_create(self.ctx.ptr, len(self.k_interp), self.k_interp, cfgp, flags, C.byref(p)))
        self.plan = _Plan(p)
        self._galaxies = {}

    def info(self):
        out = np.empty(8)
        N.check(_L().prsd_plan_info(self.plan.ptr, out))
        return dict(nodes=int(out[0]), device_bytes=int(out[1]), build_seconds=out[2], nk=int(out[3]), n_knots=int(out[4]),
                    nodes_oneloop=int(out[5]), nodes_ik=int(out[6]), nodes_ml_half=int(out[7]))

    def build_profile(self):
        """milliseconds of the plan's build stages (prsd_plan_build_profile)"""
        out = np.zeros(8)
        L = _L()
        L.prsd_plan_build_profile.argtypes = [_vp, _dp]
        L.prsd_plan_build_profile.restype = C.c_int
        N.check(L.prsd_plan_build_profile(self.plan.ptr, out))
        # the host builds the node geometry of the three 2-D operators on threads while the device integrates the 1-D ones
        return dict(geometry_uploads_w_operators_ms=out[0], enqueue_1d_operators_ms=out[2], drain_ms=out[4])

    # ------------------------------------------------------------------ spectra
    def spectra(self, k, T, n_s, amp, h, Omega0_m, Omega0_b, Tcmb=2.7255):
        """upload a batch of linear spectra: k, T [ncosmo][nspl]; the rest scalars or [ncosmo]"""
        k = np.atleast_2d(np.asarray(k, dtype=np.float64))
        T = np.atleast_2d(np.asarray(T, dtype=np.float64))
        nc = T.shape[0]
        if k.shape[0] == 1 and nc > 1:
            k = np.repeat(k, nc, axis=0)
        pars = np.zeros((nc, 8))
        for j, v in enumerate((n_s, amp, h, Omega0_m, Omega0_b, Tcmb)):
            pars[:, j] = v
        return pygcl._Spectra(self.ctx, k, T, pars)

    def run(self, spectra, result=None, sync=True):
        """everything rsd/pt_integrals.py tabulates, for every cosmology of the batch (``sync=False``: enqueue only)"""
        # a fresh result unless the caller passes one to be overwritten: a result handed out earlier must not
        # change under its holder when the engine runs another batch
        res = result if result is not None else _new_result()
        cur = res.ptr.value if isinstance(res.ptr, _vp) else res.ptr
        p = _vp(cur)
        N.check((_L().prsd_plan_run if sync else _L().prsd_plan_run_async)(self.plan.ptr, spectra.ptr, C.byref(p)))
        res.ptr = p
        res.ncosmo = spectra.ncosmo
        spectra.have_oneloop = True
        return res

    def get(self, result, name):
        what, shape = self.RESULT_SHAPES[name]
        out = np.empty(shape(result.ncosmo, len(self.k_interp)))
        N.check(_L().prsd_result_get(result.ptr, what, out, out.size))
        return out

    # ------------------------------------------------------------------- galaxy
    def galaxy(self, kmodel, config=None):
        kmodel = N.as_f64(kmodel)
        cfg = None if config is None else N.as_f64(config)
        key = (kmodel.tobytes(), None if cfg is None else cfg.tobytes())
        if key not in self._galaxies:
            p = _vp()
            N.check(_L().prsd_galaxy_create(self.ctx.ptr, None if cfg is None else cfg.ctypes.data_as(_vp), len(kmodel), kmodel,
                                            len(self.k_interp), self.k_interp, C.byref(p)))
            g = _Galaxy(p)
            g.nkm = len(kmodel)
            self._galaxies[key] = g
        return self._galaxies[key]

    def galaxy_options(self, gal, nw=0, mu_corrections=None, pair_overrides=None, loaded=None):
        """optional inputs of the following galaxy_* calls on ``gal`` with ``nw`` walkers (None clears each):
        ``mu_corrections`` [nw][10][2][nkm] added to the P[mu^2], P[mu^4] knots (correct_mu2 / correct_mu4),
        ``pair_overrides`` [nw][10][8] (use_mean_bias), ``loaded`` dict name -> [nkm] of user-loaded dark-matter terms"""
        L = _L()
        if mu_corrections is None:
            N.check(L.prsd_galaxy_set_mu_corrections(gal.ptr, 0, None))
        else:
            a = np.ascontiguousarray(mu_corrections, dtype=np.float64).reshape(nw, NPAIR, 2, gal.nkm)
            N.check(L.prsd_galaxy_set_mu_corrections(gal.ptr, nw, a.ctypes.data_as(_vp)))
        if pair_overrides is None:
            N.check(L.prsd_galaxy_set_pair_overrides(gal.ptr, 0, None))
        else:
            a = np.ascontiguousarray(pair_overrides, dtype=np.float64).reshape(nw, NPAIR, NPAIROVR)
            N.check(L.prsd_galaxy_set_pair_overrides(gal.ptr, nw, a.ctypes.data_as(_vp)))
        if not loaded:
            N.check(L.prsd_galaxy_set_loaded(gal.ptr, 0, None))
        else:
            a = np.zeros((len(LOADABLE), gal.nkm))
            mask = 0
            for name, v in loaded.items():
                j = LOADABLE.index(name)
                a[j] = v
                mask |= 1 << j
            N.check(L.prsd_galaxy_set_loaded(gal.ptr, mask, a.ctypes.data_as(_vp)))

    def galaxy_power(self, gal, spectra, result, par, stoch, k, mu, cmap=None):
        """P(k, mu) of nw walkers at paired (k, mu) points -> [nw][npts]"""
        par = np.ascontiguousarray(par, dtype=np.float64).reshape(-1, NPAR)
        nw = par.shape[0]
        stoch = np.ascontiguousarray(stoch, dtype=np.float64).reshape(nw, NPAIR, NSTOCH)
        k, mu = N.as_f64(k), N.as_f64(mu)
        assert k.shape == mu.shape
        out = np.empty((nw, len(k)))
        cm = None
        if cmap is not None:
            cm = np.ascontiguousarray(cmap, dtype=np.int32)
        N.check(_L().prsd_galaxy_power(gal.ptr, spectra.ptr, result.ptr, nw, None if cm is None else cm.ctypes.data_as(_vp),
                                       par, stoch, len(k), k, mu, out))
        return out

    def galaxy_poles(self, gal, spectra, result, par, stoch, k, ells, Nmu=40, cmap=None):
        """P_ell(k) of nw walkers -> [nw][nell][nk]; Simpson over Nmu (made odd as MultipoleTransfer does) mu in [0, 1]"""
        par = np.ascontiguousarray(par, dtype=np.float64).reshape(-1, NPAR)
        nw = par.shape[0]
        stoch = np.ascontiguousarray(stoch, dtype=np.float64).reshape(nw, NPAIR, NSTOCH)
        k = N.as_f64(k)
        ells = np.ascontiguousarray(np.atleast_1d(ells), dtype=np.int32)
        if Nmu % 2 == 0:
            Nmu += 1                      # poles.py:29
        out = np.empty((nw, len(ells), len(k)))
        cm = None if cmap is None else np.ascontiguousarray(cmap, dtype=np.int32)
        N.check(_L().prsd_galaxy_poles(gal.ptr, spectra.ptr, result.ptr, nw, None if cm is None else cm.ctypes.data_as(_vp),
                                       par, stoch, len(k), k, len(ells), ells, int(Nmu), out))
        return out

    def galaxy_chi2(self, gal, spectra, result, par, stoch, k, ells, data, cinv, Nmu=40, cmap=None, return_poles=False):
        """chi^2 of nw walkers against ``data`` [nell][nk] with precision matrix ``cinv`` (rsdfit/driver.py:780-802)"""
        par = np.ascontiguousarray(par, dtype=np.float64).reshape(-1, NPAR)
        nw = par.shape[0]
        stoch = np.ascontiguousarray(stoch, dtype=np.float64).reshape(nw, NPAIR, NSTOCH)
        k = N.as_f64(k)
        ells = np.ascontiguousarray(np.atleast_1d(ells), dtype=np.int32)
        n = len(ells)*len(k)
        data = np.ascontiguousarray(data, dtype=np.float64).reshape(n)
        cinv = np.ascontiguousarray(cinv, dtype=np.float64).reshape(n, n)
        if Nmu % 2 == 0:
            Nmu += 1
        chi2 = np.empty(nw)
        poles = np.empty((nw, len(ells), len(k))) if return_poles else None
        cm = None if cmap is None else np.ascontiguousarray(cmap, dtype=np.int32)
        N.check(_L().prsd_galaxy_chi2(gal.ptr, spectra.ptr, result.ptr, nw, None if cm is None else cm.ctypes.data_as(_vp),
                                      par, stoch, len(k), k, len(ells), ells, int(Nmu), data, cinv, chi2,
                                      None if poles is None else poles.ctypes.data_as(_vp)))
        return (chi2, poles) if return_poles else chi2

    def galaxy_power_device(self, gal, spectra, result, par, stoch, d_k, d_mu, cmap=None):
        """P(k, mu) of nw walkers at paired device-resident (k, mu) points -> CUDA tensor [nw][npts]; nothing
        but the walker parameters crosses the host boundary (input of the device-side transfer functions)"""
        import torch
        par = np.ascontiguousarray(par, dtype=np.float64).reshape(-1, NPAR)
        nw = par.shape[0]
        stoch = np.ascontiguousarray(stoch, dtype=np.float64).reshape(nw, NPAIR, NSTOCH)
        assert d_k.is_cuda and d_mu.is_cuda and d_k.dtype == torch.float64 and d_k.shape == d_mu.shape
        out = torch.empty((nw, d_k.numel()), dtype=torch.float64, device=d_k.device)
        torch.cuda.current_stream(d_k.device).synchronize()      # the library runs on its own non-blocking stream
        cm = None if cmap is None else np.ascontiguousarray(cmap, dtype=np.int32)
        N.check(_L().prsd_galaxy_power_device(gal.ptr, spectra.ptr, result.ptr, nw, None if cm is None else cm.ctypes.data_as(_vp),
                                              par, stoch, d_k.numel(), _vp(d_k.data_ptr()), _vp(d_mu.data_ptr()),
                                              _vp(out.data_ptr())))
        return out

    def chi2(self, model, data, cinv):
        """chi^2 of nw model vectors [nw][n] (NumPy array or CUDA tensor) against ``data`` [n] with precision
        matrix ``cinv`` [n][n] (rsdfit/driver.py:780-802) -> [nw]"""
        on_dev = type(model).__module__.split('.')[0] == 'torch'
        if on_dev:
            import torch
            m = model.contiguous().reshape(model.shape[0], -1)
            torch.cuda.current_stream(m.device).synchronize()    # the library runs on its own non-blocking stream
            nw, n = m.shape
            ptr = _vp(m.data_ptr())
        else:
            m = np.ascontiguousarray(model, dtype=np.float64)
            m = m.reshape(m.shape[0], -1)
            nw, n = m.shape
            ptr = m.ctypes.data_as(_vp)
        data = np.ascontiguousarray(data, dtype=np.float64).reshape(n)
        cinv = np.ascontiguousarray(cinv, dtype=np.float64).reshape(n, n)
        out = np.empty(nw)
        N.check(_L().prsd_chi2(self.ctx.ptr, nw, n, ptr, data, cinv, out, int(on_dev)))
        return out

    def galaxy_window_chi2(self, gal, spectra, result, par, stoch, transfer, k_out, data, cinv, cmap=None,
                           return_poles=False):
        """One likelihood evaluation of a survey fit for nw walkers, entirely on the device: P(k, mu) on the grid
        of ``transfer`` (a transfers.WindowFunctionTransfer / GriddedMultipoleTransfer) -> multipoles -> window
        convolution -> spline onto ``k_out`` -> chi^2 against ``data`` in the order [ell][k]
        (rsdfit/theory/base.py:612-699 + rsdfit/driver.py:780-802)"""
        import torch
        dev = 'cuda:%d' % self.ctx.device
        grid = transfer.grid
        if getattr(transfer, '_dgrid', None) is None:
            transfer._dgrid = (torch.from_numpy(np.ascontiguousarray(grid.k).ravel()).to(dev),
                               torch.from_numpy(np.ascontiguousarray(grid.mu).ravel()).to(dev))
        dk, dmu = transfer._dgrid
        P = self.galaxy_power_device(gal, spectra, result, par, stoch, dk, dmu, cmap=cmap)
        P = P.reshape(P.shape[0], grid.Nk, grid.Nmu)
        poles = transfer(P, k_out=k_out) if k_out is not None else transfer(P)          # [nw][nell][nk]
        chi2 = self.chi2(poles, data, cinv)
        return (chi2, poles) if return_poles else chi2

    def galaxy_base(self, gal, nw):
        """dark-matter / bias building blocks of the last galaxy call on the model grid -> [nw][nrows][nkm]"""
        n = C.c_int()
        L = _L()
        L.prsd_galaxy_base.argtypes = [_vp, C.c_int, C.POINTER(C.c_int), _vp]
        L.prsd_galaxy_base.restype = C.c_int
        N.check(L.prsd_galaxy_base(gal.ptr, nw, C.byref(n), None))
        out = np.empty((nw, n.value, gal.nkm))
        N.check(L.prsd_galaxy_base(gal.ptr, nw, C.byref(n), out.ctypes.data_as(_vp)))
        return out

    def galaxy_moments(self, gal, nw):
        out = np.empty((nw, NPAIR, 4, gal.nkm))
        N.check(_L().prsd_galaxy_moments(gal.ptr, nw, out))
        return out


class PTStream(object):
    """Pipelined evaluation of a stream of batches: every PT table rebuilt + P(k, mu) per cosmology, batch after
    batch, with nothing on the host waiting for the device between ``submit`` calls.

    ``submit`` copies a batch's inputs into page-locked slots and enqueues upload -> PT stage -> galaxy assembly ->
    copy of P(k, mu) to a page-locked output slot; the copy runs on a second CUDA stream under the PT kernels of the
    NEXT batch.  ``result(ticket)`` waits for that batch only and returns a view of its output slot, valid until
    ``depth`` further batches have been submitted.  Results are bit-identical to the blocking
    ``spectra -> run -> galaxy_power`` sequence (tests/test_gpu_stream.py)."""

    def __init__(self, engine, kmodel, config, ncosmo, k_T_shape, k, mu, depth=2):
        self.eng, self.depth = engine, int(depth)
        self.gal = engine.galaxy(kmodel, config)
        self.nc = int(ncosmo)
        self.nspl = int(k_T_shape)
        k, mu = N.as_f64(k), N.as_f64(mu)
        assert k.shape == mu.shape
        self.npts = len(k)
        P = N.PinnedArray
        self.h_kk, self.h_mm = P(self.npts), P(self.npts)
        self.h_kk.array[:], self.h_mm.array[:] = k, mu
        self.slots = [dict(k=P((self.nc, self.nspl)), T=P((self.nc, self.nspl)), pars=P((self.nc, 8)),
                           wpar=P((self.nc, NPAR)), stoch=P((self.nc, NPAIR, NSTOCH)), out=P((self.nc, self.npts)),
                           ticket=0) for _ in range(self.depth)]
        self.sp = None
        self.res = _new_result()
        self.n = 0

    def submit(self, k, T, pars, wpar, stoch):
        """enqueue one batch: k, T [ncosmo][nspl], pars [ncosmo][8] (prsd_spectra_create), wpar [ncosmo][NPAR],
        stoch [ncosmo][NPAIR][NSTOCH]; returns the batch's ticket"""
        L = _L()
        s = self.slots[self.n % self.depth]
        if s['ticket']:
            # the slot's previous batch: once its output has landed, its inputs were consumed long ago
            N.check(L.prsd_galaxy_wait_copy(self.gal.ptr, s['ticket']))
        for name, a in (('k', k), ('T', T), ('pars', pars), ('wpar', wpar), ('stoch', stoch)):
            s[name].array[...] = np.asarray(a).reshape(s[name].shape)
        if self.sp is None:
            self.sp = pygcl._Spectra(self.eng.ctx, s['k'].array, s['T'].array, s['pars'].array)
        else:
            N.check(L.prsd_spectra_reload(self.sp.ptr, s['k'].ptr, s['T'].ptr, s['pars'].ptr))
        self.eng.run(self.sp, result=self.res, sync=False)
        t = C.c_longlong()
        # the evaluation points go up with the first batch and stay on the device
        kk, mm = (self.h_kk.ptr, self.h_mm.ptr) if self.n == 0 else (None, None)
        N.check(L.prsd_galaxy_power_async(self.gal.ptr, self.sp.ptr, self.res.ptr, self.nc, None, s['wpar'].ptr, s['stoch'].ptr,
                                          self.npts, kk, mm, s['out'].ptr, C.byref(t)))
        s['ticket'] = t.value
        self.n += 1
        return t.value

    def result(self, ticket):
        """P(k, mu) [ncosmo][npts] of the batch that returned ``ticket`` (a view of page-locked memory)"""
        for s in self.slots:
            if s['ticket'] == ticket:
                N.check(_L().prsd_galaxy_wait_copy(self.gal.ptr, ticket))
                return s['out'].array
        raise ValueError("ticket %d is no longer held (depth = %d)" % (ticket, self.depth))

    def drain(self):
        N.check(_L().prsd_galaxy_wait(self.gal.ptr))

    def bytes_per_step(self):
        s = self.slots[0]
        # per step: the spectra and the walker parameters; the evaluation points (k, mu) go up once with the first batch
        h2d = sum(s[n].array.nbytes for n in ('k', 'T', 'pars', 'wpar', 'stoch'))
        return h2d, s['out'].array.nbytes


def galaxy_config(max_mu=4, fog_model='modified_lorentzian', use_P00_model=True, use_P01_model=True, use_P11_model=True,
                  use_Pdv_model=True, use_Pvv_model=True, use_Phm_model=True, use_tidal_bias=False,
                  use_so_correction=False, k0_low=5e-3, interpolate=False):
    cfg = np.empty(64)
    N.check(_L().prsd_galaxy_config_default(cfg))
    if fog_model not in FOG_MODELS:
        raise ValueError("`fog_model` must be one of %s" % sorted(FOG_MODELS))
    if max_mu not in (0, 2, 4, 6):
        raise ValueError("`max_mu` must be one of [0, 2, 4, 6]")
    cfg[0], cfg[1] = max_mu, FOG_MODELS[fog_model]
    cfg[2:10] = [use_P00_model, use_P01_model, use_P11_model, use_Pdv_model, use_Pvv_model, use_Phm_model,
                 use_tidal_bias, use_so_correction]
    cfg[10] = k0_low
    cfg[11] = 1.0 if interpolate else 0.0
    return cfg


def pack_walker(sigma8_z, f, alpha_par, alpha_perp, alpha_drag, b1, fs, fcB, fsB, sigma_c, sigma_sA, sigma_sB, NcBs, NsBsB,
                f_so, sigma_so, sigma_v, D, sigma8, b2):
    """one row of the walker-parameter matrix of the C-ABI (include/pyrsd_b200.h)"""
    p = np.zeros(NPAR)
    p[0:5] = sigma8_z, f, alpha_par, alpha_perp, alpha_drag
    p[5:9] = b1
    p[9:19] = fs, fcB, fsB, sigma_c, sigma_sA, sigma_sB, NcBs, NsBsB, f_so, sigma_so
    p[19], p[20], p[21] = (sigma_v if sigma_v is not None else -1.0), D, sigma8
    p[22:46] = np.asarray(b2, dtype=float).reshape(24)
    return p


def stochasticity_grid(kmodel):
    """the 20 log-spaced k the stochasticity is tabulated on (power_biased.py:357)"""
    return np.logspace(np.log10(kmodel.min()), np.log10(kmodel.max()), NSTOCH)


# polynomial nonlinear-bias defaults shipped by the reference (data/simulation_fits/b2_0?_fits_runPB.npz,
# exposed as `nonlinear_bias_defaults`, nonlinear_biasing.py:54-72): coefficients of b1^0, b1^2, b1^4 for
# b2_00 and of b1^0, b1^1, b1^2 for b2_01 (the "_a" variants, used for all terms under `use_vlah_biasing`)
_B2_00_A = (-0.9097586321233345, 0.42328044892648975, 0.008338610073394552)
_B2_01_A = (0.4541869219609199, -1.4949433955934217, 0.7001525446354208)


def default_b2_00(b1):
    return _B2_00_A[0] + _B2_00_A[1]*b1**2 + _B2_00_A[2]*b1**4


def default_b2_01(b1):
    return _B2_01_A[0] + _B2_01_A[1]*b1 + _B2_01_A[2]*b1**2


class GalaxySpectrum(object):
    """The model for the galaxy redshift-space power spectrum (power_gal.py:10-425).

    Same keyword arguments / attributes as the reference where they belong to the P(k, mu) path;
    ``params`` must be a :class:`pyrsd_b200.pygcl.Cosmology`."""

    _PARAMS = ('sigma8_z', 'f', 'alpha_par', 'alpha_perp', 'alpha_drag', 'b1_cA', 'b1_cB', 'b1_sA', 'b1_sB', 'fs', 'fcB',
               'fsB', 'sigma_c', 'sigma_s', 'sigma_sA', 'sigma_sB', 'NcBs', 'NsBsB', 'N', 'f_so', 'sigma_so', 'sigma_v',
               'sigma_v2', 'sigma_bv2', 'sigma_bv4')

    use_mean_bias = vel_disp_from_sims = correct_mu2 = correct_mu4 = False
    Pdv_model_type = 'jennings'
    vel_disp_fitter = Pmu2_correction = Pmu4_correction = Pdv_sim_cosmo = None
    _loaded_data = {}

    def __init__(self, kmin=1e-3, kmax=0.5, Nk=200, z=0., params=None, include_2loop=False, transfer_fit="CLASS",
                 max_mu=4, interpolate=True, k0_low=5e-3, Pdv_model_type='jennings', fog_model='modified_lorentzian',
                 use_so_correction=False, use_tidal_bias=False, use_mean_bias=False, vel_disp_from_sims=False,
                 correct_mu2=False, correct_mu4=False, use_vlah_biasing=True, b2_00=None, b2_01=None,
                 stochasticity=None, vel_disp_fitter=None, Pmu2_correction=None, Pmu4_correction=None, Pdv_sim_cosmo=None,
                 engine=None, **kwargs):
        if not isinstance(params, pygcl.Cosmology):
            raise TypeError("`params` must be a pyrsd_b200.pygcl.Cosmology (the linear T(k) is a host input)")
        if include_2loop:
            raise ValueError("`include_2loop` is forced to False for biased tracers (power_biased.py:47)")
        if Pdv_model_type not in ('jennings', 'sims'):
            raise ValueError("`Pdv_model_type` must be one of ['jennings', 'sims']")       # power_dm.py:638-640
        if kmin < INTERP_KMIN:
            raise ValueError("kmin = %.2e h/Mpc below PT integrals interpolation lower bound" % kmin)
        if kmax > INTERP_KMAX:
            raise ValueError("kmax = %.2e h/Mpc above PT integrals interpolation upper bound" % kmax)
        self.cosmo = params
        self.z, self.kmin, self.kmax, self.Nk = z, kmin, kmax, Nk
        self.max_mu, self.fog_model, self.use_so_correction, self.use_tidal_bias = max_mu, fog_model, use_so_correction, use_tidal_bias
        self.k0_low = k0_low
        self.include_2loop = False
        # True (the reference's default, power_dm.py:34): Zel'dovich terms from the sigma8_z x k tables of
        # rsd/hzpt/interpolated.py, built on the device in one batched launch; False: exact at every call
        self.interpolate = bool(interpolate)
        self.transfer_fit = transfer_fit
        self.use_vlah_biasing = use_vlah_biasing
        # simulation-calibrated options of HaloSpectrum (power_biased.py:26-52)
        self.use_mean_bias, self.vel_disp_from_sims = bool(use_mean_bias), bool(vel_disp_from_sims)
        self.correct_mu2, self.correct_mu4 = bool(correct_mu2), bool(correct_mu4)
        self.Pdv_model_type = Pdv_model_type
        self._loaded_data = {}              # SimLoaderMixin (rsd/sim_loader.py): name -> spline of the loaded term
        self._models = {}
        for kw in ('use_P00_model', 'use_P01_model', 'use_P11_model', 'use_Pdv_model', 'use_Pvv_model', 'use_Phm_model'):
            self._models[kw] = bool(kwargs.pop(kw, True))
        if kwargs:
            raise TypeError("unexpected keyword arguments %s" % sorted(kwargs))
        # the reference's defaults are its Gaussian-process simulation fits (use_vlah_biasing: one b2_00 and one b2_01
        # for all terms, nonlinear_biasing.py:10-23; stochasticity: power_biased.py:342-366), predicted on the device
        # from the shipped training data (pyrsd_b200/simulation.py); any callable with the same signature replaces them
        from . import simulation as _sim
        self.b2_00 = b2_00 if b2_00 is not None else _sim.b2_00
        self.b2_01 = b2_01 if b2_01 is not None else _sim.b2_01
        self.stochasticity = stochasticity if stochasticity is not None else _sim.stochasticity
        # the fits behind vel_disp_from_sims / correct_mu2 / correct_mu4 (power_biased.py:181-200), keyword callables
        # ``(b1=, sigma8_z=)`` and ``(f=, sigma8_z=, b1=, k=)``; None: the shipped Gaussian processes
        self.vel_disp_fitter, self.Pmu2_correction, self.Pmu4_correction = vel_disp_fitter, Pmu2_correction, Pmu4_correction
        # Pdv_model_type='sims': pygcl.Cosmology of the simulations (simulation.SimulationPdv); None: shipped normalisation
        self.Pdv_sim_cosmo = Pdv_sim_cosmo
        # defaults of DarkMatterSpectrum / GalaxySpectrum (power_dm.py:119-126, power_gal.py:103-123)
        self.sigma8_z = params.Sigma8_z(z)
        self.f = params.f_z(z)
        self.alpha_par = self.alpha_perp = self.alpha_drag = 1.
        self.sigma_v = None
        self.sigma_v2 = self.sigma_bv2 = self.sigma_bv4 = 0.
        self.fs, self.fcB, self.fsB = 0.10, 0.08, 0.40
        self.b1_cA, self.b1_cB, self.b1_sA, self.b1_sB = 1.85, 2.8, 2.6, 3.6
        self.sigma_c, self.sigma_s, self.sigma_sA, self.sigma_sB = 1., 5., 4.2, 6.
        self.NcBs, self.NsBsB, self.N = 3e4, 9e4, 0.
        self.f_so = self.sigma_so = 0.
        self._engine = engine
        self._spectra = None
        self._result = None

    # -------------------------------------------------------------- pickling
    def __getstate__(self):
        """everything but the device handles: the PT tables are rebuilt on first use (a millisecond here, minutes in
        the reference, which therefore pickles them: power_dm.py:720-735 ``to_npy`` / ``from_npy``)"""
        d = dict(self.__dict__)
        for k in ('_engine', '_spectra', '_result'):
            d[k] = None
        for k in ('_dm_tabs', '_b2s', '_b2s_key'):
            d.pop(k, None)
        return d

    def __setstate__(self, state):
        self.__dict__.update(state)

    def to_npy(self, filename):
        """``np.save`` of the pickled model, as the reference's ``DarkMatterSpectrum.to_npy``"""
        np.save(filename, np.array(self, dtype=object), allow_pickle=True)

    @classmethod
    def from_npy(cls, filename):
        return np.load(filename, allow_pickle=True).tolist()

    # -------------------------------------------------------------- derived
    def __getattr__(self, name):
        models = self.__dict__.get('_models', {})
        if name in models:
            return models[name]
        raise AttributeError(name)

    @property
    def k(self):
        """the spline domain of the model (power_dm.py:526-538)"""
        return np.logspace(np.log10(max(self.kmin, INTERP_KMIN)), np.log10(min(self.kmax, INTERP_KMAX)), self.Nk)

    @property
    def D(self):
        return self.cosmo.D_z(self.z)

    @property
    def _power_norm(self):
        return (self.sigma8_z/self.cosmo.sigma8())**2

    @property
    def b1_c(self):
        return self.fcB*self.b1_cB + (1. - self.fcB)*self.b1_cA

    @property
    def b1_s(self):
        return self.fsB*self.b1_sB + (1. - self.fsB)*self.b1_sA

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if k in self._models:
                self._models[k] = bool(v)
            elif k in self._PARAMS or k in ('max_mu', 'fog_model', 'use_so_correction', 'use_tidal_bias', 'b2_00', 'b2_01',
                                            'stochasticity', 'interpolate', 'k0_low', 'use_mean_bias', 'vel_disp_from_sims',
                                            'correct_mu2', 'correct_mu4', 'Pdv_model_type'):
                setattr(self, k, v)
            else:
                raise RuntimeError("failure to set parameter `%s` to value %s: unknown parameter" % (k, v))

    # ---------------------------------------------------------------- tables
    def _ensure_tables(self):
        if self._engine is None:
            self._engine = PTEngine(context=self.cosmo._ctx)
        if self._result is None:
            c = self.cosmo
            self._spectra = self._engine.spectra(c.GetDiscreteK(), c.GetDiscreteTk(), c.n_s(), c.delta_H()**2, c.h(),
                                                 c.Omega0_m(), c.Omega0_b(), c.Tcmb())
            self._result = self._engine.run(self._spectra, result=_new_result())
        return self._engine

    def initialize(self):
        """build every PT / Zel'dovich table (the reference's minutes-long step, power_dm.py:154-159)"""
        return self.power(0.5*(self.kmin + self.kmax), 0.5)

    def pt_table(self, name):
        """unnormalised PT tables on k_interp: 'I', 'J', 'K', 'ML', 'oneloop', 'sigmasq_k', 'scalars', 'X0', 'YY'"""
        eng = self._ensure_tables()
        return eng.get(self._result, name)[0]

    @property
    def sigma_lin(self):
        return np.sqrt(self._power_norm*self.pt_table('scalars')[0])

    @property
    def velocity_kurtosis(self):
        return self._power_norm**2*self.pt_table('scalars')[1]

    def _pair_biases(self):
        """(_ib1, _ib1_bar) of the ten sub-spectra (power_biased.py:219-237)"""
        b1 = [self.b1_cA, self.b1_cB, self.b1_sA, self.b1_sB]
        if not self.use_mean_bias:
            return [(b1[a], b1[b]) for a, b in PAIRS]
        return [((b1[a]*b1[b])**0.5,)*2 for a, b in PAIRS]

    def _sigmav_halo(self, b1):
        """sigmav_halo at a bias (power_biased.py:268-286); <= 0 stands for ``sigma_v``"""
        from . import simulation as _sim
        if not self.vel_disp_from_sims:
            return -1.0
        fit = self.vel_disp_fitter if self.vel_disp_fitter is not None else _sim._shared(_sim.VelocityDispersionFits)
        return float(fit(b1=b1, sigma8_z=self.sigma8_z))

    def _walker(self):
        b1 = [self.b1_cA, self.b1_cB, self.b1_sA, self.b1_sB]
        b00 = [self.b2_00(b) for b in b1]
        b01 = [self.b2_01(b) for b in b1]
        b2 = [b00, b00, b00, b00, b01, b01]
        p = pack_walker(self.sigma8_z, self.f, self.alpha_par, self.alpha_perp, self.alpha_drag, b1, self.fs, self.fcB,
                        self.fsB, self.sigma_c, self.sigma_sA, self.sigma_sB, self.NcBs, self.NsBsB, self.f_so,
                        self.sigma_so, self.sigma_v, self.D, self.cosmo.sigma8(), b2)
        p[46:50] = [self._sigmav_halo(b) for b in b1]
        return p

    def _stoch(self, kmodel):
        ks = stochasticity_grid(kmodel)
        out = np.zeros((NPAIR, NSTOCH))
        if self.stochasticity is not None:
            for p, (a, b) in enumerate(self._pair_biases()):
                lo, hi = sorted([a, b])
                out[p] = self.stochasticity(ks, self.sigma8_z, lo, hi)
        return out

    def _pair_overrides(self):
        """per-pair inputs of ``use_mean_bias``: the mean bias and everything evaluated at it"""
        if not self.use_mean_bias:
            return None
        out = np.zeros((NPAIR, NPAIROVR))
        for p, (bm, _) in enumerate(self._pair_biases()):
            b00, b01 = self.b2_00(bm), self.b2_01(bm)
            out[p] = [bm, b00, b00, b00, b00, b01, b01, self._sigmav_halo(bm)]
        return out

    def _mu_corrections(self, kmodel):
        """mu2_model_correction / mu4_model_correction on the model grid (power_biased.py:462-480): the residual
        fits at the geometric-mean bias of each pair"""
        if not (self.correct_mu2 or self.correct_mu4):
            return None
        from . import simulation as _sim
        out = np.zeros((NPAIR, 2, len(kmodel)))
        bm = np.array([(a*b)**0.5 for a, b in self._pair_biases()])[:, None]
        for j, (on, fit, cls) in enumerate(((self.correct_mu2, self.Pmu2_correction, _sim.Pmu2ResidualCorrection),
                                            (self.correct_mu4, self.Pmu4_correction, _sim.Pmu4ResidualCorrection))):
            if on:
                fit = fit if fit is not None else _sim._shared(cls)
                out[:, j] = fit(f=self.f, sigma8_z=self.sigma8_z, b1=bm, k=kmodel[None, :])
        return out

    # ---- SimLoaderMixin (rsd/sim_loader.py:44-75)
    def _load(self, name, k, P):
        """load data into a dark-matter term (``Pdv``, ``P00_mu0``, ``P01_mu2``, ``P11_mu4``): the term becomes the
        cubic spline of (k, P), evaluated on the model grid"""
        from scipy.interpolate import InterpolatedUnivariateSpline as spline
        if name not in LOADABLE:
            raise ValueError("`%s` not a valid term to be loaded; must be one of %s" % (name, list(LOADABLE)))
        self._loaded_data[name] = spline(k, P)

    def _unload(self, name):
        self._loaded_data.pop(name, None)

    def load_dm_sims(self, val):
        """context manager: P00_mu0, P01_mu2, P11_mu4 and Pdv from the simulation snapshot ``teppei_lowz`` (z = 0),
        ``teppei_midz`` (0.509) or ``teppei_highz`` (0.989) (power_dm.py:250-279)"""
        import contextlib
        from . import simulation as _sim
        tags = {'teppei_lowz': 0, 'teppei_midz': 1, 'teppei_highz': 2}
        if val not in tags:
            raise ValueError("Allowed simulations to load are %s" % sorted(tags))

        @contextlib.contextmanager
        def cm():
            d = _sim.dm_sims()
            for name, key in (('P00_mu0', 'P00_mu0'), ('P01_mu2', 'P01_mu2'), ('P11_mu4', 'P11_mu4'), ('Pdv', 'Pdv_mu0')):
                self._load(name, d['k'], d[key][tags[val]])
            try:
                yield
            finally:
                for name in ('P00_mu0', 'P01_mu2', 'P11_mu4', 'Pdv'):
                    self._unload(name)
        return cm()

    def _loaded_knots(self, kmodel):
        loaded = {n: s(kmodel) for n, s in self._loaded_data.items()}
        if 'Pdv' not in loaded and self.Pdv_model_type == 'sims' and self._models['use_Pdv_model']:
            from . import simulation as _sim
            loaded['Pdv'] = _sim.SimulationPdv(self.cosmo, self.z, self.sigma8_z, self.f, self.Pdv_sim_cosmo)(kmodel)  # power_dm.py:869-870
        return loaded

    def _state(self, kmodel):
        """everything one walker of the assembly kernels reads at the current attribute values"""
        return dict(par=self._walker(), stoch=self._stoch(kmodel), corr=self._mu_corrections(kmodel), ovr=self._pair_overrides())

    def _evaluate(self, states, kind, *args, **kw):
        """run ``states`` as one batch of walkers of the engine's ``galaxy_<kind>`` on this model's tables"""
        eng = self._ensure_tables()
        kmodel = self.k
        gal = eng.galaxy(kmodel, self._config())
        nw = len(states)
        par, stoch = np.array([s['par'] for s in states]), np.array([s['stoch'] for s in states])
        corr = None if states[0]['corr'] is None else np.array([s['corr'] for s in states])
        ovr = None if states[0]['ovr'] is None else np.array([s['ovr'] for s in states])
        loaded = self._loaded_knots(kmodel)
        opts = corr is not None or ovr is not None or bool(loaded)
        if opts:
            eng.galaxy_options(gal, nw, corr, ovr, loaded)
        try:
            out = getattr(eng, 'galaxy_' + kind)(gal, self._spectra, self._result, par, stoch, *args,
                                                 cmap=np.zeros(nw, dtype=np.int32), **kw)
        finally:
            # the handle is shared by every model of this engine with the same grid and configuration (and by the
            # streaming path): the optional inputs never outlive the call that set them
            if opts:
                eng.galaxy_options(gal)
        if np.isnan(out).any():
            raise ValueError("wavenumber range outside interpolation domain: AP-distorted k outside [%.4e, %.4e]; "
                             "try adjusting `kmin` / `kmax`" % (kmodel[0], kmodel[-1]))
        return eng, gal, out

    def _config(self):
        return galaxy_config(max_mu=self.max_mu, fog_model=self.fog_model, use_tidal_bias=self.use_tidal_bias,
                             use_so_correction=self.use_so_correction, k0_low=self.k0_low, interpolate=self.interpolate,
                             **self._models)

    def power(self, k, mu, flatten=False):
        """P(k, mu) with the broadcasting rules of ``tools.broadcast_kmu`` (rsd/tools.py:225-266)"""
        k = np.atleast_1d(np.asarray(k, dtype=np.float64))
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        if k.ndim == 1 and mu.ndim == 1:
            if len(mu) != len(k):
                k, mu = k[:, None], mu[None, :]
        elif k.ndim > 1 and mu.ndim < 2:
            mu = mu[None, :]
        elif mu.ndim > 1 and k.ndim < 2:
            k = k[:, None]
        kb, mub = np.broadcast_arrays(k, mu)
        if kb.ndim > 2:
            raise ValueError("incompatible `k`, `mu` dimensions for broadcasted; arrays should have maximum dimension of 2")
        P = self._evaluate([self._state(self.k)], 'power', np.ascontiguousarray(kb).ravel(), np.ascontiguousarray(mub).ravel())[2][0]
        P = np.squeeze(P.reshape(kb.shape))
        return np.ravel(P, order='F') if flatten else P


def _poles(self, k, ells, Nmu=40):
    """``GalaxySpectrum.poles(k, ells, Nmu=40)`` (power_dm.py:1016-1050): list of P_ell(k), one array per ell"""
    scalar = np.isscalar(ells)
    P = self._evaluate([self._state(self.k)], 'poles', k, ells, Nmu)[2][0]
    return P[0] if scalar else [P[i] for i in range(P.shape[0])]


GalaxySpectrum.poles = _poles


def _pole_method(ell):
    def method(self, k, **kwargs):
        """P_ell(k) by Simpson's rule over 41 mu in [0, 1] (the ``tools.monopole`` ... ``tetrahexadecapole`` decorators of
        rsd/tools.py:268-325 around ``power``; they use 42 samples when len(k) == 41, which the device rule -- an odd
        number of samples -- does not reproduce: 41 is used throughout)"""
        return _poles(self, np.atleast_1d(np.asarray(k, dtype=float)), ell, Nmu=41)
    return method


GalaxySpectrum.monopole = _pole_method(0)
GalaxySpectrum.quadrupole = _pole_method(2)
GalaxySpectrum.hexadecapole = _pole_method(4)
GalaxySpectrum.tetrahexadecapole = _pole_method(6)


def _apply_transfer(self, transfer):
    """the power spectrum on the grid of ``transfer`` (a transfers.* object) with the transfer applied: grid binning,
    multipoles, wedges, window convolution (power_dm.py:1046-1077)"""
    import warnings
    from . import transfers as T
    g = transfer.grid
    if isinstance(transfer, T.WindowFunctionTransfer):
        kmin, kmax = float(np.nanmin(g.k)), float(np.nanmax(g.k))
        if kmin < self.kmin:
            warnings.warn("min k of window transfer (%s) is less than model's kmin (%.2e)" % (kmin, self.kmin))
        if kmax > self.kmax:
            warnings.warn("max k of window transfer (%s) is greater than model's kmax (%.2e)" % (kmax, self.kmax))
    # the model is evaluated on the cells that hold modes (transfer.flatk / flatmu of the reference); the others carry
    # zero weight in every transfer
    ok = g.notnull
    P = np.zeros(g.shape)
    P[ok] = self.power(np.ascontiguousarray(g.k[ok]), np.ascontiguousarray(g.mu[ok]))
    return transfer(P)


GalaxySpectrum.apply_transfer = _apply_transfer


def _derivative(self, k, mu, wrt, rel=1e-4):
    """d P / d k' or d P / d mu' at the Alcock-Paczynski coordinates (k', mu') of the observed (k, mu), with the volume and
    sound-horizon factors of ``tools.alcock_paczynski`` (power_gal.py:427-441).  The reference chains analytic
    derivatives of every term; here a central difference of the undistorted model (alpha = 1) at (k', mu'), one batch of
    two point sets (truncation error ~ rel^2 = 1e-8 of the curvature scale)."""
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    if k.ndim == 1 and mu.ndim == 1 and len(mu) != len(k):
        k, mu = k[:, None], mu[None, :]
    kb, mub = np.broadcast_arrays(k, mu)
    F = self.alpha_par/self.alpha_perp
    fac = np.sqrt(1. + mub**2*(1./F**2 - 1.))
    kt, mut = (kb/self.alpha_perp*fac).ravel(), (mub/F/fac).ravel()
    saved = self.alpha_par, self.alpha_perp, self.alpha_drag
    try:
        self.alpha_par = self.alpha_perp = self.alpha_drag = 1.
        if wrt == 'k':
            h = rel*kt
            lo, hi = self.power(kt - h, mut), self.power(kt + h, mut)
        else:
            # one-sided where the step would leave [0, 1]
            h = np.full(len(mut), rel)
            up, dn = np.minimum(mut + h, 1.), np.maximum(mut - h, 0.)
            lo, hi, h = self.power(kt, dn), self.power(kt, up), 0.5*(up - dn)
    finally:
        self.alpha_par, self.alpha_perp, self.alpha_drag = saved
    d = (np.atleast_1d(hi) - np.atleast_1d(lo))/(2.*h)
    d = d*self.alpha_drag**3/(self.alpha_perp**2*self.alpha_par)
    return np.squeeze(d.reshape(kb.shape))


GalaxySpectrum.derivative_k = lambda self, k, mu: _derivative(self, k, mu, 'k')
GalaxySpectrum.derivative_mu = lambda self, k, mu: _derivative(self, k, mu, 'mu')


# the parameters that enter a walker row (pack_walker)
_GRAD_PARAMS = ('sigma8_z', 'f', 'alpha_par', 'alpha_perp', 'alpha_drag', 'b1_cA', 'b1_cB', 'b1_sA', 'b1_sB', 'fs', 'fcB', 'fsB',
                'sigma_c', 'sigma_sA', 'sigma_sB', 'NcBs', 'NsBsB', 'f_so', 'sigma_so', 'sigma_v')


def _gradient(self, k, mu, names, epsilon=1e-4):
    """d power(k, mu) / d theta for the parameters ``names`` -> array [len(names)][npts].

    The reference (``GalaxySpectrum.get_gradient`` -> ``PkmuGradient.__call__``, rsd/power/gradient.py:95-203)
    mixes hand-derived derivatives for the nuisance parameters (rsd/power/gal/derivatives/*.py) with central
    finite differences of ``power`` (absolute step ``epsilon`` = 1e-4) for the rest, one model evaluation per
    perturbed parameter vector.  Here every parameter uses the central difference and the 2 len(names) perturbed
    models are ONE batch of walkers of the assembly kernel (the PT tables are shared): the cost is one launch
    whatever the number of parameters, and the truncation error, O(epsilon^2) times the third derivative, is far
    below the 1e-5 target.  ``epsilon`` may be a scalar or one step per name; ``k`` and ``mu`` are paired points
    (or broadcast as in ``power``).  Constraints between parameters (rsdfit's ParameterSet) are the caller's
    chain rule."""
    names = [names] if isinstance(names, str) else list(names)
    for n in names:
        if n not in _GRAD_PARAMS:
            raise ValueError("no derivative of power(k, mu) with respect to `%s` (available: %s)" % (n, ', '.join(_GRAD_PARAMS)))
    eps = np.empty(len(names))
    eps[:] = epsilon
    if np.any(eps <= 0):
        raise ValueError("`epsilon` must be positive")
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    if k.ndim == 1 and mu.ndim == 1 and len(mu) != len(k):
        k, mu = k[:, None], mu[None, :]
    kb, mub = np.broadcast_arrays(k, mu)
    kmodel = self.k
    states = []
    for n, e in zip(names, eps):
        x0 = getattr(self, n)
        if x0 is None:
            raise ValueError("parameter `%s` is unset (derived from the linear spectrum): set it before differentiating" % n)
        try:
            for sgn in (+1., -1.):
                setattr(self, n, x0 + sgn*e)
                states.append(self._state(kmodel))
        finally:
            setattr(self, n, x0)
    P = self._evaluate(states, 'power', np.ascontiguousarray(kb).ravel(), np.ascontiguousarray(mub).ravel())[2]
    P = P.reshape(len(names), 2, -1)
    return (P[:, 0] - P[:, 1])/(2.*eps[:, None])


GalaxySpectrum.gradient = _gradient


def _components(self, k, mu):
    """(Pgal_cc, Pgal_cs, Pgal_ss)(k, mu): the central-central, central-satellite and satellite-satellite spectra of
    ``power = (1 - fs)^2 Pcc + 2 fs (1 - fs) Pcs + fs^2 Pss`` (power_gal.py:316-399, gal/__init__.py:24-40), each
    with the same AP treatment as ``power``.  Evaluated as ONE batch of three walkers of the assembly kernel at
    fs = 0, 1/2, 1 (fs enters nowhere else): Pcc = P(0), Pss = P(1), Pcs = 2 P(1/2) - (Pcc + Pss)/2."""
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    if k.ndim == 1 and mu.ndim == 1 and len(mu) != len(k):
        k, mu = k[:, None], mu[None, :]
    kb, mub = np.broadcast_arrays(k, mu)
    fs0 = self.fs
    states = []
    try:
        for v in (0.0, 0.5, 1.0):
            self.fs = v
            states.append(self._state(self.k))
    finally:
        self.fs = fs0
    P = self._evaluate(states, 'power', np.ascontiguousarray(kb).ravel(), np.ascontiguousarray(mub).ravel())[2]
    shape = kb.shape
    Pcc, Pss = P[0], P[2]
    Pcs = 2.0*P[1] - 0.5*(Pcc + Pss)
    return tuple(np.squeeze(x.reshape(shape)) for x in (Pcc, Pcs, Pss))


GalaxySpectrum.components = _components

_DM_ROWS = {'P_lin': 0, 'P00': 1, 'P01': 2, 'P11_mu4': 3, 'Pdv': 4, 'Pvv': 5, 'Pdd': 6}


def _dm_spectra(self):
    """the dark-matter building blocks on the model grid ``self.k`` at the current parameters: dict with
    P_lin (normed_power_lin), P00, P01, P11_mu4 (dm/P00.py, P01.py, P11.py totals), Pdd, Pdv, Pvv
    (power_dm.py:797-884 incl. the Jennings et al. fits of :676-718)"""
    kmodel = self.k
    eng, gal, _ = self._evaluate([self._state(kmodel)], 'power', kmodel[len(kmodel)//2:len(kmodel)//2 + 1], np.zeros(1))
    base = eng.galaxy_base(gal, 1)[0]
    return {name: base[row].copy() for name, row in _DM_ROWS.items()}


GalaxySpectrum.dm_spectra = _dm_spectra
GalaxySpectrum.Pgal_cc = lambda self, k, mu, flatten=False: _flat(_components(self, k, mu)[0], flatten)
GalaxySpectrum.Pgal_cs = lambda self, k, mu, flatten=False: _flat(_components(self, k, mu)[1], flatten)
GalaxySpectrum.Pgal_ss = lambda self, k, mu, flatten=False: _flat(_components(self, k, mu)[2], flatten)


def _flat(P, flatten):
    return np.ravel(P, order='F') if flatten else P


def _new_result():
    r = _Result(None)
    r.ncosmo = 0
    return r


def _z_setattr(self, name, value):
    """``redshift_params`` (power_biased.py:128-146): setting ``z`` re-derives ``f`` and / or ``sigma8_z`` from the cosmology"""
    object.__setattr__(self, name, value)
    if name == 'z' and 'cosmo' in self.__dict__:
        rp = getattr(self, 'redshift_params', ())
        if any(par not in ('f', 'sigma8_z') for par in rp):
            raise ValueError("valid parameters for redshift scaling: 'f' and 'sigma8_z'")
        if 'f' in rp:
            object.__setattr__(self, 'f', self.cosmo.f_z(value))
        if 'sigma8_z' in rp:
            object.__setattr__(self, 'sigma8_z', self.cosmo.Sigma8_z(value))


GalaxySpectrum.__setattr__ = _z_setattr

from . import dm as _dm_models  # noqa: E402,F401  (dm.py ends by importing model_api, which completes GalaxySpectrum)


def load_model(filename, show_warning=True):
    """a model saved with ``to_npy`` (``pyRSD.rsd.load_model``, rsd/__init__.py:39-60)"""
    return np.load(filename, allow_pickle=True).tolist()


_LAZY = {'DarkMatterSpectrum': 'dm', 'BiasedSpectrum': 'dm', 'HaloSpectrum': 'dm', 'QuasarSpectrum': 'qso',
         'ExtrapolatedPowerSpectrum': 'correlation', 'SmoothedXiMultipoles': 'correlation', 'transfers': 'transfers'}


def __getattr__(name):
    # the names ``pyRSD.rsd`` exports (rsd/__init__.py:21-29), resolved on first use
    if name in _LAZY:
        import importlib
        mod = importlib.import_module('.' + _LAZY[name], __package__)
        return mod if name == 'transfers' else getattr(mod, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
