"""Transfer functions between the theory P(k, mu) and what a survey measures, on the device.

Host-side mirror of ``pyRSD/rsd/transfers`` and ``pyRSD/rsd/window.py`` (same class names, constructor
arguments and call semantics); the per-walker arithmetic runs in ``libpyrsd_b200.so``:

* :class:`GriddedWedgeTransfer`, :class:`GriddedMultipoleTransfer` (``rsd/transfers/grid.py:9-319``): the mode
  counts, the mu-bin membership and the Legendre factors are folded once, here, into one weight array
  ``[nbin][Nk][Nmu]``; the device contracts it with every walker's P(k, mu) (``prsd_gridop_apply``).
* :class:`WindowConvolution` (``rsd/window.py:81-213``): coefficients of the Legendre product expansion and
  the kernel splines -- set-up, done here; :class:`WindowFunctionTransfer` (``rsd/transfers/window.py:9-132``):
  zero-padding, mcfit P2xi, the kernel mixing, mcfit xi2P and the k_out spline, batched over walkers on the
  device (``prsd_window_apply``, ``prsd_spline_resample_device``).

Inputs may be NumPy arrays (host; copied in and out) or CUDA ``torch`` tensors (used in place; a tensor comes
back).  A leading walker axis is optional: ``[Nk][Nmu]`` or ``[nw][Nk][Nmu]``.  The reference returns
``xarray.DataArray`` objects; these classes return plain arrays with the ``k`` / ``ell`` / ``mu`` coordinates as
attributes of the transfer object.  There is no CPU path.
"""
import ctypes as C
import warnings
from fractions import Fraction

import numpy as np

from . import _native as N

_vp = C.c_void_p
_registered = False


def _L():
    global _registered
    L = N.lib()
    if not _registered:
        i, d, pvp = C.c_int, C.c_double, C.POINTER(C.c_void_p)
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')
        ip = np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS')
        sig = {
            'prsd_gridop_create': (i, [_vp, i, i, i, dp, pvp]),
            'prsd_gridop_destroy': (i, [_vp]),
            'prsd_gridop_apply': (i, [_vp, i, _vp, _vp, i]),
            'prsd_window_create': (i, [_vp, i, dp, i, ip, d, i, pvp]),
            'prsd_window_destroy': (i, [_vp]),
            'prsd_window_info': (i, [_vp, ip, dp]),
            'prsd_window_set_kernel': (i, [_vp, dp]),
            'prsd_window_apply': (i, [_vp, i, _vp, _vp, i, i, _vp, _vp]),
            'prsd_spline_resample_device': (i, [_vp, i, i, dp, _vp, i, dp, _vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _registered = True
    return L


class _GridOp(N.Handle):
    _destroy = 'prsd_gridop_destroy'


class _Window(N.Handle):
    _destroy = 'prsd_window_destroy'


def _is_tensor(x):
    return type(x).__module__.split('.')[0] == 'torch'


def _ptr(x):
    return _vp(x.data_ptr()) if _is_tensor(x) else x.ctypes.data_as(_vp)


def _prepare(power, shape_tail):
    """-> (array or tensor with a leading walker axis, had_walker_axis, on_device)"""
    if _is_tensor(power):
        import torch
        if not power.is_cuda:
            raise ValueError("torch inputs must live on the GPU (pass NumPy arrays for host data)")
        p = power.to(torch.float64).contiguous()
        single = p.dim() == len(shape_tail)
        if single:
            p = p.unsqueeze(0)
        if tuple(p.shape[1:]) != tuple(shape_tail):
            raise ValueError("input of shape %s does not match the grid %s" % (tuple(power.shape), tuple(shape_tail)))
        return p, not single, True
    p = np.ascontiguousarray(power, dtype=np.float64)
    single = p.ndim == len(shape_tail)
    if single:
        p = p[None]
    if p.shape[1:] != tuple(shape_tail):
        raise ValueError("input of shape %s does not match the grid %s" % (np.shape(power), tuple(shape_tail)))
    return p, not single, False


def _empty_like(p, shape, on_device):
    if on_device:
        import torch
        return torch.empty(shape, dtype=torch.float64, device=p.device)
    return np.empty(shape)


class _CtxStream(object):
    """Run the enclosed torch operations on the library's (non-blocking) stream: they are then ordered with the
    library's kernels; the caller's current stream is joined on entry and on exit."""

    def __init__(self, ctx):
        import torch
        self.torch = torch
        sp = _vp()
        N.check(N.lib().prsd_ctx_stream(ctx.ptr, C.byref(sp)))
        self.dev = torch.device('cuda', ctx.device)
        self.ext = torch.cuda.ExternalStream(sp.value, device=self.dev)

    def __enter__(self):
        self.ext.wait_stream(self.torch.cuda.current_stream(self.dev))
        self.cm = self.torch.cuda.stream(self.ext)
        self.cm.__enter__()
        return self

    def __exit__(self, *exc):
        self.cm.__exit__(*exc)
        self.torch.cuda.current_stream(self.dev).wait_stream(self.ext)
        return False


class _NoStream(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class PkmuGrid(object):
    """A 2D grid of (k, mu): bin centres, mean k and mu of every cell and its number of modes
    (``rsd/transfers/__init__.py:5-82``)."""

    def __init__(self, coords, k, mu, modes):
        if np.shape(mu)[1] < 40:
            warnings.warn("initializing PkmuGrid with fewer mu bins than recommended; should have >= 40 to avoid inaccuracies")
        self.k_cen, self.mu_cen = [np.asarray(c, dtype=float) for c in coords]
        self.k = np.asarray(k, dtype=float)
        self.mu = np.asarray(mu, dtype=float)
        self.modes = np.asarray(modes, dtype=float)

    @property
    def shape(self):
        return self.k.shape

    @property
    def Nk(self):
        return self.shape[0]

    @property
    def Nmu(self):
        return self.shape[1]

    @property
    def notnull(self):
        return np.isfinite(self.modes) & (self.modes > 0)


def _legendre(ell, mu):
    c = np.zeros(ell + 1)
    c[ell] = 1.
    return np.polynomial.legendre.legval(mu, c)


class GriddedWedgeTransfer(object):
    """P(k, mu) on a grid -> wedges P(k, mu_bin): the mode-weighted mean over the grid cells of every mu bin
    (``rsd/transfers/grid.py:9-238``)."""

    def __init__(self, grid, mu_bounds, kmin=-np.inf, kmax=np.inf, context=None):
        self.grid = grid
        if not isinstance(mu_bounds, list) and isinstance(mu_bounds, tuple):
            mu_bounds = [mu_bounds]
        self.mu_cen = np.array([0.5*(lo + hi) for (lo, hi) in mu_bounds])
        self.k_cen = self.grid.k_cen
        # bins that end at mu = 1 include their upper edge (grid.py:52-55)
        self.mu_bounds = [(lo, 1.0005*hi) if np.isclose(hi, 1.0) else (lo, hi) for (lo, hi) in mu_bounds]
        self.mu_edges = np.asarray(self.mu_bounds).ravel()
        if not all(x <= y for x, y in zip(self.mu_edges, self.mu_edges[1:])):
            raise ValueError("specified `mu` bounds are not monotonically increasing")
        self.kmin = np.empty(self.N2)
        self.kmin[:] = -np.inf if kmin is None else kmin
        self.kmax = np.empty(self.N2)
        self.kmax[:] = np.inf if kmax is None else kmax
        self._ctx = context if context is not None else N.default_context()
        self._op = None

    @property
    def N1(self):
        return self.grid.Nk

    @property
    def N2(self):
        return len(self.mu_bounds)

    @property
    def in_k_range(self):
        k_cen = self.grid.k_cen
        return (k_cen[:, None] >= self.kmin) & (k_cen[:, None] <= self.kmax)

    @property
    def coords(self):
        return np.broadcast_arrays(self.k_cen[:, None], self.mu_cen[None, :])

    @property
    def flatcoords(self):
        return [x[self.in_k_range] for x in self.coords]

    @property
    def size(self):
        return len(self.flatcoords[0])

    def _membership(self):
        """[nbin][Nk][Nmu] booleans: grid cell in mu bin b (np.digitize on the ravelled bounds, every other bin,
        grid.py:143-166)"""
        g = self.grid
        dig = np.digitize(g.mu, self.mu_edges)
        return np.array([(dig == 2*b + 1) & g.notnull for b in range(len(self.mu_bounds))])

    def _extra(self):
        return np.ones((len(self.mu_bounds),) + self.grid.shape)

    def _weights(self):
        g = self.grid
        member = self._membership()
        modes = np.where(g.notnull, g.modes, 0.)
        w = member*modes[None]
        norm = w.sum(axis=-1)
        self._empty = norm == 0.                                 # NaN in the reference (0/0)
        with np.errstate(invalid='ignore', divide='ignore'):
            w = np.where(norm[..., None] > 0, w*self._extra()/norm[..., None], 0.)
        return np.ascontiguousarray(w)

    def _operator(self):
        if self._op is None:
            w = self._weights()
            if (self._empty & self.in_k_range.T).any():
                raise ValueError("NaN values in %s result within valid k range!" % self.__class__.__name__)
            p = _vp()
            N.check(_L().prsd_gridop_create(self._ctx.ptr, w.shape[0], w.shape[1], w.shape[2], w, C.byref(p)))
            self._op = _GridOp(p)
        return self._op

    def __call__(self, power):
        """``power`` [Nk][Nmu] or [nw][Nk][Nmu] -> [Nk][N2] or [nw][Nk][N2] (NaN outside [kmin, kmax] for host
        inputs, as the reference; device tensors are returned unmasked in the layout [nw][N2][Nk])"""
        op = self._operator()
        p, multi, dev = _prepare(power, self.grid.shape)
        nw = p.shape[0]
        with (_CtxStream(self._ctx) if dev else _NoStream()):
            out = _empty_like(p, (nw, self.N2, self.grid.Nk), dev)
            N.check(_L().prsd_gridop_apply(op.ptr, nw, _ptr(p), _ptr(out), int(dev)))
        if dev:
            return out if multi else out[0]
        out = np.ascontiguousarray(np.swapaxes(out, 1, 2))
        out[:, ~self.in_k_range] = np.nan
        return out if multi else out[0]


class GriddedMultipoleTransfer(GriddedWedgeTransfer):
    """P(k, mu) on a grid -> multipoles: the mode-weighted mean of (2 ell + 1) L_ell(mu) P over mu in [0, 1]
    (``rsd/transfers/grid.py:241-319``)."""

    def __init__(self, grid, ells, kmin=-np.inf, kmax=np.inf, context=None):
        self.ells = np.array(ells, ndmin=1)
        GriddedWedgeTransfer.__init__(self, grid, [(0., 1.)], kmin=kmin, kmax=kmax, context=context)

    @property
    def N2(self):
        return len(self.ells)

    @property
    def coords(self):
        return np.broadcast_arrays(self.k_cen[:, None], self.ells[None, :])

    @property
    def legendre_weights(self):
        return np.array([(2*ell + 1)*_legendre(ell, self.grid.mu) for ell in self.ells])

    def _membership(self):
        m = GriddedWedgeTransfer._membership(self)
        return np.repeat(m, len(self.ells), axis=0)

    def _extra(self):
        return np.where(self.grid.notnull, self.legendre_weights, 0.)


# ------------------------------------------------------------------------------------- window function
def _poch_half(p):
    """(1/2)_p 2^p / p! = (2p - 1)!! / p!  -- the G(p) of Wilson et al. 2017 as an exact fraction"""
    r = Fraction(1)
    for j in range(1, p + 1):
        r *= Fraction(2*j - 1, j)
    return r


def get_coefficients(ell, ellprime):
    """L_ell L_ellprime = sum_q c_q L_q expansion weighted by (2 ell + 1)/(2 q + 1): returns (q values, coefficients)
    in increasing q (``rsd/window.py:29-79``; Wilson et al. 2017 eq. 22-25)."""
    qs, cs = [], []
    for p in range(0, min(ell, ellprime) + 1):
        q = ell + ellprime - 2*p
        c = _poch_half(ell - p)*_poch_half(p)*_poch_half(ellprime - p)/_poch_half(ell + ellprime - p)
        c *= Fraction(2*(ell + ellprime) - 4*p + 1, 2*(ell + ellprime) - 2*p + 1)
        c *= Fraction(2*ell + 1, 2*q + 1)
        qs.append(q)
        cs.append(float(c))
    return qs[::-1], cs[::-1]


class WindowConvolution(object):
    """The window-function kernels of Wilson et al. 2017 (``rsd/window.py:81-213``): for a convolved multipole
    ``ell`` and a leaking multipole ``ellprime`` the kernel is a fixed linear combination of the measured window
    multipoles, splined in the separation; outside the measured range it is the identity."""

    def __init__(self, s, W, max_ellprime=4, max_ell=4):
        from scipy.interpolate import InterpolatedUnivariateSpline
        self.s = np.asarray(s, dtype=float)
        self.smin, self.smax = self.s.min(), self.s.max()
        self.W = np.asarray(W, dtype=float)
        self.max_ellprime, self.max_ell = max_ellprime, max_ell
        self.splines = {}
        for ell in range(0, max_ell + 1, 2):
            kern = []
            for ellprime in range(0, max_ellprime + 1, 2):
                qvals, coeffs = get_coefficients(ell, ellprime)
                if max(qvals)//2 >= self.W.shape[1]:
                    raise ValueError("window multipoles up to ell = %d are needed (%d columns given)" % (max(qvals), self.W.shape[1]))
                kern.append(sum(c*self.W[:, q//2] for q, c in zip(qvals, coeffs)))
            self.splines[ell] = [InterpolatedUnivariateSpline(self.s, kk) for kk in kern]

    def _get_kernel(self, ell, r):
        splines = self.splines[ell]
        toret = np.zeros((len(r), len(splines)))
        idx = (r >= self.smin) & (r <= self.smax)
        for i, spl in enumerate(splines):
            toret[idx, i] = spl(r[idx])
            if i == ell//2:
                toret[~idx, i] = 1.0
        return toret

    def kernel_matrix(self, ells, r):
        """[nell][npoles][len(r)], npoles = max_ellprime/2 + 1"""
        return np.array([self._get_kernel(ell, r).T for ell in ells])

    def __call__(self, ells, r, xi, order='F'):
        """host evaluation (used for set-up checks only; the product path is WindowFunctionTransfer)"""
        npoles = self.max_ellprime//2 + 1
        xi = np.asarray(xi)
        if xi.shape[1] > npoles:
            xi = xi[..., :npoles]
        elif xi.shape[1] < npoles:
            raise ValueError("shape mismatch between kernel and number of xi multipoles; "
                             "please provide the first %d even multipoles" % npoles)
        return np.asarray([np.einsum('ij,ij->i', xi, self._get_kernel(ell, r)) for ell in ells]).T.copy(order=order)


def convolve_multipoles(k, Pell, ells, convolver, qbias=0.7, dry_run=False, legacy=True):
    """Window-convolved power multipoles, the older route of ``rsd/window.py:214-307`` (``WindowFunctionTransfer`` is the
    current one).  ``Pell`` [Nk][Nell] on ``k``; ``convolver`` a ``WindowConvolution``.

    * ``legacy=True``: multipoles with fewer than 500 k are first splined to 500 log-spaced points (FITPACK, as the
      reference); xi_ell on the window's own separations by the trapezoid rule, the window kernels, and back to the input
      k by the trapezoid rule -- the transforms are ``pygcl.pk_to_xi`` / ``xi_to_pk`` with ``method=TRAPZ`` (device).
      Returns ``(k, Pell_conv)`` with Pell_conv [Nk][Nell].
    * ``legacy=False``: biased FFTLog there and back on the input (log-spaced) ``k`` (``pygcl.ComputeXiLM_fftlog`` with
      q = +-qbias).  Returns ``(k, Pell_conv)``."""
    from . import pygcl
    k = np.asarray(k, dtype=float)
    Pell = np.asarray(Pell, dtype=float)
    ells = [int(ell) for ell in ells]
    sign = [(-1.)**(ell//2) for ell in ells]
    if not legacy:
        fwd = [pygcl.ComputeXiLM_fftlog(ell, 2, k, Pell[:, i], qbias) for i, ell in enumerate(ells)]
        # Reference behaviour kept: every multipole's transform has its own low-ringing kr, hence its own output grid, but
        # rsd/window.py:236-251 lets each call overwrite ONE `rr` (and `kk`) array -- the kernels are evaluated on, and the
        # backward transforms start from, the grid of the LAST multipole (the grids differ by a factor ~1 +- 4e-4)
        rr = fwd[-1][0]
        xi = np.asfortranarray(np.array([sg*f[1] for sg, f in zip(sign, fwd)]).T)
        xi_conv = xi.copy() if dry_run else convolver(ells, rr, xi, order='F')
        back = [pygcl.ComputeXiLM_fftlog(ell, 2, rr, np.ascontiguousarray(xi_conv[:, i]), -qbias) for i, ell in enumerate(ells)]
        return back[-1][0], np.asfortranarray(np.array([sg*(2*np.pi)**3*b[1] for sg, b in zip(sign, back)]).T)

    if len(ells) != Pell.shape[-1]:
        raise ValueError("shape mismatch between multipole numbers and number of multipoles provided")
    if not all(ell in (0, 2, 4, 6) for ell in ells):
        raise ValueError("valid `ell` values are [0,2,4,6]")
    s = convolver.s
    k_out = k if k.ndim == 2 else np.repeat(k[:, None], len(ells), axis=1)
    if k_out.shape[-1] != len(ells):
        raise ValueError("input `k_out` must have %d columns for ell=%s multipoles" % (len(ells), str(ells)))
    if len(k) < 500:
        from scipy import interpolate as _interp
        k_hi = np.logspace(np.log10(k.min()), np.log10(k.max()), 500)
        Pell = np.array([_interp.splev(k_hi, _interp.splrep(k, Pell[:, i], k=3, s=0)) for i in range(len(ells))]).T
        k = k_hi
    TRAPZ = pygcl.IntegrationMethods.TRAPZ
    xi = np.array([pygcl.pk_to_xi(ell, k, np.ascontiguousarray(Pell[:, i]), s, smoothing=0., method=TRAPZ)
                   for i, ell in enumerate(ells)]).T
    xi_conv = xi.copy() if dry_run else convolver(ells, s, xi, order='F')
    conv = np.array([pygcl.xi_to_pk(ell, s, np.ascontiguousarray(xi_conv[:, i]), np.ascontiguousarray(k_out[:, i]), smoothing=0.,
                                    method=TRAPZ) for i, ell in enumerate(ells)]).T
    return k_out[:, 0], conv


class WindowFunctionTransfer(GriddedMultipoleTransfer):
    """Unconvolved P(k, mu) on the internal grid -> window-convolved multipoles
    (``rsd/transfers/window.py:9-132``).

    ``window``: array whose first column is the separation ``s`` and the others the even window multipoles
    Q_0, Q_2, ...; the remaining arguments as in the reference."""

    def __init__(self, window, ells, kmin=1e-4, kmax=0.7, Nk=1024, Nmu=40, max_ellprime=4, context=None):
        k = np.logspace(np.log10(kmin), np.log10(kmax), Nk)
        mu_edges = np.linspace(0., 1., Nmu + 1)
        mu = 0.5*(mu_edges[1:] + mu_edges[:-1])
        grid_k, grid_mu = np.meshgrid(k, mu, indexing='ij')
        grid = PkmuGrid([k, mu], grid_k, grid_mu, np.ones_like(grid_k))
        # the reference passes (kmin, kmax) on, which blanks an end point whenever logspace rounds it outside the
        # interval and then poisons the FFT with NaN; the grid is the range here
        GriddedMultipoleTransfer.__init__(self, grid, ells, context=context)
        window = np.asarray(window, dtype=float)
        self.convolver = WindowConvolution(window[:, 0], window[:, 1:], max_ellprime=max_ellprime, max_ell=max(self.ells))
        ells_l = [int(e) for e in self.ells]
        npoles = max_ellprime//2 + 1
        if len(ells_l) < npoles or ells_l[:npoles] != list(range(0, max_ellprime + 1, 2)):
            raise ValueError("shape mismatch between kernel and number of xi multipoles; "
                             "please provide the first %d even multipoles" % npoles)
        # zero-padding in k up to 100 h/Mpc with the same log spacing (window.py:84-93)
        dk = np.diff(np.log10(k))[0]
        newk = 10**(np.arange(np.log10(k.max()) + dk, 2 + 0.5*dk, dk))
        self.k_pad = np.concatenate([k, newk])
        self._win = None

    def _window(self):
        if self._win is None:
            p = _vp()
            ells = np.ascontiguousarray(self.ells, dtype=np.int32)
            N.check(_L().prsd_window_create(self._ctx.ptr, len(self.k_pad), self.k_pad, len(ells), ells, 1.5, 0, C.byref(p)))
            self._win = _Window(p)
            info = np.zeros(4, dtype=np.int32)
            self.rr = np.empty(len(self.k_pad))
            N.check(_L().prsd_window_info(p, info, self.rr))
            self.N_fft = int(info[0])
            nell, npoles = len(ells), self.convolver.max_ellprime//2 + 1
            kern = np.zeros((nell, nell, len(self.k_pad)))
            kern[:, :npoles] = self.convolver.kernel_matrix(ells, self.rr)      # columns beyond npoles are dropped
            N.check(_L().prsd_window_set_kernel(p, np.ascontiguousarray(kern)))
        return self._win

    def __call__(self, power, k_out=None, extrap=False, dry_run=False, no_convolution=False, return_xi=False):
        """``power`` [Nk][Nmu] or [nw][Nk][Nmu] on ``self.grid`` -> convolved multipoles [len(k)][nell] (or with a
        leading walker axis) on ``k_out`` if given, else on the zero-padded grid ``self.k_pad``; device tensors
        come back as [nw][nell][len(k)]"""
        if extrap:
            # the reference passes `extrap` to mcfit, whose power-law padding is f[-1] (f[-1] / f[-2])^j (extern/mcfit/
            # mcfit.py:170-178); this transfer zero-pads P_ell from kmax to 100 h/Mpc first (transfers/window.py:84-92), so
            # f[-1] / f[-2] = 0 / 0 and every output of the reference is NaN (checked: profiles/r02_summary.md).  There is
            # nothing to be equal to.
            raise NotImplementedError("extrap=True is not meaningful for WindowFunctionTransfer: the reference's power-law "
                                      "padding divides 0 / 0 on the zero-padded high-k tail and returns NaN everywhere")
        op = self._operator()
        p, multi, dev = _prepare(power, self.grid.shape)
        if return_xi and dev:
            raise ValueError("return_xi is only available for host (NumPy) inputs: the correlation functions are copied out "
                             "through host buffers")
        nw, nell, Nk, Npad = p.shape[0], len(self.ells), self.grid.Nk, len(self.k_pad)
        with (_CtxStream(self._ctx) if (dev or k_out is not None) else _NoStream()):
            poles = _empty_like(p, (nw, nell, Nk), dev)
            N.check(_L().prsd_gridop_apply(op.ptr, nw, _ptr(p), _ptr(poles), int(dev)))
            # copy over with zeros
            padded = _empty_like(p, (nw, nell, Npad), dev)
            padded[..., :Nk] = poles
            padded[..., Nk:] = 0.
            xi = xic = None
            if no_convolution:
                conv = padded
            else:
                win = self._window()
                conv = _empty_like(p, (nw, nell, Npad), dev)
                if return_xi and not dev:
                    xi, xic = np.empty((nw, nell, Npad)), np.empty((nw, nell, Npad))
                N.check(_L().prsd_window_apply(win.ptr, nw, _ptr(padded), _ptr(conv), 1 if dry_run else 0, int(dev),
                                               None if xi is None else _ptr(xi), None if xic is None else _ptr(xic)))
            if k_out is not None:
                k_out = N.as_f64(k_out)
                if dev:
                    res = _empty_like(p, (nw, nell, len(k_out)), True)
                    N.check(_L().prsd_spline_resample_device(self._ctx.ptr, nw*nell, Npad, self.k_pad, _ptr(conv), len(k_out),
                                                             k_out, _ptr(res)))
                else:
                    import torch
                    d_in = torch.from_numpy(conv).to('cuda:%d' % self._ctx.device)
                    d_res = torch.empty((nw, nell, len(k_out)), dtype=torch.float64, device=d_in.device)
                    N.check(_L().prsd_spline_resample_device(self._ctx.ptr, nw*nell, Npad, self.k_pad, _ptr(d_in), len(k_out),
                                                             k_out, _ptr(d_res)))
                    res = d_res.cpu().numpy()
            else:
                res = conv
        if dev:
            return res if multi else res[0]
        res = np.ascontiguousarray(np.swapaxes(res, 1, 2))
        if return_xi:
            return (res if multi else res[0]), xi, xic
        return res if multi else res[0]


# ------------------------------------------------------------------------- continuous-model transfers
def simpson_weights(x):
    """weights w with ``SimpsIntegrate(x, y) == w @ y`` for an ODD number of samples (DiscreteQuad.cpp:3-21, 40-66):
    per pair of intervals (h0, h1): (h0 + h1)/6 * [2 - h1/h0, (h0 + h1)^2/(h0 h1), 2 - h0/h1]"""
    x = np.asarray(x, dtype=float)
    if len(x) % 2 == 0 or len(x) < 3:
        raise ValueError("Simpson weights need an odd number (>= 3) of samples")
    h = np.diff(x)
    h0, h1 = h[0::2], h[1::2]
    hsum = h0 + h1
    w = np.zeros(len(x))
    w[0:-2:2] += hsum/6.*(2. - h1/h0)
    w[1:-1:2] += hsum/6.*hsum*hsum/(h0*h1)
    w[2::2] += hsum/6.*(2. - h0/h1)
    return w


class MultipoleTransfer(GriddedMultipoleTransfer):
    """P(k, mu) sampled on ``k`` x ``linspace(0, 1, Nmu)`` -> P_ell(k) = (2 ell + 1) int_0^1 L_ell P dmu by Simpson's
    rule (``rsd/transfers/poles.py:8-69``: ``pygcl.SimpsIntegrate`` row by row); Nmu is made odd as the reference
    does.  ``GalaxySpectrum.poles`` fuses the same rule behind the assembly kernel; this class serves P(k, mu) arrays
    that already exist."""

    def __init__(self, k, ells, Nmu=40, context=None):
        if Nmu % 2 == 0:
            Nmu += 1
        k = np.asarray(k, dtype=float)
        mu = np.linspace(0., 1., Nmu, endpoint=True)
        grid_k, grid_mu = np.meshgrid(k, mu, indexing='ij')
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            grid = PkmuGrid([k, mu], grid_k, grid_mu, np.ones_like(grid_k))
        GriddedMultipoleTransfer.__init__(self, grid, ells, context=context)

    def _weights(self):
        w = simpson_weights(self.grid.mu_cen)
        self._empty = np.zeros((len(self.ells), self.grid.Nk), dtype=bool)
        return np.ascontiguousarray(self.legendre_weights*w[None, None, :])


class WedgeTransfer(GriddedWedgeTransfer):
    """P(k, mu) sampled on ``k`` x ``linspace(0, 1, Nmu + 1)`` -> the plain mean over the samples of every wide mu
    bin (``rsd/transfers/wedges.py:7-108``)."""

    def __init__(self, k, mu_bounds, Nmu=40, context=None):
        k = np.asarray(k, dtype=float)
        mu = np.linspace(0., 1., Nmu + 1, endpoint=True)
        grid_k, grid_mu = np.meshgrid(k, mu, indexing='ij')
        grid = PkmuGrid([k, mu], grid_k, grid_mu, np.ones_like(grid_k))
        GriddedWedgeTransfer.__init__(self, grid, mu_bounds, context=context)
