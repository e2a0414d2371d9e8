"""The C-ABI shared library loads on a CPU-only machine and exports every symbol declared in
include/pyrsd_b200.h; without a GPU the context constructor fails LOUDLY (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'pyrsd_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(prsd_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    from pyrsd_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        from pyrsd_b200 import build
        build.build_product()
    return _native.lib()


def test_header_declares_the_boundary():
    names = declared_symbols()
    assert len(names) >= 35
    for must in ('prsd_eval_imn', 'prsd_eval_jmn', 'prsd_eval_kmn', 'prsd_eval_imn_oneloop', 'prsd_oneloop_tables',
                 'prsd_zeldovich', 'prsd_fftlog', 'prsd_compute_xilm', 'prsd_plan_run', 'prsd_galaxy_power'):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_oracle_in_product():
    """the product package must never import or link the oracle / the CPU emulation"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'pyrsd_b200')):
        if os.path.basename(dirpath) == 'build':
            continue
        for f in files:
            if f == 'build.py':       # builds the checker too (building the oracle is not using it)
                continue
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                # strip comments / docstrings mentioning the test infrastructure
                code = re.sub(r'//.*', '', txt)
                assert not re.search(r'(import|from|#include|CDLL|dlopen)[^\n]*(gcl_ref|host_emul|libprsd_emul|oracle)', code), f


def test_context_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = C.c_void_p()
    rc = lib.prsd_ctx_create(0, C.byref(p))
    assert rc == -3                      # PRSD_ERR_NO_DEVICE
    assert b'no CPU fallback' in lib.prsd_last_error()


def test_static_queries_work_without_gpu(lib):
    from pyrsd_b200 import _native
    x = _native.knot_grid()
    assert len(x) == 2457
    cfg = _native.quad_config_default()
    assert len(cfg) == 56 and cfg[0] == 1e-8 and cfg[21] == 12 and cfg[25] == 16 and cfg[29] == 12 and cfg[41] == 10 and cfg[42] == 7 and cfg[45] == 10 and cfg[46] == 9 and cfg[49] == 10 and cfg[50] == 8


def test_rsd_module_exports_the_reference_names():
    """``from pyRSD.rsd import ...`` (rsd/__init__.py:21-29) works against ``pyrsd_b200.rsd``; every model class answers the
    reference's public methods that do not need a device to exist"""
    from pyrsd_b200 import rsd
    for name in ('GalaxySpectrum', 'DarkMatterSpectrum', 'BiasedSpectrum', 'HaloSpectrum', 'QuasarSpectrum',
                 'ExtrapolatedPowerSpectrum', 'SmoothedXiMultipoles', 'transfers', 'load_model'):
        assert getattr(rsd, name) is not None
    for cls in (rsd.GalaxySpectrum, rsd.BiasedSpectrum, rsd.DarkMatterSpectrum, rsd.QuasarSpectrum):
        for meth in ('power', 'poles', 'monopole', 'quadrupole', 'hexadecapole', 'derivative_k', 'derivative_mu', 'apply_transfer'):
            assert callable(getattr(cls, meth)), (cls.__name__, meth)
    for meth in ('Pgal_cAcA', 'Pgal_cBsA', 'Pgal_sBsB', 'Pgal_cc', 'preserve', 'use_spt', 'cache_override', 'to_npy', 'from_npy'):
        assert hasattr(rsd.GalaxySpectrum, meth), meth
    with pytest.raises(AttributeError):
        rsd.no_such_name
