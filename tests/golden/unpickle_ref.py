"""Stub-unpickler for the reference's shipped model pickle.

Reads ``docs/data/galaxy_power.npy`` of the reference (a numpy object array
wrapping a Python-2 era pickle of a fully initialised ``GalaxySpectrum``)
WITHOUT importing pyRSD / astropy / scipy-era classes: every unknown class is
replaced by a dict-like stub that just records its pickled state.
This is test infrastructure (golden-vector generation), never product code.
"""
import io
import pickle
import types

import numpy as np


class _Placeholder(object):
    def __init__(self, owner, name):
        self.owner, self.name = owner, name

    def __call__(self, *a, **k):
        return None

    def __reduce__(self):
        return (_Placeholder, (None, self.name))


_STUB_CACHE = {}


def _make_stub(module, name):
    key = (module, name)
    if key in _STUB_CACHE:
        return _STUB_CACHE[key]

    class Stub(dict):
        _ref_class = "%s.%s" % (module, name)

        def __init__(self, *args, **kwargs):
            dict.__init__(self)
            self._args = args
            self._kwargs = kwargs
            self._state = None
            self._items = []

        def __setstate__(self, state):
            self._state = state

        def append(self, x):
            self._items.append(x)

        def extend(self, xs):
            self._items.extend(xs)

        def __getattr__(self, item):
            if item.startswith('__') and item.endswith('__'):
                raise AttributeError(item)
            return _Placeholder(self, item)

        def __repr__(self):
            return "<Stub %s>" % self._ref_class

        def __reduce_ex__(self, protocol):
            raise TypeError("stubs are read-only")

    Stub.__name__ = name
    _STUB_CACHE[key] = Stub
    return Stub


class _TolerantArray(np.ndarray):
    """ndarray subclass accepting the longer state tuples of astropy Quantity."""

    def __setstate__(self, state):
        # astropy Quantity pickles (ndarray_state, extra_dict)
        if len(state) == 2 and isinstance(state[0], tuple):
            state = state[0]
        np.ndarray.__setstate__(self, tuple(state[:5]))


_STUB_ROOTS = ('pyRSD', 'scipy', 'george', 'sklearn', 'pandas', 'astropy',
               'xarray', 'lmfit', 'emcee')


class RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith('numpy.core'):
            module = module.replace('numpy.core', 'numpy._core', 1)
        root = module.split('.')[0]
        if root == 'astropy' and name == 'Quantity':
            return _TolerantArray
        if root in _STUB_ROOTS:
            return _make_stub(module, name)
        if module == '__builtin__':
            module = 'builtins'
        if module == 'copy_reg':
            module = 'copyreg'
        return super().find_class(module, name)


def load_reference_model(path):
    with open(path, 'rb') as f:
        version = np.lib.format.read_magic(f)
        assert version == (1, 0), version
        np.lib.format.read_array_header_1_0(f)
        up = RefUnpickler(f, encoding='latin1')
        return up.load()
