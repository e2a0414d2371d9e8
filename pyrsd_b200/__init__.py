"""pyrsd_b200 -- B200-native (sm_100a) implementation of the perturbation-theory hot path of
nickhand/pyRSD behind the reference's own ``pygcl`` / ``rsd`` Python API.

The compute path is hand-written CUDA reached through a flat C-ABI
(``include/pyrsd_b200.h`` -> ``pyrsd_b200/libpyrsd_b200.so``); there is no CPU fallback.
"""
__version__ = '0.1.0'
