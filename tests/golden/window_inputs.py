"""Synthetic inputs shared by tests/golden/make_window_ref.py and the transfer-function tests."""
import numpy as np


def synthetic_power(k, mu, b=1.9, f=0.78, sigma=3.8):
    """Kaiser + Lorentzian FoG on a smooth spectrum with a BAO-like wiggle"""
    k, mu = np.broadcast_arrays(k, mu)
    P0 = 2.1e4*(k/0.02)/(1. + (k/0.02)**2.55)*(1. + 0.04*np.sin(k/0.0095)*np.exp(-(k/0.25)**2))
    return (b + f*mu**2)**2*P0/(1. + 0.5*(k*mu*sigma)**2)**2 + 310.


def synthetic_window(ns=320):
    s = np.logspace(0., np.log10(3400.), ns)
    x = s/900.
    Q0 = np.exp(-x**2)*(1. - 0.12*x)
    Q2 = -0.21*x*np.exp(-0.8*x**2)
    Q4 = 0.09*x**2*np.exp(-0.9*x**2)
    Q6 = -0.03*x**3*np.exp(-x**2)
    Q8 = 0.012*x**4*np.exp(-1.1*x**2)
    return np.column_stack([s, Q0, Q2, Q4, Q6, Q8])
