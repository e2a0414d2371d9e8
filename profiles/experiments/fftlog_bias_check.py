"""biased FFTLog (q != 0) on the device vs the C restatement of the Fortran (oracle): forward and backward, l = 0, 2, 4"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from pyrsd_b200 import pygcl
from oracle import gcl_ref as R
k = np.logspace(-4, 1, 1024)
P = 1e4*(k/0.02)/(1 + (k/0.02)**2.5)
for N in (1024, 700):
    kk = np.logspace(-4, 1, N)
    PP = 1e4*(kk/0.02)/(1 + (kk/0.02)**2.5)
    for l in (0, 2, 4):
        for q in (0.0, 0.7, -0.7, 0.3):
            r1, x1 = pygcl.ComputeXiLM_fftlog(l, 2, kk, PP, q, 0.)
            r2, x2 = R.ComputeXiLM_fftlog(l, 2, kk, PP, q, 0.)
            print(N, l, q, 'r %.1e' % np.max(np.abs(r1/r2 - 1)), 'xi %.2e' % (np.max(np.abs(x1 - x2))/np.max(np.abs(x2))),
                  'ratio mid %.6f' % (x1[N//2]/x2[N//2]))
