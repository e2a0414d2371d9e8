"""BASELINE configs[3]: 16384 FFTLog Hankel transforms on a 4096-point log grid (device-resident), CUDA-event timing"""
import sys, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np, torch
from pyrsd_b200 import _native as N, pygcl
L = pygcl._L()
vp = C.c_void_p
L.prsd_fftlog_device.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, vp, vp, C.POINTER(C.c_double)]
L.prsd_ctx_stream.argtypes = [vp, C.POINTER(vp)]
ctx = N.default_context(0)
sptr = vp(); N.check(L.prsd_ctx_stream(ctx.ptr, C.byref(sptr)))
stream = torch.cuda.ExternalStream(sptr.value, device=torch.device('cuda', 0))
for Nn, nb in ((4096, 16384), (1024, 65536)):
    dlnr = np.log(1e7)/Nn
    x = np.exp((np.arange(Nn) - Nn/2)*dlnr)
    sig = np.exp(np.random.default_rng(3).uniform(np.log(0.3), np.log(30.), nb))
    a = torch.from_numpy(x[None, :]**1.5*np.exp(-x[None, :]**2/(2*sig[:, None]**2))).cuda()
    out = torch.empty_like(a)
    kr = C.c_double()
    flush = torch.empty(64*1024*1024, dtype=torch.float32, device='cuda')
    for mu in (0.5,):
        for it in range(3):
            N.check(L.prsd_fftlog_device(ctx.ptr, nb, Nn, mu, 0.0, dlnr, 1.0, 1, a.data_ptr(), out.data_ptr(), C.byref(kr)))
        torch.cuda.synchronize()
        ts = []
        for it in range(5):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            N.check(L.prsd_fftlog_device(ctx.ptr, nb, Nn, mu, 0.0, dlnr, 1.0, 1, a.data_ptr(), out.data_ptr(), C.byref(kr)))
            e1.record(stream); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        gb = 16.0*Nn*nb/1e9
        print('N=%d batch=%d: %.3f ms  %.2f M transforms/s  %.1f GB/s of %d (HBM copy peak)  %.2f TFLOP/s (5 N log2 N + 3 N)' % (
            Nn, nb, ms, nb/ms/1e3, gb/(ms*1e-3), 6453, nb*(5*Nn*np.log2(Nn) + 3*Nn)/(ms*1e-3)/1e12))
