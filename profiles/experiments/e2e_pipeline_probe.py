"""where does the pipelined end-to-end step lose time?  host enqueue time per submit, GPU time per step (events on the
library stream), and the throughput with results consumed late / never (pure enqueue)"""
import os, sys, time
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from pyrsd_b200 import _native as N, rsd
from pyrsd_b200.synthetic import synthetic_batch
B = 256
data = synthetic_batch(B, 0)
eng = rsd.PTEngine()
L = rsd._L()
sptr = C.c_void_p()
N.check(L.prsd_ctx_stream(eng.ctx.ptr, C.byref(sptr)))
stream = torch.cuda.ExternalStream(sptr.value, device=torch.device('cuda', 0))
for depth in (2, 3):
    st = rsd.PTStream(eng, data['kmodel'], rsd.galaxy_config(), B, data['k'].shape[1], data['kk'], data['mm'], depth=depth)
    for mode in ('consume one late', 'consume at the end'):
        for rep in range(2):
            steps = 10
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            host = []
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            evs[0].record(stream)
            tickets = []
            for i in range(steps):
                h0 = time.perf_counter()
                tickets.append(st.submit(data['k'], data['T'], data['pars'], data['wpar'], data['stoch']))
                host.append(time.perf_counter() - h0)
                evs[i + 1].record(stream)
                if mode == 'consume one late' and i >= 1:
                    st.result(tickets[i - 1])
            st.drain()
            wall = time.perf_counter() - t0
            torch.cuda.synchronize()
        gpu = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        print('depth %d, %-18s: wall %.2f ms/step; host submit ms %s; main-stream ms between step ends %s' %
              (depth, mode, 1e3*wall/steps, np.array2string(1e3*np.array(host), precision=2), np.array2string(np.array(gpu), precision=2)), flush=True)
