"""what does the per-step exchange of the multi-GPU bench cost?  grid-reduction kernel and NCCL all-gather timed apart
(torchrun, one process per GPU)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
from pyrsd_b200 import _native as N, transfers
L = N.lib(); transfers._L()
ctx = N.default_context(local)
sptr = C.c_void_p(); N.check(L.prsd_ctx_stream(ctx.ptr, C.byref(sptr)))
stream = torch.cuda.ExternalStream(sptr.value, device=dev)
B, NK, NMU = 256, 1000, 40
gop = C.c_void_p()
N.check(L.prsd_gridop_create(ctx.ptr, 1, NK, NMU, np.full((1, NK, NMU), 1.0/NMU), C.byref(gop)))
d_out = torch.rand((B, NK*NMU), dtype=torch.float64, device=dev)
d_send = torch.empty((B, NK), dtype=torch.float64, device=dev)
d_recv = torch.empty((world*B, NK), dtype=torch.float64, device=dev)
flush = torch.empty(64*1024*1024, dtype=torch.float32, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot/n
def grid(): N.check(L.prsd_gridop_apply(gop, B, d_out.data_ptr(), d_send.data_ptr(), 1))
def gather():
    with torch.cuda.stream(stream):
        dist.all_gather_into_tensor(d_recv, d_send)
def both(): grid(); gather()
r = [timeit(grid), timeit(gather), timeit(both)]
if rank == 0:
    print('world %d: grid reduction %.3f ms, all-gather %.3f ms, both %.3f ms' % (world, r[0], r[1], r[2]), flush=True)
dist.destroy_process_group()
