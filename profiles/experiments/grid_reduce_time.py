"""device time of the grid-reduction kernel on the benchmark's shape (256 x 1000 x 40), inputs alternating between two
82 MB arrays (> L2 together), 40 back-to-back launches between one event pair"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pyrsd_b200 import _native as N, transfers
L = transfers._L()
ctx = N.default_context(0)
dev = torch.device('cuda', 0)
sptr = C.c_void_p(); N.check(N.lib().prsd_ctx_stream(ctx.ptr, C.byref(sptr)))
stream = torch.cuda.ExternalStream(sptr.value, device=dev)
B, NK, NMU = 256, 1000, 40
gop = C.c_void_p()
N.check(L.prsd_gridop_create(ctx.ptr, 1, NK, NMU, np.full((1, NK, NMU), 1.0/NMU), C.byref(gop)))
a = [torch.rand((B, NK*NMU), dtype=torch.float64, device=dev) for _ in range(2)]
out = torch.empty((B, NK), dtype=torch.float64, device=dev)
def run(n):
    for i in range(n):
        N.check(L.prsd_gridop_apply(gop, B, C.c_void_p(a[i & 1].data_ptr()), C.c_void_p(out.data_ptr()), 1))
run(4); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(40); e1.record(stream); torch.cuda.synchronize()
print(os.environ.get('PRSD_GRID_STAGED', 'rows'), '%.4f ms per launch' % (e0.elapsed_time(e1)/40))
