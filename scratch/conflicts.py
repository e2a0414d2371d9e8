import sys
sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from tests.host_emul import emul_lib
E = emul_lib.load()
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS'); dp = emul_lib.dp
E.L.emul_nodes.restype = C.c_longlong
E.L.emul_nodes.argtypes = [C.c_int, dp, C.c_double, C.c_longlong, ip, ip, dp]
E.L.emul_set_config_default()
def nodes(k, qmax, ml=False):
    E.L.emul_set_config_default()
    if ml: E.L.emul_use_ml_variant()
    cap = 200000
    o = np.empty(cap, np.int32); j = np.empty(cap, np.int32); w = np.empty(cap)
    n = E.L.emul_nodes(1, np.array([k]), qmax, cap, o, j, w)
    return o[:n].copy(), j[:n].copy()
def wavefronts(addr):
    """addr: array of 8-byte word addresses for consecutive lanes; returns total wavefronts over half-warps"""
    tot = 0
    n = len(addr)
    for i in range(0, n, 16):
        a = addr[i:i+16]
        ua = np.unique(a)
        banks = ua % 16
        tot += np.bincount(banks, minlength=16).max()
    return tot
def cost(o, j):
    # 4 table reads (same pattern) + 1 PaS read, per slot
    return 4*wavefronts(j) + wavefronts(o), 5*((len(j)+15)//16), wavefronts(j)/((len(j)+15)//16), wavefronts(o)/((len(j)+15)//16)
def permute(o, j):
    # items = nodes sharing a stencil start (broadcast reads); a half-warp takes whole items of distinct residues
    from collections import defaultdict, deque
    items = defaultdict(list)
    for idx in range(len(j)): items[j[idx]].append(idx)
    queues = [deque() for _ in range(16)]
    for j0 in sorted(items): queues[j0 % 16].append(deque(items[j0]))
    order = []
    remaining = len(j)
    start = 0
    while remaining:
        room = 16
        used = set()
        # residues with most remaining nodes first
        res = sorted(range(16), key=lambda r: -sum(len(it) for it in queues[r]))
        for r in res:
            if room == 0: break
            if not queues[r]: continue
            it = queues[r][0]
            take = min(room, len(it))
            for _ in range(take): order.append(it.popleft())
            if not it: queues[r].popleft()
            room -= take; remaining -= take
        # nothing else fits without a conflict: the half-warp stays partially conflicting only if room > 0 and nodes remain
        while room and remaining:
            r = max(range(16), key=lambda q: sum(len(it) for it in queues[q]))
            it = queues[r][0]
            take = min(room, len(it))
            for _ in range(take): order.append(it.popleft())
            if not it: queues[r].popleft()
            room -= take; remaining -= take
    return np.array(order)
for name, ks, qmax, ml in (('1loop', [1e-3, 0.03, 0.15, 0.6, 3.0], 1e5, False), ('ml', [1e-3, 0.03, 0.15, 0.6], 10., True)):
    for k in ks:
        o, j = nodes(k, qmax, ml)
        c0, ideal, t0, p0 = cost(o, j)
        p = permute(o, j)
        c1, _, t1, p1 = cost(o[p], j[p])
        print('%s k=%g nodes %d: wavefronts/ideal orig %.2f (table %.2f pa %.2f)  permuted %.2f (table %.2f pa %.2f)' % (name, k, len(j), c0/ideal, t0, p0, c1/ideal, t1, p1))
