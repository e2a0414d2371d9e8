"""Walker / cosmology sharding over the GPUs of one box.

The reference distributes one likelihood evaluation per MPI rank through ``emcee.utils.MPIPool``
(``pyRSD/rsdfit/util/mpi_manager.py``, ``rsdfit/engines/mcmc.py``): walkers are independent, so here every rank
(one process per GPU, ``torch.distributed``) owns a contiguous block of walkers, runs the same batched kernels on
its block, and the per-walker results are gathered once per ensemble step -- NCCL over NVLink on GPUs, gloo on
CPU tensors (tests).  There is no collective on the data path itself.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """contiguous block [lo, hi) of rank `rank`: the first n % world ranks hold one extra item"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %d of %d" % (rank, world))
    base, extra = divmod(int(n), world)
    lo = rank*base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world):
    return [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]


def gather_rows(local, n_total, group=None):
    """every rank contributes its block of rows ``local`` [n_local, ...]; returns [n_total, ...] in walker order on
    every rank.  Blocks may be ragged (n_total not a multiple of the world size): they are padded to the largest
    block for the fixed-size all-gather and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_total:
            raise ValueError("single process holds %d of %d rows" % (local.shape[0], n_total))
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, its block has %d" % (rank, local.shape[0], sizes[rank]))
    nmax = max(sizes)
    tail = tuple(local.shape[1:])
    send = local
    if local.shape[0] != nmax:
        send = torch.zeros((nmax,) + tail, dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    recv = torch.empty((world*nmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(s == nmax for s in sizes):
        return recv
    recv = recv.view((world, nmax) + tail)
    return torch.cat([recv[r, :sizes[r]] for r in range(world)], dim=0)


def max_over_ranks(value, device=None, group=None):
    """max of a python float over ranks (timings are reported as the max over ranks)"""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class _DeviceArray(object):
    """a device pointer as a ``__cuda_array_interface__`` holder (for ``torch.as_tensor``)"""
    def __init__(self, ptr, shape, owner):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}
        self._owner = owner


class PeerExchange(object):
    """The per-step exchange over NVLink / NVSwitch peer memory (``prsd_exchange_*``, csrc/exchange.cu): every rank's
    gridded-transfer kernel writes its per-walker results into its slot of its own receive buffer, a push kernel stores the
    slot into every peer's buffer, a flag exchange follows; no library collective on the data path.  ``torch.distributed`` carries the 128 handle bytes once at set-up.

    ``n_per_rank`` doubles per rank and step (walkers x bins x k of the gridded transfer)."""

    def __init__(self, context, n_per_rank, group=None):
        import ctypes as C
        import numpy as np
        from . import _native as N
        from . import transfers as _T
        self._C, self._N, self._L = C, N, _T._L()
        L = self._L
        vp = C.c_void_p
        L.prsd_exchange_create.argtypes = [vp, C.c_int, C.c_int, C.c_longlong, C.POINTER(vp)]
        L.prsd_exchange_destroy.argtypes = [vp]
        L.prsd_exchange_handles.argtypes = [vp, C.c_void_p]
        L.prsd_exchange_open.argtypes = [vp, C.c_void_p]
        L.prsd_gridop_apply_exchange.argtypes = [vp, C.c_int, vp, vp]
        L.prsd_exchange_received.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int)]
        for f in ('prsd_exchange_create', 'prsd_exchange_destroy', 'prsd_exchange_handles', 'prsd_exchange_open',
                  'prsd_gridop_apply_exchange', 'prsd_exchange_received'):
            getattr(L, f).restype = C.c_int
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.n_per_rank = int(n_per_rank)
        self._ctx = context
        self.ptr = vp()
        # every phase that can fail on one rank alone is followed by an agreement, so that the ranks never sit in different
        # collectives: either every rank owns a mapped exchange after this constructor or every rank raises
        tdev = torch.device('cpu')
        if on and dist.get_backend(group) == 'nccl':
            tdev = torch.device('cuda', torch.cuda.current_device())

        def agree(err, what):
            flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=tdev)
            if on:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if int(flag.item()):
                self.close()
                raise RuntimeError("PeerExchange: %s failed on %s" % (what, "this rank: %s" % err if err is not None else "another rank"))

        err = None
        try:
            N.check(L.prsd_exchange_create(context.ptr, self.world, self.rank, self.n_per_rank, C.byref(self.ptr)))
        except Exception as e:                                             # noqa: BLE001
            err = e
        agree(err, "allocating the receive buffer")
        if self.world > 1:
            mine = np.zeros(128, dtype=np.uint8)
            N.check(L.prsd_exchange_handles(self.ptr, mine.ctypes.data))
            send = torch.from_numpy(mine).to(tdev)
            recv = torch.empty(self.world*128, dtype=torch.uint8, device=tdev)
            dist.all_gather_into_tensor(recv, send, group=group)
            allh = np.ascontiguousarray(recv.cpu().numpy())
            try:
                N.check(L.prsd_exchange_open(self.ptr, allh.ctypes.data))
            except Exception as e:                                         # noqa: BLE001
                err = e
            agree(err, "mapping the peers' buffers (CUDA IPC)")
            dist.barrier(group=group)              # every rank has mapped every buffer before the first step

    def apply(self, gridop_ptr, nw, d_power_ptr):
        """enqueue the gridded transfer of ``nw`` walkers + the exchange on the context's stream"""
        self._N.check(self._L.prsd_gridop_apply_exchange(gridop_ptr, int(nw), d_power_ptr, self.ptr))

    def received(self, check=False):
        """torch tensor [world][n_per_rank] over the receive buffer of the latest step (valid behind the step on the
        context's stream); ``check``: synchronise and raise if a wait for a peer timed out"""
        C = self._C
        p, st = C.c_void_p(), C.c_int(0)
        self._N.check(self._L.prsd_exchange_received(self.ptr, C.byref(p), C.byref(st) if check else None))
        if check and st.value != 0:
            raise RuntimeError("PeerExchange: a rank did not deliver its step within 2 s")
        return torch.as_tensor(_DeviceArray(p.value, (self.world, self.n_per_rank), self), device=torch.device('cuda', self._ctx.device))

    def close(self):
        if getattr(self, 'ptr', None):
            self._L.prsd_exchange_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
