"""N > 1 host logic on CPU: world_size 2 and 3 over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyrsd_b200 import parallel


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 2048, 2049):
        for world in (1, 2, 3, 8):
            blocks = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == parallel.shard_sizes(n, world)
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_bounds(n_total, rank, world)
        # the "result" of walker w is a row that depends only on w: what an ensemble step produces
        local = torch.stack([torch.arange(5, dtype=torch.float64) + 10.0*w for w in range(lo, hi)]) if hi > lo \
            else torch.zeros((0, 5), dtype=torch.float64)
        full = parallel.gather_rows(local, n_total)
        t = parallel.max_over_ranks(1.0 + rank)
        np.save(os.path.join(out_dir, 'r%d.npy' % rank), full.numpy())
        np.save(os.path.join(out_dir, 't%d.npy' % rank), np.array([t]))
        with pytest.raises(ValueError):
            parallel.gather_rows(torch.zeros((hi - lo + 1, 5), dtype=torch.float64), n_total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_total', [(2, 8), (2, 9), (3, 10)])
def test_gather_rows_over_gloo(tmp_path, world, n_total):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    expect = np.stack([np.arange(5.0) + 10.0*w for w in range(n_total)])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ('r%d.npy' % r)), expect)        # same, ordered result on every rank
        assert np.load(tmp_path / ('t%d.npy' % r))[0] == float(world)            # max over ranks


def test_single_process_passthrough():
    x = torch.arange(12, dtype=torch.float64).view(4, 3)
    assert parallel.gather_rows(x, 4) is x
    assert parallel.max_over_ranks(2.5) == 2.5
    with pytest.raises(ValueError):
        parallel.gather_rows(x, 5)


class _NoContext(object):
    ptr = None
    device = 0


def _exchange_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # without a device context the allocation fails on every rank; the agreement after it must turn that into an
        # exception on every rank -- nobody is left waiting in the handle all-gather
        try:
            parallel.PeerExchange(_NoContext(), 16)
            outcome = 'created'
        except RuntimeError as e:
            outcome = 'raised: ' + str(e)[:40]
        dist.barrier()
        with open(os.path.join(out_dir, 'x%d.txt' % rank), 'w') as f:
            f.write(outcome)
    finally:
        dist.destroy_process_group()


def test_peer_exchange_setup_fails_on_every_rank_together(tmp_path):
    """host logic of the NVLink peer exchange (parallel.PeerExchange): a set-up failure is agreed upon by all ranks"""
    port = _free_port()
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / ('x%d.txt' % r)).read().startswith('raised: PeerExchange: allocating')
