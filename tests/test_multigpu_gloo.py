"""N>1 host logic on CPU: world_size-2 gloo process group.  Shards must tile the batch exactly, the constants
broadcast must deliver rank 0's bytes, and the counter reduction must sum / max correctly."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cvxpygen_b200.distributed import broadcast_constants, reduce_counters, shard_bounds, shard_params


def test_shard_bounds_tile_the_batch():
    for B in (0, 1, 7, 100000, 1000003):
        for W in (1, 2, 3, 8):
            b = [shard_bounds(B, W, r) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == B
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    blob = bytes([rank + 1]) * 4096                     # each rank starts with a different blob
    got = broadcast_constants(blob, src=0)
    params = np.arange(23 * 3, dtype=np.float64).reshape(23, 3)
    mine = shard_params(params, world, rank)
    red = reduce_counters({'n_solved': float(len(mine)), 'sum_iter': float(mine.sum()), 'max_ms': 10.0 + rank})
    q.put((rank, got == bytes([1]) * 4096, len(mine), red))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    assert [n for _, _, n, _ in res] == [12, 11]
    for _, _, _, red in res:
        assert red['n_solved'] == 23.0 and red['sum_iter'] == float(np.arange(69).sum()) and red['max_ms'] == 11.0
