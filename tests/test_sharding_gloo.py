"""CPU tier: the N > 1 path (one process per GPU, no data-path collective) with world_size = 2 on gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spcies_b200 import sysmodel
from spcies_b200.sharding import reduce_scalar, shard_bounds


def test_shard_bounds_cover_batch_exactly():
    for B in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            sl = shard_bounds(B, world)
            assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            assert max(h - l for l, h in sl) - min(h - l for l, h in sl) <= (B + world - 1) // world


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys = sysmodel.oscillating_masses_sys()
    lo, hi = shard_bounds(B, world)[rank]
    batch = sysmodel.synthetic_batch(sys, B, seed=11)            # same seeded batch on every rank; each takes its slice
    x0 = batch['x0'][lo:hi]
    # stand-in for the per-rank solve: a per-instance function of the inputs only (instances are independent)
    result = x0.sum(axis=1)
    np.save(os.path.join(out_dir, f'r{rank}.npy'), result)
    t_max = reduce_scalar(0.1 * (rank + 1), 'max')
    n_sum = reduce_scalar(hi - lo, 'sum')
    if rank == 0:
        np.save(os.path.join(out_dir, 'meta.npy'), np.array([t_max, n_sum]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reassembles_the_batch(tmp_path):
    B, world = 1001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    sys = sysmodel.oscillating_masses_sys()
    full = sysmodel.synthetic_batch(sys, B, seed=11)['x0'].sum(axis=1)
    got = np.concatenate([np.load(tmp_path / f'r{r}.npy') for r in range(world)])
    assert np.array_equal(got, full)                              # slices are disjoint, ordered and cover the batch
    t_max, n_sum = np.load(tmp_path / 'meta.npy')
    assert t_max == 0.2 and n_sum == B                            # max-over-ranks time, summed instance count
