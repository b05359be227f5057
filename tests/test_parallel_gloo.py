"""world_size-2 gloo test of the multi-GPU plumbing (batch sharding + scalar all-reduce), on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import avsep_b200 as av
from avsep_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(total, rank, world)
    # per-mixture "losses": a rank only sees its shard, the reduced mean must equal the global mean
    per_mix = torch.arange(total, dtype=torch.float32) * 0.25 + 1.0
    local = per_mix[lo:hi]
    g = parallel.global_mean_loss(local.mean(), hi - lo)
    sums = dict(sdr=float(local.sum()), sir=float((2 * local).sum()))
    m = parallel.global_metrics(sums, hi - lo)
    mx = parallel.max_over_ranks(float(rank + 1))
    # long-clip windows: rank r owns windows r, r + world, ...; every rank ends up with all of them, in order
    nw = 5
    mine = parallel.shard_windows(nw, rank, world)
    local = torch.stack([torch.full((2, 3), float(i)) for i in mine])
    allw = parallel.gather_windows(local, nw, rank, world)
    assert allw.shape == (nw, 2, 3) and [float(allw[i, 0, 0]) for i in range(nw)] == [0.0, 1.0, 2.0, 3.0, 4.0]
    q.put((rank, lo, hi, float(g), m, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for total in (0, 1, 7, 80, 128, 1024):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_windows(10, 1, 8) == [1, 9]


def test_two_rank_scalar_reduction():
    total, world = 7, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    per_mix = np.arange(total) * 0.25 + 1.0
    assert [o[1:3] for o in out] == [(0, 4), (4, 7)]
    for o in out:
        assert abs(o[3] - per_mix.mean()) < 1e-6
        assert abs(o[4]["sdr"] - per_mix.mean()) < 1e-9 and abs(o[4]["sir"] - 2 * per_mix.mean()) < 1e-9
        assert o[5] == 2.0


def test_single_process_is_identity():
    g = parallel.global_mean_loss(torch.tensor(0.5), 3)
    assert abs(float(g) - 0.5) < 1e-7
    assert parallel.global_metrics(dict(sdr=6.0), 3)["sdr"] == 2.0
    assert parallel.max_over_ranks(1.5) == 1.5
