"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes shard a batch, form
finite-difference gradients locally and gather per-series results in shard order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from statespacer_b200 import shard


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 8, 1000001):
        for world in (1, 2, 3, 8):
            cuts = [shard.shard_bounds(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 2, 2)


def test_fd_gradient_matches_quadratic():
    k, B, h = 3, 5, 1e-3
    rng = np.random.default_rng(0)
    theta = rng.normal(size=(B, k))
    A = rng.normal(size=(k,))

    def f(t):
        return (t ** 2 * A).sum(axis=1)

    vals = [f(theta)]
    for i in range(k):
        for s in (1, -1):
            t = theta.copy()
            t[:, i] += s * h
            vals.append(f(t))
    g = shard.fd_gradient(torch.tensor(np.stack(vals)), h).numpy()
    assert np.allclose(g, 2 * theta * A, atol=1e-9)
    with pytest.raises(ValueError):
        shard.fd_gradient(torch.zeros(4, 3), h)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard.shard_bounds(total, rank, world)
    # every rank "evaluates" its own series: row i of the global table is [i, 10 i, -i]
    idx = torch.arange(a, b, dtype=torch.float64)
    local = torch.stack([idx, 10 * idx, -idx], dim=1)
    full = shard.gather_series(local, total)
    ok = full.shape == (total, 3) and bool(torch.equal(full[:, 0], torch.arange(total, dtype=torch.float64))) \
        and bool(torch.equal(full[:, 1], 10 * full[:, 0]))
    # sum of loglik over the whole job = all-reduce of local sums (the other reduction the path uses)
    s = local[:, 0].sum().reshape(1)
    dist.all_reduce(s)
    ok = ok and float(s) == total * (total - 1) / 2
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_world2_gloo(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
