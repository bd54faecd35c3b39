"""world_size-2 gloo tests (CPU) of the N > 1 path's host logic: shard ranges, and that row-sharded sufficient
statistics combined by ONE all-reduce per EM step followed by a replicated update reproduce the unsharded
reference step -- the scheme libgplvm_b200 runs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gplvmhaskell_b200.dist import shard_range, packed_stats_length  # noqa: E402
from oracle import ppca as oppca  # noqa: E402


def test_shard_ranges_cover_everything():
    for n in (1, 7, 150, 1 << 20, (1 << 23) + 3):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(n, world, r) for r in range(world)]
            assert sum(c for _, c in got) == n
            pos = 0
            for b, c in got:
                assert b == pos or c == 0
                pos += c
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_packed_lengths_match_survey():
    assert packed_stats_length(256, 16, missing=False) == (256 + 16) * 16        # 34.8 KB
    assert packed_stats_length(256, 16, missing=True) == 256 * 152 + 257 * 16 + 136
    assert packed_stats_length(4, 3, missing=False) == (4 + 8) * 8


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    d, n, m = 12, 203, 3
    Y = rng.standard_normal((d, m)) @ rng.standard_normal((m, n)) + rng.standard_normal((d, 1)) + 0.3 * rng.standard_normal((d, n))
    mask = rng.random((d, n)) < 0.25
    mask[0] = False
    Yn = np.where(mask, np.nan, Y)
    W = rng.standard_normal((d, m))
    b, c = shard_range(n, world, rank)

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    # dense: one-time all-reduces (feature sums -> mean, |Xc|^2), then one packed all-reduce per step
    mean = allreduce(Y[:, b:b + c].sum(axis=1)) / n
    Xc = Y[:, b:b + c] - mean[:, None]
    total = float(allreduce(np.array([np.sum(Xc ** 2)]))[0])
    Wd, s2 = W.copy(), 1.3
    for _ in range(3):
        packed = allreduce(oppca.dense_stats(Xc, Wd, s2))
        Wd, s2, _, _ = oppca.dense_update_from_stats(packed, total, n, Wd, s2)
    # masked: constants once, then one packed all-reduce per step
    consts = allreduce(oppca.missed_constants(Yn[:, b:b + c]))
    mu, Wm, s2m = np.zeros(d), W.copy(), 1.3
    for _ in range(3):
        packed, _ = oppca.missed_stats(Yn[:, b:b + c], mu, Wm, s2m)
        packed = allreduce(packed)
        mu, Wm, s2m, _, _ = oppca.missed_update_from_stats(packed, consts, Wm, s2m)
    if rank == 0:
        q.put((Y, Yn, W, Wd, s2, mu, Wm, s2m))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_equal_unsharded_reference_step():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    Y, Yn, W, Wd, s2, mu, Wm, s2m = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = oppca.em_steps_fast(Y, W, 1.3, ("left", 2))          # literal reference order, unsharded
    assert np.max(np.abs(Wd - ref.W)) < 1e-11 * np.max(np.abs(ref.W))
    assert abs(s2 - ref.variance) < 1e-11 * ref.variance
    refm = oppca.em_steps_missed(Yn, W, 1.3, ("left", 2))      # literal per-observation closures, unsharded
    assert np.max(np.abs(Wm - refm.W)) < 1e-10 * np.max(np.abs(refm.W))
    assert abs(s2m - refm.variance) < 1e-11 * refm.variance
