"""Multi-GPU plumbing: one process per GPU (torchrun), observations row-sharded, one packed all-reduce per
EM step inside libgplvm_b200.so over its own NCCL communicator.  This module only computes shard ranges and
bootstraps the library's communicator through torch.distributed (rank 0's NCCL unique id is broadcast)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

from . import _lib


def shard_range(n_obs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of observations owned by `rank`: same rule as gplvm_session_create uses for the GPUs
    of one process (ceil(n / world) per shard, last shard short)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-n_obs // world)
    begin = min(per * rank, n_obs)
    return begin, max(0, min(per, n_obs - begin))


def packed_stats_length(D: int, M: int, missing: bool, fused_dense: bool = True) -> int:
    """Length (doubles) of the per-step all-reduced statistics buffer."""
    if missing:
        nv = M * (M + 1) // 2
        return D * (M + nv) + (D + 1) * M + nv
    mp = 16 if (fused_dense and D in (64, 128, 256) and M <= 16) else -(-M // 8) * 8
    return (D + mp) * mp


def init_library_from_torch(local_rank: int) -> Tuple[int, int]:
    """Create the library's communicator for this torch.distributed job.  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        _lib.check(lib.gplvm_dist_init(0, 1, local_rank, None))
        return 0, 1
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local_rank) if dist.get_backend() == "nccl" else torch.device("cpu")
    idbuf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        _lib.check(lib.gplvm_dist_unique_id(raw))
        idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
    _lib.check(lib.gplvm_dist_init(rank, world, local_rank, raw))
    return rank, world
