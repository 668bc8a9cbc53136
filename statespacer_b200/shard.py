"""Multi-GPU plumbing of the batched path: one process per GPU, series sharded in contiguous blocks, no
collective inside the recursion (series, parameter vectors and draws are independent -- SURVEY.md 8e).
The only exchange is the gather of per-series results at the end of a sweep; it goes through
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of `total` items owned by `rank`; the first total % world ranks hold one
    more item."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def fd_gradient(loglik_v: torch.Tensor, h: float) -> torch.Tensor:
    """Central-difference gradient from the (1 + 2k, B) output of a batched sweep whose parameter vectors are
    ordered [theta, theta + h e_1, theta - h e_1, theta + h e_2, ...] (stats::optim's scheme,
    R/statespacer.R:977-982).  Returns (B, k)."""
    if loglik_v.shape[0] % 2 != 1:
        raise ValueError("expected 1 + 2k parameter vectors")
    return ((loglik_v[1::2] - loglik_v[2::2]) / (2.0 * h)).T.contiguous()


def gather_series(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather of per-series rows (B_local, c) into (total, c), ranks in shard order.  Uneven shards are
    padded to the largest one for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if local.shape[0] != total:
            raise ValueError("single process: local block must hold every series")
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    if local.shape[0] != sizes[rank][1] - sizes[rank][0]:
        raise ValueError("local block does not match shard_bounds")
    biggest = max(b - a for a, b in sizes)
    padded = local
    if local.shape[0] < biggest:
        padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    if all(b - a == biggest for a, b in sizes):
        return out
    return torch.cat([out[r * biggest: r * biggest + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)
