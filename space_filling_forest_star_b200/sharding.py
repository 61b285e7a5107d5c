"""Multi-GPU sharding of pose / edge / query batches: one process per GPU, torch.distributed for the plumbing.

The obstacle BVH, robot mesh and node set are replicated on every GPU (a few MB); a batch is split into
``world_size`` contiguous, equally padded index ranges; each rank computes its slice with its own engine handle and
the slices are exchanged with ONE all-gather (NCCL over NVLink on GPUs; gloo in the CPU unit tests).  There is no
other data-path collective: every pose verdict and every query row is independent.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """-> (begin, end, padded_len): contiguous equal slices, the last ones possibly short or empty."""
    per = (n + world - 1) // world if n > 0 else 0
    b = min(rank * per, n)
    e = min(b + per, n)
    return b, e, per


def sharded_rows(n: int, row_shape: Tuple[int, ...], dtype: torch.dtype, device, compute: Callable[[int, int, torch.Tensor], None],
                 group=None, pad_value=0) -> torch.Tensor:
    """Generic "split, compute local slice, all-gather" driver.

    ``compute(begin, end, out)`` fills ``out[: end - begin]`` (a view of this rank's padded slice).  Returns the full
    ``[n, *row_shape]`` tensor on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e, per = shard_bounds(n, rank, world)
    local = torch.full((per, *row_shape), pad_value, dtype=dtype, device=device)
    if e > b:
        compute(b, e, local)
    if world == 1:
        return local[:n]
    full = torch.empty((world * per, *row_shape), dtype=dtype, device=device)
    dist.all_gather_into_tensor(full, local, group=group)
    return full[:n]


def sharded_collide(env, poses: torch.Tensor, group=None) -> torch.Tensor:
    """poses: CUDA tensor [n][6] replicated (or at least valid for this rank's slice) -> uint8 [n] on every rank."""
    n = poses.shape[0]

    def compute(b, e, out):
        env.collide_device(poses[b:e].contiguous(), out=out[: e - b])

    return sharded_rows(n, (), torch.uint8, poses.device, compute, group)


def sharded_edges(env, starts: torch.Tensor, ends: torch.Tensor, sample_dist: float, rot_mode: int, group=None) -> torch.Tensor:
    m = starts.shape[0]

    def compute(b, e, out):
        env.edges_device(starts[b:e].contiguous(), ends[b:e].contiguous(), sample_dist, rot_mode, free_out=out[: e - b])

    return sharded_rows(m, (), torch.uint8, starts.device, compute, group)


def sharded_knn(index, queries: torch.Tensor, k: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Node set replicated, query rows split.  (ids, d2) rows are packed as one int32 pair tensor for a single gather."""
    nq = queries.shape[0]

    def compute(b, e, out):
        ids = out[: e - b, :, 0]
        d2 = out[: e - b, :, 1].view(torch.float32)
        i_tmp, d_tmp = index.knn_device(queries[b:e].contiguous(), k)
        ids.copy_(i_tmp)
        d2.copy_(d_tmp)

    packed = sharded_rows(nq, (k, 2), torch.int32, queries.device, compute, group)
    return packed[:, :, 0].contiguous(), packed[:, :, 1].contiguous().view(torch.float32)
