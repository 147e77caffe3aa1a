"""Multi-GPU plumbing: one process per GPU (torchrun), frames sharded by contiguous ranges, no
data-path collective except ONE all-gather of the output vertices (NCCL over NVLink on GPUs, gloo
in the CPU tests).  The reference's only parallelism is a single-process nn.DataParallel wrapper
(/root/reference/acr/main.py:61) that degenerates to one GPU."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of frames owned by `rank`; remainders go to the low ranks."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_vertices(verts: torch.Tensor, counts: torch.Tensor, out: torch.Tensor = None,
                    counts_out: torch.Tensor = None):
    """All-gather the dense per-rank vertex buffers (R, 778, 3) (R = worst-case rows, identical on
    every rank) and the (8,) int32 count vectors.  Returns (world, R, 778, 3), (world, 8).
    Asynchronous w.r.t. the host on NCCL (enqueued on the current stream)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if out is None:
        out = torch.empty((world,) + tuple(verts.shape), dtype=verts.dtype, device=verts.device)
    if counts_out is None:
        counts_out = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
    if world == 1:
        out[0].copy_(verts)
        counts_out[0].copy_(counts)
        return out, counts_out
    dist.all_gather_into_tensor(out.view(-1), verts.contiguous().view(-1))
    dist.all_gather_into_tensor(counts_out.view(-1), counts.contiguous().view(-1))
    return out, counts_out


def compact_gathered(gathered: torch.Tensor, counts: torch.Tensor) -> List[torch.Tensor]:
    """Valid rows of every rank's shard, in rank order (counts[:, 2] = L+R of each shard)."""
    n = counts[:, 2].tolist()
    return [gathered[r, : int(n[r])] for r in range(gathered.shape[0])]
