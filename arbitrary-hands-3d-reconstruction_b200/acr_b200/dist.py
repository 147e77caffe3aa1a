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


class PeerVertexGather:
    """Vertex all-gather fused into the MANO kernel: a symmetric-memory buffer (world, R, 778, 3) whose peer
    addresses (and, with NVLS, its multicast address) are handed to ``acr_b200_mano_forward_gather``; the
    kernel's epilogue stores every vertex into all ranks' buffers over NVLink (``multimem.st`` through the
    NVSwitch when multicast is available, per-peer stores otherwise).  ``finish()`` is the cross-rank barrier
    that makes the stores of all ranks visible (symmetric-memory signal pads, enqueued on the current stream)."""

    def __init__(self, rows: int, device, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.rows = int(rows)
        self.buf = symm_mem.empty((self.world, self.rows, 778, 3), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        mc = 0
        if use_multicast:
            try:
                if self.hdl.has_multicast_support:
                    mc = int(self.hdl.multicast_ptr or 0)
            except Exception:
                mc = 0
        self.multicast_ptr = mc
        self.dst_row_offset = self.rank * self.rows

    @property
    def mode(self) -> str:
        return "multimem.st (NVLS multicast)" if self.multicast_ptr else "peer stores"

    def finish(self) -> None:
        self.hdl.barrier(channel=0)

    def gathered(self) -> torch.Tensor:
        return self.buf
