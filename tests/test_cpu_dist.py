"""CPU, world_size 2 over gloo: the sharding + vertex all-gather host logic of the N>1 path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acr_b200.dist import compact_gathered, gather_vertices, shard_range


def test_shard_range_covers_everything():
    for total in (1, 7, 256, 1024, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R = 6                                   # worst-case rows per rank (2 * local batch)
    n_valid = 3 + rank                      # ragged number of detected hands per shard
    verts = torch.full((R, 778, 3), float(rank + 1))
    verts[n_valid:] = -1.0                  # rows beyond L+R are garbage by contract
    verts[:, 0, 0] = torch.arange(R) + 100 * rank
    counts = torch.zeros(8, dtype=torch.int32)
    counts[2] = n_valid
    g, c = gather_vertices(verts, counts)
    parts = compact_gathered(g, c)
    ok = g.shape == (world, R, 778, 3) and [p.shape[0] for p in parts] == [3, 4]
    for r, p in enumerate(parts):
        ok = ok and bool((p[:, 1:, :] == r + 1).all()) and p[:, 0, 0].tolist() == [100 * r + i for i in range(3 + r)]
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_vertex_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
