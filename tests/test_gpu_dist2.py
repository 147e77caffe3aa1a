"""2 GPUs (run with `gpurun --gpus 2`): the vertex all-gather fused into the MANO kernel (symmetric-memory
peer / multimem stores over NVLink) must equal an NCCL all-gather of the same vertices."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, use_mc):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ACR_B200_SYNTHETIC_MANO="1")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from acr_b200 import ops
        from acr_b200.dist import PeerVertexGather, gather_vertices
        from acr_b200.synth import make_synthetic_mano
        ml = ops.pack_mano_model(make_synthetic_mano("left"), True, dev)
        mr = ops.pack_mano_model(make_synthetic_mano("right"), False, dev)
        R = 64
        g = torch.Generator().manual_seed(10 + rank)
        poses = (torch.randn(R, 48, generator=g) * 0.5).to(dev)
        betas = torch.randn(R, 10, generator=g).to(dev)
        ht = (torch.arange(R) % 2).int().to(dev)
        n_valid = 40 + 7 * rank
        n_dev = torch.tensor([n_valid], dtype=torch.int32, device=dev)
        pg = PeerVertexGather(R, dev, use_multicast=use_mc)
        pg.buf.fill_(-7.0)
        torch.cuda.synchronize()
        dist.barrier()
        out = ops.mano_forward(ml, mr, poses, betas, ht, 1, 9, n_dev=n_dev, peers=pg)
        pg.finish()
        torch.cuda.synchronize()
        counts = torch.zeros(8, dtype=torch.int32, device=dev)
        counts[2] = n_valid
        ref, cnt = gather_vertices(out["verts"], counts)
        torch.cuda.synchronize()
        ok = True
        for r in range(world):
            nv = int(cnt[r, 2])
            ok = ok and torch.equal(pg.gathered()[r, :nv], ref[r, :nv])
            ok = ok and bool((pg.gathered()[r, nv:] == -7.0).all())      # rows >= n_dev are never written
        q.put((rank, bool(ok), pg.mode))
        dist.barrier()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_mc", [False, True])
def test_fused_vertex_all_gather_2gpu(use_mc):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, use_mc)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    print("fused gather:", res)
    assert all(r[1] for r in res), res
