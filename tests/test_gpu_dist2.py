"""2 GPUs (run with `gpurun --gpus 2`; log kept in profiles/): the vertex all-gather fused into the MANO kernel
(symmetric-memory 16-byte peer / multimem stores over NVLink, double-buffered slots, arrival flags, counts carried
along -- no barrier, no NCCL call) must equal an NCCL all-gather of the same vertices, step after step."""
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
        pg = PeerVertexGather(R, dev, use_multicast=use_mc)
        nfl = world * R * 778 * 3 * 4
        for s in range(2):                                   # sentinel in the vertex area of both slots (flags stay 0)
            pg.buf[s * pg.slot_bytes: s * pg.slot_bytes + nfl].view(torch.float32).fill_(-7.0)
        torch.cuda.synchronize()
        dist.barrier()
        ht = (torch.arange(R) % 2).int().to(dev)
        ok, why = True, ""

        def launch(step):
            g = torch.Generator().manual_seed(1000 * step + rank)
            poses = (torch.randn(R, 48, generator=g) * 0.5).to(dev)
            betas = torch.randn(R, 10, generator=g).to(dev)
            n_valid = 40 + 7 * rank - 3 * step                # shrinking: rows beyond keep OLDER data, counts say so
            counts = torch.zeros(8, dtype=torch.int32, device=dev)
            counts[2], counts[0], counts[1] = n_valid, n_valid // 2, n_valid - n_valid // 2
            if (step + rank) % 2 == 0:
                torch.cuda._sleep(int(3e7))                   # skew the ranks (~15 ms) to exercise the drift
            out = ops.mano_forward(ml, mr, poses, betas, ht, 1, 9, n_dev=counts[2:3], peers=pg, counts=counts)
            return out, counts

        # (a) every step consumed: finish(), compare every shard with an NCCL all-gather of the same step
        for step in range(1, 7):
            out, counts = launch(step)
            pg.finish()
            ref, cnt = gather_vertices(out["verts"], counts)
            torch.cuda.synchronize()
            got, gcnt = pg.gathered(), pg.counts()
            if not torch.equal(gcnt, cnt):
                ok, why = False, f"step {step}: counts {gcnt.tolist()} != {cnt.tolist()}"
            for r in range(world):
                nv = int(cnt[r, 2])
                if not torch.equal(got[r, :nv], ref[r, :nv]):
                    ok, why = False, f"step {step}: shard {r} differs"
            if step == 1:
                for r in range(world):
                    if not bool((got[r, int(cnt[r, 2]):] == -7.0).all()):
                        ok, why = False, "rows >= n_dev were written"
        # (b) four launches back to back without any wait in between (ranks drift), then one finish()
        last = None
        for step in range(7, 11):
            last = launch(step)
        pg.finish()
        ref, cnt = gather_vertices(last[0]["verts"], last[1])
        torch.cuda.synchronize()
        for r in range(world):
            nv = int(cnt[r, 2])
            if not (torch.equal(pg.gathered()[r, :nv], ref[r, :nv]) and torch.equal(pg.counts(), cnt)):
                ok, why = False, f"back-to-back: shard {r} differs"
        q.put((rank, bool(ok), why or pg.mode))
        dist.barrier()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_mc", [False, True])
def test_fused_vertex_all_gather_2gpu(use_mc):
    """Ten steps on 2 GPUs with skewed ranks: double-buffered slots, flags instead of barriers, counts carried in
    the symmetric buffer; every shard of every consumed step must equal the NCCL all-gather bit for bit."""
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
