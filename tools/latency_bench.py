#!/usr/bin/env python
"""Small-batch latency of the whole hot path, eager launches vs one CUDA graph (serving / webcam mode)."""
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"), ROOT):
    sys.path.insert(0, p)
os.environ.setdefault("ACR_B200_SYNTHETIC_MANO", "1")
import torch  # noqa: E402
from acr.config import args as cfg  # noqa: E402
from acr.main import ACR  # noqa: E402
from acr_b200.synth import load_bn_calibration, make_synthetic_mano, synth_state_dict  # noqa: E402

cfg().return_maps = False
app = ACR(state_dict=synth_state_dict(0, bn_stats=load_bn_calibration(0)),
          mano_assets={"left": make_synthetic_mano("left"), "right": make_synthetic_mano("right")})
for B in (1, 4, 16):
    g = torch.Generator().manual_seed(B)
    frames = torch.randint(0, 256, (B, 512, 512, 3), generator=g, dtype=torch.uint8).cuda()
    offs = torch.tensor([[512., 512, 0, 0, 0, 0, 0, 0, 0, 0]]).repeat(B, 1).cuda()

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
            torch.cuda.synchronize()          # per-frame latency: the caller waits for each result
        return (time.perf_counter() - t0) / n * 1e3

    eager = timeit(lambda: app.fused_forward(frames, offs))
    replay = app.capture_graph(B)
    bufs_e, mano_e = app.fused_forward(frames, offs)
    torch.cuda.synchronize()
    ve = mano_e["verts"].clone()
    bufs_g, mano_g = replay(frames, offs)
    torch.cuda.synchronize()
    same = bool(torch.equal(ve[: int(bufs_e.counts[2])], mano_g["verts"][: int(bufs_g.counts[2])]))
    graph = timeit(lambda: replay(frames, offs))
    print(json.dumps({"batch": B, "eager_ms": round(eager, 3), "graph_ms": round(graph, 3),
                      "graph_fps": round(B / graph * 1e3, 1), "identical_results": same}), flush=True)
