#!/usr/bin/env python
"""MANO kernel alone: achieved HBM GB/s (algorithmic 9 820 B/hand, 19 324 B/hand with the projection
outputs) and FP32 rate at several hand counts."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
from acr_b200 import ops  # noqa: E402
from acr_b200.synth import make_synthetic_mano  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
ml = ops.pack_mano_model(make_synthetic_mano("left"), True, "cuda")
mr = ops.pack_mano_model(make_synthetic_mano("right"), False, "cuda")
for n in tuple(int(v) for v in os.environ.get('MANO_BENCH_N', '512,65536').split(',')):
    g = torch.Generator().manual_seed(0)
    poses = (torch.randn(n, 48, generator=g) * 0.5).cuda()
    betas = torch.randn(n, 10, generator=g).cuda()
    cam = (torch.rand(n, 3, generator=g) + 0.5).cuda()
    offs = torch.tensor([[512., 512, 0, 0, 0, 0, 0, 0, 0, 0]]).repeat(n, 1).cuda()
    ht = (torch.arange(n) >= n // 2).int().cuda()
    for _ in range(3):
        out = ops.mano_forward(ml, mr, poses, betas, ht, 1, 9, cam, offs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 20
    torch.cuda.synchronize()
    e0.record()
    for _ in range(it):
        out = ops.mano_forward(ml, mr, poses, betas, ht, 1, 9, cam, offs)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3   # includes the output allocations of the wrapper
    gbs = n * 19324 / us / 1e3
    print(json.dumps({"hands": n, "us": round(us, 2), "ns_per_hand": round(us * 1e3 / n, 2), "achieved_GBs": round(gbs, 1),
                      "hbm_peak_GBs": peaks["hbm_gbs"], "frac": round(gbs / peaks["hbm_gbs"], 4),
                      "fp32_TFLOPs": round(n * 1.152e6 / us / 1e6, 2)}))
