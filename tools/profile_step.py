#!/usr/bin/env python
"""One warm-up step + one profiled step of the hot path, for ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... \\
        python tools/profile_step.py --batch 256 --range      (cudaProfilerStart/Stop around the LAST step)
"""
import argparse
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"), ROOT):
    sys.path.insert(0, p)
os.environ.setdefault("ACR_B200_SYNTHETIC_MANO", "1")
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--range", action="store_true", help="cudaProfilerStart/Stop around the last step")
a = ap.parse_args()
from acr.config import args as cfg  # noqa: E402
from acr.main import ACR  # noqa: E402
from acr_b200.synth import load_bn_calibration, make_synthetic_mano, synth_state_dict  # noqa: E402
cfg().model_precision = a.dtype
cfg().return_maps = False
app = ACR(state_dict=synth_state_dict(0, bn_stats=load_bn_calibration(0)),
          mano_assets={"left": make_synthetic_mano("left"), "right": make_synthetic_mano("right")})
g = torch.Generator().manual_seed(0)
frames = torch.randint(0, 256, (a.batch, 512, 512, 3), generator=g, dtype=torch.uint8).cuda()
offs = torch.tensor([[512., 512, 0, 0, 0, 0, 0, 0, 0, 0]]).repeat(a.batch, 1).cuda()
for i in range(a.steps):
    last = a.range and i == a.steps - 1
    if last:
        torch.cuda.cudart().cudaProfilerStart()
    app.fused_forward(frames, offs)
    torch.cuda.synchronize()
    if last:
        torch.cuda.cudart().cudaProfilerStop()
print("launches per step:", app.model.engine(a.batch, frames.device).num_launches + 4)
