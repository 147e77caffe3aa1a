#!/usr/bin/env python
"""Time single conv ops through the C ABI (acr_b200_run_op via a 1-op plan) at full batch.
    python tools/conv_bench.py "32,32,3,1,128,1" "64,64,3,1,64,0" ...   (cin,cout,k,s,H,residual[,W[,flags]])
W defaults to H; flags = conv flag bits (4 = ACR_CONV_XPAIR).  ACR_B200_CONV_DIAG=1|2|4 times the diagnostic
kernel instances (issuer without MMAs / epilogue only recycling / epilogue without math and stores)."""
import ctypes as C
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from acr_b200 import lib as L  # noqa: E402
from tests.helpers import ctensor, rup  # noqa: E402

B = int(os.environ.get("BATCH", "256"))
lib = L.load()
for spec in sys.argv[1:]:
    fields = [int(v) for v in spec.split(",")]
    cin, cout, k, s, H, res = fields[:6]
    W = fields[6] if len(fields) > 6 else H
    flags = fields[7] if len(fields) > 7 else 0
    cin_pad, cout_pad = rup(cin, 16), rup(cout, 16)
    Ho, Wo = H // s, W // s
    s2x = bool(flags & 8)       # ACR_CONV_S2X: "32,cout,3,2,H,0,W,8" = dense 32-channel input read as x-pairs (H, W/2, 64)
    if s2x:
        assert cin == 32 and s == 2 and k == 3
        cin_pad = 64
    in_bytes = B * H * W * (32 if s2x else cin_pad) * 2
    out_bytes = B * Ho * Wo * cout_pad * 2
    off_r = rup(in_bytes, 1024)
    off_o = rup(off_r + out_bytes, 1024)
    arena = (torch.randn((off_o + out_bytes) // 2 + 512, device="cuda") * 0.5).to(torch.bfloat16)
    wts = (torch.randn(cout_pad * k * k * cin_pad + 1024, device="cuda") * 0.05).to(torch.bfloat16)
    bias_off = rup(cout_pad * k * k * cin_pad * 2, 256)
    blob = torch.zeros(bias_off + cout_pad * 4 + 256, dtype=torch.uint8, device="cuda")
    blob[: cout_pad * k * k * cin_pad * 2] = wts.view(torch.uint8)[: cout_pad * k * k * cin_pad * 2]
    op = L.Op()
    op.kind = L.OP_CONV
    op.n_in = 2 if res else 1
    op.in_[0] = ctensor(0, 64, H, W // 2, 64, L.DT_BF16) if s2x else ctensor(0, cin, H, W, cin_pad, L.DT_BF16)
    op.in_[1] = ctensor(off_r, cout, Ho, Wo, cout_pad, L.DT_BF16)
    op.out = ctensor(off_o, cout, Ho, Wo, cout_pad, L.DT_BF16)
    op.shift[0] = flags
    op.w_offset[0], op.w_offset[1] = 0, bias_off
    op.k, op.stride, op.relu, op.has_residual = k, s, 1, res
    op.cin_pad, op.cout_pad = cin_pad, cout_pad
    plan = C.c_void_p()
    ops = (L.Op * 1)(op)
    L.check(lib.acr_b200_plan_create(ops, 1, B, arena.data_ptr(), arena.numel() * 2, blob.data_ptr(), blob.numel(),
                                     L.DT_BF16, C.byref(plan)), "plan_create")
    st = torch.cuda.current_stream().cuda_stream
    dummy = torch.zeros(16, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        L.check(lib.acr_b200_plan_run(plan, dummy.data_ptr(), st), "run")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        L.check(lib.acr_b200_plan_run(plan, dummy.data_ptr(), st), "run")
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * Ho * Wo * cout * cin * k * k * B
    by = in_bytes + out_bytes * (2 if res else 1)
    print(f"{spec:>22s}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s (algorithmic)", flush=True)
    lib.acr_b200_plan_destroy(plan)
