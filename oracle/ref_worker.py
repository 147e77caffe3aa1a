#!/usr/bin/env python
"""ORACLE (test infrastructure): times the reference's own CPU pipeline in a process of its own.

bench.py's B200 arm has OUR drop-in ``acr`` / ``mano`` packages imported; the reference's packages have the same
names, so the CPU baseline runs here, in a subprocess, and prints one JSON line:

    python oracle/ref_worker.py --batch 8 --steps 3 [--check]

Implementation timed: the snapshot oracle/_ref (oracle/make_ref.py) through oracle/ref_harness.py.  ``--check``
additionally verifies the snapshot against the committed golden (tests/golden/net_golden.npz) and exits non-zero
on a mismatch -- the pin that says the thing being timed IS the reference."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0, help="0 = sweep on the timed batch")
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "acr")):
        print(json.dumps({"unavailable": "no oracle/_ref snapshot (run oracle/make_ref.py in the build container)"}))
        return 0
    import numpy as np
    from oracle import ref_harness
    sys.path.insert(0, ref_harness.PKG)
    from acr_b200.synth import load_bn_calibration, synth_state_dict
    sys.path.remove(ref_harness.PKG)
    sd = synth_state_dict(0, bn_stats=load_bn_calibration(0))
    run = ref_harness.load_reference_pipeline(ref_root, sd)
    import torch
    cores = os.cpu_count()
    if a.check:
        g = np.load(os.path.join(ROOT, "tests", "golden", "net_golden.npz"))
        gi = torch.Generator().manual_seed(123)
        img = torch.randint(0, 256, (2, 512, 512, 3), generator=gi, dtype=torch.uint8)
        out = run(img, torch.tensor([[512., 512, 0, 0, 0, 0, 0, 0, 0, 0]] * 2))
        err = float(np.abs(out["verts"].numpy() - g["verts"]).max() / np.abs(g["verts"]).max())
        same = bool((out["l_centers_pred"].numpy() == g["l_centers_pred"]).all())
        print(json.dumps({"check": "snapshot vs tests/golden/net_golden.npz", "verts_rel_err": err, "same_centres": same}))
        return 0 if (err < 1e-4 and same) else 1
    gi = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (a.batch, 512, 512, 3), generator=gi, dtype=torch.uint8)
    offs = torch.tensor([[512., 512, 0, 0, 0, 0, 0, 0, 0, 0]]).repeat(a.batch, 1)
    step = lambda: run(img, offs)
    sweep = {}
    if a.threads:
        best = a.threads
        torch.set_num_threads(best)
        step()
    else:
        step()                                    # warm-up, not attributed
        best, best_t = None, None
        # more than 64 threads oversubscribe these small convs badly (measured: 128 threads = 82 s per 8-frame pass
        # against 2.1 s with 32), so the sweep stops at 64
        for n in sorted({min(cores, 64), 48, 32, 24, 16, 8}, reverse=True):
            if n > cores:
                continue
            torch.set_num_threads(n)
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            sweep[n] = round(dt, 3)
            if best_t is None or dt < best_t:
                best, best_t = n, dt
        torch.set_num_threads(best)
    for _ in range(a.warmup):
        step()
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"kind": "reference", "batch": a.batch, "steps": a.steps, "threads": best, "cores": cores, "sweep": sweep,
                      "s_per_step_mean": sum(ts) / len(ts), "s_per_step_min": min(ts),
                      "what": "the unmodified reference (oracle/_ref snapshot: acr.model.ACR.forward + ResultParser.parse + "
                              "ManoLayer x2 + projection), torch fp32 CPU"}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
