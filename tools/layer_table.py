#!/usr/bin/env python
"""Per-layer-class conv timing table from an ncu launch list (gpu__time_duration.sum CSV)."""
import collections
import csv
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"))
from acr_b200 import lib as L  # noqa: E402
from acr_b200.engine import Engine  # noqa: E402

path, batch = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 256
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"]) for r in csv.DictReader(lines)
        if r.get("Metric Name") == "gpu__time_duration.sum"]
stems = [i for i, r in enumerate(rows) if "stem_kernel" in r[0]]
step = rows[stems[-1]:]
tot = sum(r[1] for r in step)
agg = collections.OrderedDict()
for name, ns, grid in step:
    a = agg.setdefault(name.split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += ns
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")
print(f"| **total** | {len(step)} | {tot / 1e6:.3f} | 100% |\n")
convs = [r for r in step if "conv_tc" in r[0]]
cops = [r for r in Engine(None, 1, "cpu", dry_run=True).recs if r["kind"] == L.OP_CONV]
assert len(cops) == len(convs), (len(cops), len(convs))
agg = collections.OrderedDict()
for r, (name, ns, grid) in zip(cops, convs):
    x, y, at = r["ins"][0], r["out"], r["attrs"]
    cin = 109 if "fold_side" in at else x.C
    key = (cin, y.C, at["k"], at["s"], x.H, bool(at["residual"]))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += ns
    a[2] += 2.0 * y.H * y.W * y.C * cin * at["k"] ** 2 * batch
print("| cin | cout | k | s | H_in | res | n | total ms | avg us | TFLOP/s |\n|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k[0]} | {k[1]} | {k[2]} | {k[3]} | {k[4]} | {'y' if k[5] else ''} | {v[0]} | {v[1] / 1e6:.2f} | {v[1] / v[0] / 1e3:.0f} | {v[2] / v[1] / 1e3:.0f} |")
