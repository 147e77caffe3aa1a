#!/usr/bin/env python
"""Tables from ncu CSVs of one step (tools/ncu_conv.sh):
    python tools/layer_table.py kernels <launch list csv>          per-kernel totals (gpu__time_duration.sum)
    python tools/layer_table.py convs <conv traffic csv> [batch]   per-layer-class conv table: time, TFLOP/s, measured
                                                                   DRAM GB/s and tensor-pipe activity"""
import collections
import csv
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"))
from acr_b200 import lib as L  # noqa: E402
from acr_b200.engine import Engine  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "%": 1.0}


def launches(path):
    per = collections.OrderedDict()
    for r in csv.DictReader(l for l in open(path) if not l.startswith("==")):
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    return list(per.values())


mode, path = sys.argv[1], sys.argv[2]
rows = launches(path)
if mode == "kernels":
    rows = [r for r in rows if "at::" not in r["name"]]   # drop torch helper kernels (fills, copies)
    tot = sum(r["gpu__time_duration.sum"] for r in rows)
    agg = collections.OrderedDict()
    for r in rows:
        name = r["name"].replace("acr::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
        a = agg.setdefault(name.split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += r["gpu__time_duration.sum"]
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {len(rows)} | {tot / 1e6:.3f} | 100% |")
else:
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    cops = [r for r in Engine(None, 1, "cpu", dry_run=True).recs if r["kind"] == L.OP_CONV]
    assert len(cops) == len(rows), (len(cops), len(rows))
    agg = collections.OrderedDict()
    for r, m in zip(cops, rows):
        x, y, at = r["ins"][0], r["out"], r["attrs"]
        cin = 109 if "fold_side" in at else (27 if "stem" in at else x.C)
        key = (cin, y.C, at["k"], at["s"], x.H, bool(at["residual"]))
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        ns = m["gpu__time_duration.sum"]
        a[0] += 1
        a[1] += ns
        a[2] += 2.0 * y.H * y.W * y.C * cin * at["k"] ** 2 * batch
        a[3] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a[4] += ns * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a[5] += batch * (x.H * x.W * x.C * 2 + y.H * y.W * y.C * (4 if y.dtype == "f32" else 2) * (2 if at["residual"] else 1))
        a[5] += sum(batch * t.H * t.W * t.C * 2 for t in r["ins"][1:]) if at.get("extra") else 0
    print("| cin | cout | k | s | H_in | res | n | total ms | avg us | TFLOP/s | DRAM GB/s (measured) | DRAM / algorithmic bytes | tensor pipe active |")
    print("|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[0]} | {k[1]} | {k[2]} | {k[3]} | {k[4]} | {'y' if k[5] else ''} | {v[0]} | {v[1] / 1e6:.2f} | {v[1] / v[0] / 1e3:.0f} | "
              f"{v[2] / v[1] / 1e3:.0f} | {v[3] / v[1]:.0f} | {v[3] / v[5]:.2f} | {v[4] / v[1]:.0f}% |")
    t = sum(v[1] for v in agg.values())
    print(f"\ntotal {t / 1e6:.2f} ms over {len(rows)} launches, {sum(v[2] for v in agg.values()) / t / 1e3:.0f} TFLOP/s, "
          f"DRAM {sum(v[3] for v in agg.values()) / 1e9:.1f} GB ({sum(v[3] for v in agg.values()) / t:.0f} GB/s)")
