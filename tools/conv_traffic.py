#!/usr/bin/env python
"""Summarise an ncu CSV of the conv launches of ONE step into profiles/<tag>_conv_traffic.json.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,\\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_tc \\
        -s <convs per step> -c <convs per step> --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_step.py --batch 256
    python tools/conv_traffic.py gpurun_out/conv_traffic.csv profiles/r1_conv_traffic.json [batch]

`traffic_bytes` is what bench.py reports as roofline.traffic; `algorithmic_bytes` = every conv reads its input
(+ residual) once and writes its output once, logical channel counts."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"))
from acr_b200 import lib as L  # noqa: E402
from acr_b200.engine import Engine  # noqa: E402

src, dst = sys.argv[1], sys.argv[2]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "%": 1.0}
per = collections.OrderedDict()
for r in csv.DictReader(l for l in open(src) if not l.startswith("==")):
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    per.setdefault(r["ID"], {})[r["Metric Name"]] = v
rows = list(per.values())
convs = [r for r in Engine(None, 1, "cpu", dry_run=True).recs if r["kind"] == L.OP_CONV]
assert len(rows) == len(convs), (len(rows), len(convs))
alg = 0
for r in convs:
    x, y, at = r["ins"][0], r["out"], r["attrs"]
    alg += batch * x.H * x.W * x.C * 2 + batch * y.H * y.W * y.C * (4 if y.dtype == "f32" else 2) * (2 if at["residual"] else 1)
    alg += sum(batch * t.H * t.W * t.C * 2 for t in r["ins"][1:]) if at.get("extra") else 0     # folded fuse terms, read once
rd = sum(r["dram__bytes_read.sum"] for r in rows)
wr = sum(r["dram__bytes_write.sum"] for r in rows)
t = sum(r["gpu__time_duration.sum"] for r in rows)
tp = sum(r["gpu__time_duration.sum"] * r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for r in rows) / t
sys.path.insert(0, ROOT)
import bench as _bench  # noqa: E402  (the stamp bench.py checks before it reports roofline.traffic)
out = {"conv_build_id": os.environ.get("CONV_BUILD_ID") or _bench.conv_build_id(), "batch": batch, "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc, batch {batch}, one step ({len(rows)} launches)",
       "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr, "algorithmic_bytes": alg, "launches": len(rows),
       "time_ms_under_ncu": t / 1e6, "tensor_pipe_active_pct_time_weighted": tp}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
