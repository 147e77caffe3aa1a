"""CPU: the committed ncu evidence under profiles/ and the tools that summarise it stay consistent
(bench.py reads roofline.traffic from profiles/r1_conv_traffic.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
R1_PLAN = {"ACR_B200_MERGE_STEMS": "0"}      # the round-1 captures are of the round-1 plan (eight separate head stems)


def _run(*args, env=None):
    r = subprocess.run([sys.executable] + list(args), cwd=ROOT, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_conv_traffic_json_is_what_the_tool_derives_from_the_launch_csv(tmp_path):
    dst = str(tmp_path / "t.json")
    _run("tools/conv_traffic.py", "profiles/r1_final_conv_launches.csv", dst, env=R1_PLAN)
    got, ref = json.load(open(dst)), json.load(open(os.path.join(ROOT, "profiles", "r1_conv_traffic.json")))
    assert {k: got[k] for k in ref} == ref          # (the tool now also stamps the kernel build and the batch)
    assert ref["launches"] == 347
    assert 0.8 < ref["traffic_bytes"] / ref["algorithmic_bytes"] < 1.1       # no re-read waste


def test_layer_tables_cover_one_whole_step():
    convs = _run("tools/layer_table.py", "convs", "profiles/r1_final_conv_launches.csv", env=R1_PLAN)
    assert "over 347 launches" in convs
    kernels = _run("tools/layer_table.py", "kernels", "profiles/r1_final_launches.csv")
    for k in ("conv_tc_kernel<64, __nv_bfloat16, 7>", "pool_kernel", "im2col_stem_kernel", "mano_forward_kernel",
              "parse_top1_kernel", "fuse_kernel"):
        assert k in kernels, k


def test_committed_bench_lines_carry_the_contract_keys():
    for name in ("r1_final_bench_default.json", "r1_final_bench_2gpu_fused.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        assert d["e2e"]["h2d_bytes_per_step"] == 256 * 512 * 512 * 3 and d["gpu_launches"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_final_bench_default.json")))
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    r = json.load(open(os.path.join(ROOT, "profiles", "r1_final_bench_reference.json")))
    assert r["impl"] == "reference" and r["e2e"]["h2d_bytes_per_step"] == 0
