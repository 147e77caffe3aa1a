"""CPU: the committed ncu evidence under profiles/ and the tools that summarise it stay consistent
(bench.py reads roofline.traffic from the newest profiles/r*_conv_traffic.json whose kernel-build stamp matches)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
R1_PLAN = {"ACR_B200_MERGE_STEMS": "0", "ACR_B200_FOLD_FUSE": "0", "ACR_B200_STEM_FUSED": "0"}   # the round-1 captures are of the round-1 plan


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


def test_round2_conv_traffic_and_tables_match_the_current_plan(tmp_path):
    """The round-2 captures are of the current plan (340 conv launches: merged head stems, conv1 in its own fused kernel)."""
    dst = str(tmp_path / "t.json")
    _run("tools/conv_traffic.py", "profiles/r2_final_conv_launches.csv", dst, env={"CONV_BUILD_ID": "x"})     # the current plan
    got, ref = json.load(open(dst)), json.load(open(os.path.join(ROOT, "profiles", "r2_conv_traffic.json")))
    for k in ("traffic_bytes", "algorithmic_bytes", "launches", "dram_read_bytes", "dram_write_bytes", "batch"):
        assert got[k] == ref[k], k
    assert ref["launches"] == 340 and len(ref["conv_build_id"]) == 12
    assert 0.8 < ref["traffic_bytes"] / ref["algorithmic_bytes"] < 1.1
    convs = _run("tools/layer_table.py", "convs", "profiles/r2_final_conv_launches.csv")
    assert "over 340 launches" in convs
    kernels = _run("tools/layer_table.py", "kernels", "profiles/r2_final_launches.csv")
    for k in ("conv_tc_kernel<64, __nv_bfloat16, 23>", "conv_tc_kernel<64, __nv_bfloat16, 19>", "conv_tc_kernel<64, __nv_bfloat16, 34>",
              "stem_tc_kernel<__nv_bfloat16>", "pool_tc_kernel<__nv_bfloat16>", "mano_forward_kernel", "fuse_kernel"):
        assert k in kernels, k
    assert "im2col" not in kernels


def test_round2_bench_lines_carry_the_contract_keys():
    keys = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "roofline_mano", "mano_verts_max_abs_err")
    for name, n in (("r2_final_bench_default.json", 1), ("r2_bench_2gpu_fused.json", 2), ("r2_bench_8gpu_fused.json", 8),
                    ("r2_bench_8gpu_nccl.json", 8), ("r2_bench_fp16.json", 1), ("r2_bench_hrnet_w48.json", 1)):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read())
        for k in keys:
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["e2e"]["h2d_bytes_per_step"] == 256 * 512 * 512 * 3 and d["gpu_launches"] > 0
        assert d["e2e"]["d2h_bytes_per_step"] > 512 * 778 * 3 * 4 + 512 * 21 * 3 * 4          # verts AND joints AND parameters
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert {"bound", "achieved", "peak", "unit", "frac", "at_65536_hands"} <= set(d["roofline_mano"])
        assert d["mano_verts_max_abs_err"] < 1e-6
        if n > 1:
            assert d["gather_check"] is True and len(d["ms_per_step_by_rank"]) == n
    d = json.loads(open(os.path.join(ROOT, "profiles", "r2_final_bench_default.json")).read())
    # the roofline's flops are those of the 340 conv_tc launches only: conv1 runs in stem_tc_kernel (timed as kind 11), so
    # its 2 * 256 * 256 * 64 * 27 flops per image must not be credited to the conv kernels' time
    stem_gflop = 2.0 * 256 * 256 * 64 * 27 / 1e9 * 256
    assert abs(d["roofline"]["algorithmic_gflop_per_launch_set"] - (25819.4979 - stem_gflop)) < 0.5
    assert "340 launches" in d["roofline"]["kernel"] and "11" in d["profile_ms_by_kind"]
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "reference" and {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    r = json.loads(open(os.path.join(ROOT, "profiles", "r2_final_bench_reference.json")).read())
    assert r["impl"] == "reference" and r["cpu_baseline"]["kind"] == "reference" and r["e2e"]["h2d_bytes_per_step"] == 0


def test_product_kernels_are_tcgen05_in_sass():
    """cuobjdump of the in-tree library (cross-compiled here by build()): every instance of the conv, stem and pooling GEMM
    kernels issues tcgen05.mma (UTCHMMA) with TMEM loads and mbarriers, the operand / result movers are TMA, and warp-level
    HMMA exists only in the mma.sync pooling kernel kept as an opt-out."""
    lib = os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200", "lib", "libacr_b200.so")
    if not (os.path.exists(lib) and os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("library not built or no cuobjdump")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import sass_audit
        rows = sass_audit.audit(lib)
    finally:
        sys.path.pop(0)
    gemm = {n: r for n, r in rows.items() if n.startswith(("conv_tc_kernel<", "stem_tc_kernel<", "pool_tc_kernel<"))}
    assert len(gemm) >= 40
    for n, r in gemm.items():
        assert r["UTCHMMA"] > 0 and r["LDTM"] > 0 and r["UTCBAR"] > 0 and r["SYNCS"] > 0 and r["HMMA"] == 0, n
        assert r["UTMALDG"] > 0 or n.startswith("stem_tc_kernel<"), n          # the stem builds its operand from uint8 itself
        assert r["UTMASTG"] > 0 or n.startswith("pool_tc_kernel<"), n          # the pooling result is 32 KB of fp32 per CTA
    assert {n for n, r in rows.items() if r["HMMA"]} == {"pool_kernel<__half>", "pool_kernel<__nv_bfloat16>"}
    assert any(r["FFMA2"] for n, r in rows.items() if n.startswith("mano_forward_kernel"))
