"""CPU: the C-ABI library builds/loads and exports every symbol of include/acr_b200.h; host-side
packers and the plan builder's bookkeeping behave (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from acr_b200 import lib as L
from acr_b200.netspec import build_acr_spec, conv_flops_per_image
from acr_b200.synth import make_synthetic_mano, synth_state_dict
from tests.helpers import pack_conv_host, u16_to_float

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    from acr_b200.build import build
    build()
    return L.load()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "acr_b200.h")).read()
    declared = set(re.findall(r"\b(acr_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/acr_b200.h but not exported"
    assert set(L.EXPORTS) <= declared
    assert b"sm_100a" in lib.acr_b200_version()


def test_struct_sizes_match_c(lib):
    # sizeof(acr_b200_tensor)=32, sizeof(acr_b200_op)=456 with the header's field order
    assert ctypes.sizeof(L.Tensor) == 32
    assert ctypes.sizeof(L.Op) == 4 * 2 + 32 * 9 + 8 * 12 + 4 * 6 + 4 * 4 + 4 * 2 + 4 * 4


def test_error_reporting(lib):
    rc = lib.acr_b200_mano_forward(None, None, None, None, None, 1, None, 4, 9, None, None, None, None, None,
                                   None, None, None, None)
    assert rc == -1 and b"null" in lib.acr_b200_last_error()
    rc = lib.acr_b200_mano_forward(None, None, None, None, None, 1, None, 0, 9, None, None, None, None, None,
                                   None, None, None, None)
    assert rc == 0   # empty batch is a no-op (reference: "if empty, return empty", mano_wrapper.py:43)


def test_pack_conv_folds_bn(lib):
    g = np.random.default_rng(0)
    w = g.standard_normal((5, 7, 3, 3)).astype(np.float32)
    cb = g.standard_normal(5).astype(np.float32)
    bn = [g.random(5).astype(np.float32) + 0.5, g.standard_normal(5).astype(np.float32),
          g.standard_normal(5).astype(np.float32), g.random(5).astype(np.float32) + 0.5]
    wp, bias = pack_conv_host(w, cb, bn, 16, 16, L.DT_BF16)
    sc = bn[0] / np.sqrt(bn[3] + 1e-5)
    ref_w = (w * sc[:, None, None, None]).transpose(0, 2, 3, 1).reshape(5, 9, 7)    # [co][tap][ci]
    got = u16_to_float(wp, L.DT_BF16).numpy().reshape(16, 9, 16)
    assert np.abs(got[:5, :, :7] - ref_w).max() <= np.abs(ref_w).max() * 2 ** -8
    assert (got[5:] == 0).all() and (got[:, :, 7:] == 0).all()
    assert np.allclose(bias[:5], bn[1] - bn[2] * sc + cb * sc, atol=1e-6) and (bias[5:] == 0).all()


def test_mano_pack_model_host(lib):
    a = make_synthetic_mano("left")
    n = lib.acr_b200_mano_model_floats()
    out = np.zeros(n, np.float32)
    arrs = [np.ascontiguousarray(a[k], np.float32) for k in
            ("shapedirs", "posedirs", "v_template", "J_regressor", "weights", "hands_mean")]
    assert lib.acr_b200_mano_pack_model(*[x.ctypes.data for x in arrs], 1, out.ctypes.data) == 0
    NVP = 784
    dirs = out[:145 * 3 * NVP].reshape(145, 3, NVP)
    assert np.array_equal(dirs[:135, :, :778], a["posedirs"].transpose(2, 1, 0))
    sd = a["shapedirs"].copy()
    sd[:, 0, :] *= -1
    assert np.array_equal(dirs[135:, :, :778], sd.transpose(2, 1, 0))
    jt = out[145 * 3 * NVP + 3 * NVP + 16 * NVP:][:48].reshape(16, 3)
    assert np.allclose(jt, a["J_regressor"] @ a["v_template"], atol=1e-7)


def test_spec_matches_reference_counts():
    spec = build_acr_spec()
    assert len(spec.params) == 2067                       # 2067 tensors
    n = sum(int(np.prod(s)) for s, _ in spec.params.values())
    assert abs(n - 30.31e6) < 0.2e6
    assert abs(conv_flops_per_image(spec) / 1e9 - 102.12) < 0.05
    sd = synth_state_dict(3)
    assert list(sd.keys()) == list(spec.params.keys())


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from acr_b200 import ops
    with pytest.raises(L.AcrB200Error):
        ops.rot6d_to_aa(torch.zeros(2, 6))
    from acr_b200.engine import Engine
    with pytest.raises(L.AcrB200Error):
        Engine({}, 1, "cpu")


def test_xpair_weight_expansion_has_the_corners_the_kernel_multiplies(lib):
    """ACR_CONV_XPAIR: the conv kernel only multiplies the [N 0..31][K 32..63] corner of the kx=0 taps and the
    [N 32..63][K 0..31] corner of the kx=2 taps of an x-paired 32->32 conv.  Everything it skips must be zero in
    the packed weights, and the centre taps must be what they were (engine._pack_conv(pair=True))."""
    from acr_b200.engine import Engine, _Blob
    rng = np.random.default_rng(5)
    sd = {"c.weight": rng.standard_normal((32, 32, 3, 3)).astype(np.float32)}
    eng = Engine(None, 1, "cpu", dry_run=True)
    blob = _Blob()
    w_off, _ = eng._pack_conv(sd, blob, "c", None, False, 64, 64, pair=True)
    wp = np.frombuffer(blob.tobytes(), np.uint16, count=64 * 9 * 64, offset=w_off).reshape(64, 3, 3, 64)  # [n][ky][pt][k]
    left, centre, right = wp[:, :, 0, :], wp[:, :, 1, :], wp[:, :, 2, :]
    assert not left[32:].any() and not left[:, :, :32].any() and left[:32, :, 32:].all()
    assert not right[:32].any() and not right[:, :, 32:].any() and right[32:, :, :32].all()
    assert centre.all()
    # even output pixel <- odd pixel of the left pair through the original kx = 0 tap (bf16-rounded)
    ref = torch.from_numpy(sd["c.weight"][:, :, :, 0]).bfloat16().view(torch.int16).numpy().view(np.uint16)   # [co][ci][ky]
    assert (left[:32, :, 32:] == ref.transpose(0, 2, 1)).all()


def test_plan_records_of_the_engine():
    """Launch-plan shape of the whole network (dry run, no GPU): op kinds, tensor-core stem, x-paired convs."""
    from acr_b200.engine import Engine
    eng = Engine(None, 2, "cpu", dry_run=True)
    kinds = [r["kind"] for r in eng.recs]
    # 348 Conv2d of the reference: conv1 runs as its own fused tcgen05 kernel (csrc/stem_tc.cu), the two
    # contact_layers[4|5] convs are folded 1x1 launches (counted), the eight head stems run as two merged convs (-6)
    assert kinds.count(L.OP_CONV) == 340 and kinds.count(L.OP_STEM_TC) == 1 and kinds.count(L.OP_IM2COL_STEM) == 0
    assert kinds.count(L.OP_STEM) == 0 and kinds[0] == L.OP_STEM_TC
    assert kinds.count(L.OP_FUSE) == 23 and kinds.count(L.OP_POOL) == 1 and kinds.count(L.OP_PARTHEAD) == 1
    assert eng.n_ops == len(kinds) == 368
    # ACR_B200_STEM_FUSED=0: round 1's im2col (27 normalised taps -> 32 channels) + a 1x1 tcgen05 conv
    os.environ["ACR_B200_STEM_FUSED"] = "0"
    try:
        k0 = [r["kind"] for r in Engine(None, 2, "cpu", dry_run=True).recs]
    finally:
        del os.environ["ACR_B200_STEM_FUSED"]
    assert k0.count(L.OP_CONV) == 341 and k0.count(L.OP_IM2COL_STEM) == 1 and len(k0) == 369
    # frame sizes whose tile counts are not powers of two keep the im2col stem (stem_tc.cu indexes tiles with shifts)
    k384 = [r["kind"] for r in Engine(None, 1, "cpu", dry_run=True, input_size=384).recs]
    assert k384.count(L.OP_STEM_TC) == 0 and k384.count(L.OP_IM2COL_STEM) == 1
    # opt-in (ACR_B200_FOLD_FUSE=1): the 15 fuse sums of the coarser outputs run inside the stride-2 conv of their finer neighbour
    from acr_b200.netspec import build_acr_spec
    spec = build_acr_spec(512, fold_fuse=True)
    folded = [o for o in spec.ops if o.kind == "conv" and o.attrs.get("extra")]
    assert sum(o.kind == "fuse" for o in spec.ops) == 8 and len(folded) == 15
    assert all(o.attrs["s"] == 2 and o.attrs["relu"] and 2 <= len(o.ins) <= 4 for o in folded)
    assert list(spec.params.items()) == list(eng.spec.params.items())          # same registry, same order
    assert eng.arena_bytes == 2 * 26 * 2 ** 20
    merged = [r for r in eng.recs if r["kind"] == L.OP_CONV and r["attrs"].get("merged")]
    assert len(merged) == 2 and all(r["out"].C == 256 and len(r["attrs"]["w"]) == 4 for r in merged)
    c32 = [r for r in eng.recs if r["kind"] == L.OP_CONV and r["attrs"]["k"] == 3 and r["attrs"]["s"] == 1
           and r["ins"][0].C == 32 and r["out"].C == 32 and r["ins"][0].H == 128]
    assert len(c32) == 64                                  # the x-paired class: 32 BasicBlocks of branch 0
    old = Engine(None, 2, "cpu", dry_run=True, stem_on_tensor_cores=False)
    assert [r["kind"] for r in old.recs].count(L.OP_STEM) == 1 and old.n_ops == 368


def test_s2x_weight_packing_places_each_tap_in_its_k_half(lib):
    """ACR_CONV_S2X: a 3x3 stride-2 conv of a dense 32-channel tensor reads x-pairs (128-byte row = even pixel's
    channels | odd neighbour's).  Tap (ky,kx) multiplies k-steps {0,1} (kx = 1: the even half) or {2,3} (kx = 0 / 2: the
    odd half of pair ox-1 / pair ox), so the packed weights must carry its 32 input channels at K offset 32*(kx != 1) and
    zeros in the other half (engine._pack_conv(s2x=True))."""
    from acr_b200.engine import Engine, _Blob
    rng = np.random.default_rng(6)
    sd = {"c.weight": rng.standard_normal((48, 32, 3, 3)).astype(np.float32)}
    eng = Engine(None, 1, "cpu", dry_run=True)
    blob = _Blob()
    w_off, _ = eng._pack_conv(sd, blob, "c", None, False, 64, 48, s2x=True)
    wp = np.frombuffer(blob.tobytes(), np.uint16, count=48 * 9 * 64, offset=w_off).reshape(48, 3, 3, 64)     # [n][ky][kx][k]
    ref = torch.from_numpy(sd["c.weight"]).bfloat16().view(torch.int16).numpy().view(np.uint16)              # [co][ci][ky][kx]
    for kx in range(3):
        used, unused = (slice(0, 32), slice(32, 64)) if kx == 1 else (slice(32, 64), slice(0, 32))
        assert not wp[:, :, kx, unused].any()
        assert (wp[:, :, kx, used] == ref[:, :, :, kx].transpose(0, 2, 1)).all()


def test_merged_head_stems_concatenate_weights_and_slice_outputs():
    """The eight head stems run as two N = 256 convs: slice j of the wide output must be conv j (weights / biases
    concatenated along cout in order), its consumers read channel slices [64 j, 64 j + 64) of the 256-wide tensor, and
    the registration order of the parameters (which the seeded weights and the goldens depend on) is the reference's."""
    from acr_b200.engine import Engine
    from acr_b200.netspec import build_acr_spec
    a, b = build_acr_spec(512, merge_stems=True), build_acr_spec(512, merge_stems=False)
    assert list(a.params.items()) == list(b.params.items())                  # same keys, shapes AND order
    eng = Engine(None, 1, "cpu", dry_run=True)
    merged = [r for r in eng.recs if r.get("attrs", {}).get("merged")]
    assert [r["attrs"]["w"] for r in merged] == [[f"{s}_final_layers.{i}.0.0" for i in (1, 2, 3, 4)] for s in "lr"]
    for r in merged:
        wide = r["out"]
        users = [q for q in eng.recs if any((t.base is wide) for t in q["ins"])]
        assert sorted({t.c_off for q in users for t in q["ins"] if t.base is wide}) == [0, 64, 128, 192]
        assert len(users) == 8                      # conv1 of the first BasicBlock and conv2's residual, per head
