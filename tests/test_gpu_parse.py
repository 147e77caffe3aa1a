"""GPU parity: centre parsing / sampling / 6D->aa kernels vs the reference goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN
from tests.test_oracle_golden import make_parse_case

pytestmark = pytest.mark.gpu

INT_KEYS = ("reorganize_idx", "l_centers_pred", "r_centers_pred", "left_hand_num", "right_hand_num", "output_hand_type")
F_KEYS = ("params_pred", "detection_flag", "l_centers_conf", "r_centers_conf")


def _run(maps_np, meta_ids=None):
    from acr.result_parser import ResultParser
    rp = ResultParser()
    outputs = {k: torch.from_numpy(v).cuda() for k, v in maps_np.items()}
    B = maps_np["l_center_map"].shape[0]
    meta = {"batch_ids": torch.arange(B) if meta_ids is None else meta_ids, "offsets": torch.zeros(B, 10)}
    out, meta = rp.parse(outputs, meta, {})
    return out


@pytest.mark.parametrize("case", ["both", "no_left", "mixed", "none", "far"])
def test_parse_golden(case):
    g = np.load(os.path.join(GOLDEN, "parse_golden.npz"))
    B = int(g[f"{case}__B"])
    out = _run(make_parse_case(case, B))
    for k in INT_KEYS:
        got = out[k].cpu().numpy()
        ref = g[f"{case}__{k}"]
        assert got.shape == ref.shape and (got == ref).all(), k      # bit-exact index work
    for k in F_KEYS:
        got = out[k].cpu().numpy()
        assert got.shape == g[f"{case}__{k}"].shape, k
        assert np.abs(got - g[f"{case}__{k}"]).max() < 1e-6, k
    for k in ("cam", "global_orient", "hand_pose", "betas", "poses"):
        assert np.abs(out["params_dict"][k].cpu().numpy() - g[f"{case}__pd_{k}"]).max() < 5e-5, k


@pytest.mark.parametrize("B,seed", [(1, 0), (7, 1), (256, 2), (1500, 3)])
def test_parse_random_vs_oracle(B, seed):
    from oracle import parse_ref
    g = np.random.default_rng(seed)
    maps = {}
    for s in "lr":
        cm = (g.standard_normal((B, 1, 64, 64)) * 0.12).astype(np.float32)
        on = g.random(B) < 0.7
        for b in np.nonzero(on)[0]:
            cm[b, 0, g.integers(0, 64), g.integers(0, 64)] = 0.5 + g.random()
        maps[f"{s}_center_map"] = cm
        maps[f"{s}_params_maps"] = g.standard_normal((B, 109, 64, 64)).astype(np.float32)
        maps[f"{s}_prior_maps"] = (g.standard_normal((B, 106, 64, 64)) * 0.1).astype(np.float32)
    meta_ids = torch.arange(B) * 3 + 1
    out = _run(maps, meta_ids)
    ref = parse_ref.parse(maps, meta_ids.numpy())
    for k in INT_KEYS:
        assert (out[k].cpu().numpy() == ref[k]).all(), k
    assert np.abs(out["params_pred"].cpu().numpy() - ref["params_pred"]).max() < 1e-6
    assert (out["detection_flag"].cpu().numpy() == ref["detection_flag"]).all()
    a, b = out["params_dict"]["poses"].cpu().numpy(), ref["params_dict"]["poses"]
    assert np.mean(np.abs(a - b) < 1e-4) > 0.999
