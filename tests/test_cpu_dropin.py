"""CPU: host-side mirror of the reference interface -- state-dict / checkpoint compatibility, the
chumpy-free MANO pickle reader, config shim.  No GPU compute."""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from acr_b200.synth import make_synthetic_mano, synth_state_dict


def test_model_state_dict_is_checkpoint_compatible(tmp_path):
    """Reference checkpoints store keys as 'module.'+key inside {'model_state_dict': ...}
    (acr/utils.py:1106-1168); load_model must fill every one of the 2067 tensors."""
    from acr.model import ACR
    from acr.utils import load_model
    sd = synth_state_dict(5)
    ckpt = {"model_state_dict": {"module." + k: v.clone() for k, v in sd.items()}}
    path = os.path.join(tmp_path, "wild.pkl")
    torch.save(ckpt, path)
    model = load_model(path, ACR().eval(), prefix="module.", drop_prefix="")
    got = model.state_dict()
    assert set(got.keys()) == set(sd.keys()) and len(got) == 2067
    for k in ("backbone.conv1.weight", "backbone.stage4.2.fuse_layers.0.3.1.running_var", "contact_layers.5.bias",
              "l_final_layers.4.1.1.0.bn2.weight", "cam_shape_layers.3.weight"):
        assert torch.equal(got[k], sd[k]), k
    with pytest.raises(ValueError):
        load_model(os.path.join(tmp_path, "missing.pkl"), model)


def test_model_exposes_reference_attributes():
    from acr.model import ACR
    m = ACR()
    for attr in ("backbone", "l_final_layers", "r_final_layers", "contact_layers", "cam_shape_layers",
                 "segmentation_layers", "_result_parser"):
        assert hasattr(m, attr), attr
    assert hasattr(m.backbone, "hand_segm")
    assert m._result_parser.params_num == 109
    with pytest.raises(RuntimeError):          # no CPU fallback: forward needs .cuda()
        m({"image": torch.zeros(1, 512, 512, 3, dtype=torch.uint8), "offsets": torch.zeros(1, 10),
           "batch_ids": torch.arange(1)})


def test_mano_pickle_reader_without_chumpy(tmp_path):
    """MANO_*.pkl holds chumpy objects; the reader must get the arrays out with chumpy absent."""
    a = make_synthetic_mano("right")
    chumpy = types.ModuleType("chumpy")
    ch = types.ModuleType("chumpy.ch")

    Ch = type("Ch", (object,), {"__init__": lambda self, x: setattr(self, "x", x), "__module__": "chumpy.ch",
                                 "__qualname__": "Ch"})
    ch.Ch = Ch
    chumpy.ch = ch
    sys.modules["chumpy"], sys.modules["chumpy.ch"] = chumpy, ch
    try:
        import scipy.sparse as sp
        dd = {k: (Ch(v) if k in ("shapedirs", "posedirs", "v_template", "weights") else v) for k, v in a.items()
              if k != "side"}
        dd["J_regressor"] = sp.csc_matrix(a["J_regressor"])
        path = os.path.join(tmp_path, "MANO_RIGHT.pkl")
        with open(path, "wb") as f:
            pickle.dump(dd, f, protocol=2)
    finally:
        del sys.modules["chumpy"], sys.modules["chumpy.ch"]
    from mano.assets import get_asset, load_mano_pkl
    got = load_mano_pkl(path)
    for k in ("shapedirs", "posedirs", "v_template", "weights", "J_regressor", "hands_mean", "hands_components"):
        assert np.array_equal(got[k], a[k]), k
    assert got["f"].dtype == np.int64 and got["f"].shape == (1538, 3)
    assert np.array_equal(get_asset(str(tmp_path), "right")["posedirs"], a["posedirs"])
    os.environ.pop("ACR_B200_SYNTHETIC_MANO", None)
    with pytest.raises(FileNotFoundError):
        get_asset(str(tmp_path), "left")


def test_manolayer_buffers_match_reference_shapes():
    from mano.manolayer import ManoLayer
    layer = ManoLayer(ncomps=45, center_idx=9, side="left", use_pca=False, flat_hand_mean=False,
                      asset=make_synthetic_mano("left"))
    shapes = {"th_betas": (1, 10), "th_shapedirs": (778, 3, 10), "th_posedirs": (778, 3, 135),
              "th_v_template": (1, 778, 3), "th_J_regressor": (16, 778), "th_weights": (778, 16),
              "th_faces": (1538, 3), "th_hands_mean": (1, 45), "th_comps": (45, 45), "th_selected_comps": (45, 45)}
    for k, s in shapes.items():                       # SURVEY.md 8b / mano/manolayer.py:65-93
        assert tuple(getattr(layer, k).shape) == s, k
    assert layer.th_faces.dtype == torch.int64
    with pytest.raises(NotImplementedError):
        ManoLayer(root_rot_mode="rot6d", asset=make_synthetic_mano("left"))


def test_config_shim():
    from acr.config import ConfigContext, args, parse_args
    a = args()
    assert (a.centermap_size, a.centermap_conf_thresh, a.align_idx, a.rot_dim, a.cam_dim) == (64, 0.35, 9, 6, 3)
    ns = parse_args(["--model_precision", "fp16", "--centermap_conf_thresh", "0.5", "--unknown_flag", "1"])
    assert ns.model_precision == "fp16" and ns.centermap_conf_thresh == 0.5
    with ConfigContext(ns) as cur:
        assert args() is cur and args().model_precision == "fp16"
    ConfigContext(parse_args([]))
    assert args().model_precision == "bf16"


def test_reorganize_results_packaging():
    """acr.utils.reorganize_results (acr/utils.py:1226-1271): per-image list of per-hand dicts, fp16 payloads."""
    from acr.utils import reorganize_results
    n = 4
    g = torch.Generator().manual_seed(0)
    out = {"detection_flag_cache": torch.ones(n, dtype=torch.bool), "cam_trans": torch.randn(n, 3, generator=g),
           "j3d": torch.randn(n, 21, 3, generator=g), "verts": torch.randn(n, 778, 3, generator=g),
           "pj2d": torch.randn(n, 21, 2, generator=g), "pj2d_org": torch.randn(n, 21, 2, generator=g),
           "output_hand_type": torch.tensor([0, 0, 1, 1], dtype=torch.int32),
           "params_dict": {"cam": torch.randn(n, 3, generator=g), "poses": torch.randn(n, 48, generator=g),
                           "betas": torch.randn(n, 10, generator=g)}}
    idx = np.array([0, 1, 0, 1])
    res = reorganize_results(out, ["a.jpg", "b.jpg", "a.jpg", "b.jpg"], idx)
    assert set(res) == {"a.jpg", "b.jpg"} and len(res["a.jpg"]) == 2
    assert res["a.jpg"][0]["hand_type"] == 0 and res["a.jpg"][1]["hand_type"] == 1
    assert res["b.jpg"][1]["verts"].dtype == np.float16 and res["b.jpg"][1]["verts"].shape == (778, 3)
    assert np.array_equal(res["b.jpg"][0]["poses"], out["params_dict"]["poses"][1].numpy().astype(np.float16))


def test_reorganize_and_save_results_match_the_reference_golden(tmp_path):
    """Result packaging: reorganize_results / save_results against pack_golden.npz, written by the
    unmodified reference (acr/utils.py:1226-1271, 124-129) on a seeded batch with undetected rows and mixed hands."""
    import os
    import pickle

    import numpy as np
    import torch

    from acr.config import args
    from acr.utils import reorganize_results, save_results
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pack_golden.npz"))
    outputs = {"detection_flag_cache": torch.from_numpy(g["detection_flag_cache"]),
               "params_dict": {k: torch.from_numpy(g["pd_" + k]) for k in ("cam", "poses", "betas")},
               "output_hand_type": torch.from_numpy(g["output_hand_type"])}
    for k in ("cam_trans", "j3d", "verts", "pj2d", "pj2d_org"):
        outputs[k] = torch.from_numpy(g[k])
    res = reorganize_results(outputs, [str(p) for p in g["img_paths"]], g["reorganize_idx"])
    assert list(res.keys()) == [str(k) for k in g["result_keys"]]
    for name, hands in res.items():
        assert len(hands) == int(g[f"res__{name}__n"])
        for i, hd in enumerate(hands):
            keys = sorted(k.split("__")[-1] for k in g.files if k.startswith(f"res__{name}__{i}__"))
            assert sorted(hd.keys()) == keys
            for k, v in hd.items():
                ref = g[f"res__{name}__{i}__{k}"]
                assert np.asarray(v).dtype == ref.dtype and np.array_equal(np.asarray(v), ref), (name, i, k)
    old = args().model_path
    args().model_path = str(g["save_model_path"])
    try:
        save_results("some/folder/clip7", str(tmp_path), res)
    finally:
        args().model_path = old
    assert os.listdir(tmp_path) == [str(g["save_name"])]
    with open(tmp_path / str(g["save_name"]), "rb") as f:
        back = pickle.load(f)
    assert list(back.keys()) == list(res.keys()) and np.array_equal(back["a.jpg"][0]["verts"], res["a.jpg"][0]["verts"])
