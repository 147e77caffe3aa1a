"""ORACLE (test infrastructure): import harness for the UNMODIFIED reference.

Used by tests/golden/make_golden.py (root = /root/reference, in the build container) and by oracle/ref_worker.py
(root = oracle/_ref, the snapshot that travels to the GPU box).  Steps: reference root on sys.path + chdir (MANO
root is the relative 'mano/', acr/mano_wrapper.py:22; yml path acr/config.py:24), argv set before import
(acr/config.py:232), stub modules for absent imports the hot path never executes (acr/utils.py:2,23-24,
mano/manolayer.py:2,322), np.float/np.int restored (acr/utils.py:493), ``ready_arguments`` replaced by a seeded
synthetic MANO asset (mano/manolayer.py:350-394 needs chumpy + the licence-gated pickle), ``.cuda()`` neutralised
for CPU runs (acr/model.py:35,39; acr/result_parser.py:36,...)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
PKG = os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200")


def import_reference(ref_root, argv0="ref_harness"):
    """-> torch, with the reference's ``acr`` / ``mano`` packages importable (and OUR drop-in packages of the
    same names NOT on the path: they would shadow the reference's namespace package ``mano``)."""
    sys.argv = [argv0]
    if PKG in sys.path:
        sys.path.remove(PKG)
    sys.path.insert(0, PKG)
    import acr_b200.synth  # noqa: F401   (needs only numpy / torch; stays importable through sys.modules)
    sys.path.remove(PKG)
    for m in [k for k in sys.modules if k == "acr" or k.startswith("acr.") or k == "mano" or k.startswith("mano.")]:
        del sys.modules[m]
    sys.path.insert(0, ref_root)
    os.chdir(ref_root)
    for name in ("h5py", "imgaug", "imgaug.augmenters", "chumpy", "chumpy.ch"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
    sys.modules["imgaug.augmenters"].compute_paddings_to_reach_aspect_ratio = lambda *a, **k: None
    sys.modules["chumpy"].Ch = object
    sys.modules["chumpy"].ch = sys.modules["chumpy.ch"]
    np.float = float
    np.int = int
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import acr.config  # noqa: F401  (parses argv + demo.yml at import)
    import mano.manolayer as ml
    from acr_b200.synth import make_synthetic_mano

    class _R:  # mimic chumpy's ``.r``
        def __init__(self, a):
            self.r = a

    def fake_ready_arguments(path, posekey4vposed="pose"):
        import scipy.sparse as sp
        side = "left" if "LEFT" in path else "right"
        a = make_synthetic_mano(side)
        d = {k: _R(v) for k, v in a.items() if k in ("betas", "shapedirs", "posedirs", "v_template", "weights")}
        d["hands_components"] = a["hands_components"]
        d["hands_mean"] = a["hands_mean"]
        d["J_regressor"] = sp.csc_matrix(a["J_regressor"])
        d["f"] = a["f"]
        d["kintree_table"] = a["kintree_table"]
        return d

    ml.ready_arguments = fake_ready_arguments
    return torch


def load_reference_pipeline(ref_root, sd):
    """-> run(image uint8 (B,512,512,3), offsets (B,10)) -> dict: the reference's ACR.forward (backbone, heads,
    ResultParser.parse) followed by its two ManoLayers and the projection, i.e. MANOWrapper.forward minus the host
    cv2.solvePnPRansac loop (section-8f scope; its INVALID_TRANS NameError fires for off-image joints,
    acr/utils.py:425,504) -- the same sequence tests/golden/make_golden.py uses for net_golden.npz."""
    torch = import_reference(ref_root)
    import acr.model as ref_model
    import acr.utils as ref_utils
    from acr.mano_wrapper import MANOWrapper
    model = ref_model.ACR().eval()
    model.load_state_dict(sd, strict=True)
    mw = MANOWrapper().eval()

    def run(image, offsets):
        B = image.shape[0]
        meta = {"image": image.clone(), "offsets": offsets.clone(), "batch_ids": torch.arange(B)}
        with torch.no_grad():
            out = model(meta, mode="parsing", calc_loss=False)
            Ln, Rn = int(out["left_hand_num"]), int(out["right_hand_num"])
            pd = out["params_dict"]
            lv, lj, _ = mw.mano_layer["l"](pd["poses"][:Ln], th_betas=pd["betas"][:Ln])
            rv, rj, _ = mw.mano_layer["r"](pd["poses"][Ln:Ln + Rn], th_betas=pd["betas"][Ln:Ln + Rn])
            out["verts"], out["j3d"] = torch.cat([lv, rv]), torch.cat([lj, rj])
            pj = ref_utils.batch_orth_proj(out["j3d"], pd["cam"], mode="2d")[:, :, :2]
            out["pj2d_org"] = ref_utils.convert_kp2d_from_input_to_orgimg(pj, out["meta_data"]["offsets"])
        return out
    return run
