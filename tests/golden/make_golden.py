#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference) on seeded synthetic inputs.  Runs only in the build container (the
reference does not exist on the GPU box); the fixtures it writes are committed and are
what pins the oracle (oracle/*.py) and, through it, the CUDA path.

    python tests/golden/make_golden.py          # writes *.npz next to this file

Harness steps follow SURVEY.md section 8c: reference root on sys.path + chdir, argv set
before import, stub modules for absent imports that the hot path never executes,
``ready_arguments`` replaced by a synthetic MANO asset, ``.cuda()`` neutralised on CPU.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
PKG = os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200")
REF = "/root/reference"


def import_reference():
    """The harness itself lives in oracle/ref_harness.py (shared with the reference arm of bench.py)."""
    sys.argv_saved = list(sys.argv)
    sys.path.insert(0, ROOT)
    from oracle import ref_harness
    return ref_harness.import_reference(REF, "make_golden")


def pack_golden(torch, ref_utils):
    """reorganize_results (acr/utils.py:1226-1271) + save_results (:124-129) of the reference on a seeded synthetic
    batch: 7 hand rows over 3 images, two undetected rows, mixed hand types -> pack_golden.npz (inputs + the
    reference's per-image / per-hand fp16 payloads, flattened) and the bytes layout of the pickle dump."""
    import pickle
    import tempfile
    rng = np.random.default_rng(33)
    N = 7
    f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    outputs = {"detection_flag_cache": torch.tensor([1, 1, 0, 1, 1, 0, 1], dtype=torch.bool),
               "params_dict": {"cam": f(N, 3), "poses": f(N, 48), "betas": f(N, 10)},
               "cam_trans": f(N, 3), "j3d": f(N, 21, 3), "verts": f(N, 778, 3) * 0.1, "pj2d": f(N, 21, 2),
               "pj2d_org": f(N, 21, 2) * 300, "output_hand_type": torch.tensor([0, 0, 0, 1, 1, 1, 0], dtype=torch.int32)}
    # rows surviving the detection filter keep their order: reorganize_idx / img_paths index the FILTERED rows
    reorganize_idx = np.array([0, 2, 0, 1, 2])
    img_paths = ["a.jpg", "c.jpg", "a.jpg", "b.jpg", "c.jpg"]
    res = ref_utils.reorganize_results(outputs, img_paths, reorganize_idx)
    flat = {"reorganize_idx": reorganize_idx, "img_paths": np.array(img_paths),
            "detection_flag_cache": outputs["detection_flag_cache"].numpy(), "output_hand_type": outputs["output_hand_type"].numpy(),
            "cam_trans": outputs["cam_trans"].numpy(), "j3d": outputs["j3d"].numpy(), "verts": outputs["verts"].numpy(),
            "pj2d": outputs["pj2d"].numpy(), "pj2d_org": outputs["pj2d_org"].numpy()}
    for k, v in outputs["params_dict"].items():
        flat["pd_" + k] = v.numpy()
    flat["result_keys"] = np.array(list(res.keys()))
    for name, hands in res.items():
        flat[f"res__{name}__n"] = len(hands)
        for i, hd in enumerate(hands):
            for k, v in hd.items():
                flat[f"res__{name}__{i}__{k}"] = np.asarray(v)
    # save_results: file name convention + pickle of the dict
    import acr.config as ref_cfg
    with tempfile.TemporaryDirectory() as d:
        ref_utils.save_results("some/folder/clip7", d, res)
        names = os.listdir(d)
        assert len(names) == 1
        flat["save_name"] = np.array(names[0])
        flat["save_model_path"] = np.array(ref_cfg.args().model_path)
        flat["save_conf"] = np.array(ref_cfg.args().centermap_conf_thresh)
        with open(os.path.join(d, names[0]), "rb") as fh:
            back = pickle.load(fh)
        assert list(back.keys()) == list(res.keys())
    np.savez_compressed(os.path.join(HERE, "pack_golden.npz"), **flat)
    print("pack_golden:", {k: len(v) for k, v in res.items()}, str(flat["save_name"]))


def main():
    torch = import_reference()
    torch.manual_seed(0)
    from acr_b200.synth import synth_state_dict
    import acr.model as ref_model
    import acr.utils as ref_utils
    from acr.mano_wrapper import MANOWrapper
    from mano.manolayer import batch_rodrigues
    if "--only-pack" in sys.argv_saved:
        return pack_golden(torch, ref_utils)

    # ------------------------------------------------------------------ rotations
    rng = np.random.default_rng(7)
    r6 = rng.standard_normal((64, 6)).astype(np.float32)
    r6[0] = [1, 0, 0, 1, 0, 0]                    # identity
    r6[1] = [1, 0, 0, -1, 0, 0]                   # 180 deg about x  (trace = -1)
    r6[2] = [-1, 0, 0, 1, 0, 0]                   # 180 deg about y
    r6[3] = [-1, 0, 0, -1, 0, 0]                  # 180 deg about z
    r6[4] = [1e-9, 0, 0, 0, 0, 1e-9]              # degenerate (below normalize eps)
    r6[5] = [1, 1, 2, 2, 3, 3]                    # parallel columns
    r6[6] = [0, 0, 0, 0, 0, 0]
    aa = ref_utils.rot6D_to_angular(torch.from_numpy(r6.reshape(4, 96))).numpy()
    aa_in = (rng.standard_normal((40, 3)) * 1.2).astype(np.float32)
    aa_in[0] = 0
    aa_in[1] = [np.pi, 0, 0]
    aa_in[2] = [1e-7, -1e-7, 1e-7]
    aa_in[3] = [0, 3.1415925, 0]
    rod = batch_rodrigues(torch.from_numpy(aa_in)).numpy()
    np.savez_compressed(os.path.join(HERE, "rot_golden.npz"), rot6d=r6.reshape(4, 96), aa=aa,
                        aa_in=aa_in, rodrigues=rod)

    # ------------------------------------------------------- temporal smoothing (acr/utils.py:1466-1527)
    filt = {0: ref_utils.create_OneEuroFilter(4.0), 1: ref_utils.create_OneEuroFilter(4.0)}
    T = 8
    rng_s = np.random.default_rng(21)       # own stream: the other fixtures keep their inputs
    seq_p = (rng_s.standard_normal((T, 2, 48)) * 0.3).astype(np.float32) + (rng_s.standard_normal((1, 2, 48)) * 0.5).astype(np.float32)
    seq_b = (rng_s.standard_normal((T, 2, 10)) * 0.2).astype(np.float32)
    det = np.ones((T, 2), np.float32)
    det[3, 0] = 0                       # left hand lost in frame 3: not filtered, history untouched
    out_p, out_b = seq_p.copy(), seq_b.copy()
    for t in range(T):
        for sid in range(2):
            if det[t, sid] == 0:
                continue
            pp, bb = ref_utils.smooth_results(filt[sid], torch.from_numpy(seq_p[t, sid].copy()), torch.from_numpy(seq_b[t, sid].copy()))
            out_p[t, sid], out_b[t, sid] = pp.numpy(), bb.numpy()
    np.savez_compressed(os.path.join(HERE, "smooth_golden.npz"), poses=seq_p, betas=seq_b, det=det, out_poses=out_p, out_betas=out_b)

    # ----------------------------------------------------------------------- MANO
    mw = MANOWrapper().eval()
    n = 10
    poses = (rng.standard_normal((n, 48)) * 0.5).astype(np.float32)
    betas = rng.standard_normal((n, 10)).astype(np.float32)
    poses[0] = 0
    betas[0] = 0
    poses[1, :3] = [np.pi, 0, 0]
    poses[2] *= 4.0
    cam = np.stack([rng.uniform(0.5, 1.5, n), rng.uniform(-.5, .5, n), rng.uniform(-.5, .5, n)], 1).astype(np.float32)
    offsets = np.tile(np.array([512, 512, 0, 0, 0, 0, 0, 0, 0, 0], np.float32), (n, 1))
    offsets[3] = [1920, 1920, 0, 0, 0, 0, 420, 0, 420, 0]
    L, R = 4, 6
    with torch.no_grad():
        lv, lj, lc = mw.mano_layer["l"](torch.from_numpy(poses[:L]), th_betas=torch.from_numpy(betas[:L]))
        rv, rj, rc = mw.mano_layer["r"](torch.from_numpy(poses[L:]), th_betas=torch.from_numpy(betas[L:]))
        verts, j3d = torch.cat([lv, rv]), torch.cat([lj, rj])
        vc = ref_utils.batch_orth_proj(verts, torch.from_numpy(cam), mode="3d", keep_dim=True)
        pj = ref_utils.batch_orth_proj(j3d, torch.from_numpy(cam), mode="2d")[:, :, :2]
        pjo = ref_utils.convert_kp2d_from_input_to_orgimg(pj, torch.from_numpy(offsets))
    # camera translation: the reference's closed-form estimator (its fall-back for cv2.solvePnPRansac),
    # called exactly like estimate_translation does (acr/utils.py:404-406, 489-517)
    j3d_np, j2d_np = j3d.numpy(), (pj.numpy() + 1) * 256
    ct = np.zeros((n, 3), np.float32)
    for i in range(n):
        m = (j2d_np[i, :, -1] > -2.) * (j3d_np[i, :, -1] != -2.)
        ct[i] = ref_utils.estimate_translation_np(j3d_np[i][m], j2d_np[i][m], m[m].astype(np.float32),
                                                  focal_length=1265, img_size=np.array([512., 512.]))
    np.savez_compressed(os.path.join(HERE, "mano_golden.npz"), poses=poses, betas=betas, cam=cam, offsets=offsets,
                        L=L, R=R, verts=verts.numpy(), j3d=j3d.numpy(), center=torch.cat([lc, rc]).numpy(),
                        verts_camed=vc.numpy(), pj2d=pj.numpy(), pj2d_org=pjo.numpy(), cam_trans=ct)

    # ------------------------------------------------------------- parser on synthetic maps
    from acr.result_parser import ResultParser
    rp = ResultParser()
    cases = {}
    for name, B, kill in (("both", 3, {}), ("no_left", 2, {"l": [0, 1]}), ("mixed", 4, {"l": [1], "r": [0, 3]}),
                          ("none", 2, {"l": [0, 1], "r": [0, 1]}), ("far", 1, {})):
        g = np.random.default_rng({"both": 1, "no_left": 2, "mixed": 3, "none": 4, "far": 5}[name])
        maps = {}
        for s in "lr":
            cm = (g.standard_normal((B, 1, 64, 64)) * 0.1).astype(np.float32)
            for b in range(B):
                if b in kill.get(s, []):
                    continue
                y, x = g.integers(0, 64, 2)
                if name == "far":
                    y, x = (2, 3) if s == "l" else (60, 58)
                cm[b, 0, y, x] = 0.9 + 0.05 * b
                if 0 < y < 63:
                    cm[b, 0, y + 1, x] = 0.8          # suppressed by NMS
            maps[f"{s}_center_map"] = cm
            maps[f"{s}_params_maps"] = g.standard_normal((B, 109, 64, 64)).astype(np.float32)
            maps[f"{s}_prior_maps"] = (g.standard_normal((B, 106, 64, 64)) * 0.1).astype(np.float32)
        outs = {k: torch.from_numpy(v.copy()) for k, v in maps.items()}
        meta = {"batch_ids": torch.arange(B), "offsets": torch.zeros(B, 10), "image": torch.zeros(B, 1)}
        o, _ = rp.parse(outs, meta, {})
        keep = {}
        for k in ("params_pred", "detection_flag", "reorganize_idx", "l_centers_pred", "r_centers_pred",
                  "l_centers_conf", "r_centers_conf", "left_hand_num", "right_hand_num", "output_hand_type"):
            keep[k] = o[k].numpy()
        for k, v in o["params_dict"].items():
            keep["pd_" + k] = v.numpy()
        cases[name] = (B, keep)
    flat = {}
    for name, (B, keep) in cases.items():
        flat[f"{name}__B"] = B
        for k, v in keep.items():
            flat[f"{name}__{k}"] = v
    np.savez_compressed(os.path.join(HERE, "parse_golden.npz"), **flat)

    # --------------------------------------------------- full network, calibrated BN stats
    sd = synth_state_dict(0)
    model = ref_model.ACR().eval()
    missing = model.load_state_dict(sd, strict=True)
    print("state-dict keys match the reference:", missing, len(sd))
    gi = torch.Generator().manual_seed(0)
    calib = torch.randint(0, 256, (2, 512, 512, 3), generator=gi, dtype=torch.uint8)
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = 1.0
    model.train()
    with torch.no_grad():
        model.head_forward(model.backbone(calib))
    model.eval()
    stats = {k: v.numpy() for k, v in model.state_dict().items()
             if k.endswith(("running_mean", "running_var"))}
    np.savez_compressed(os.path.join(HERE, "bn_calib_seed0.npz"), **stats)
    sd = synth_state_dict(0, bn_stats=stats)
    model.load_state_dict(sd, strict=True)
    model.eval()

    gi = torch.Generator().manual_seed(123)
    img = torch.randint(0, 256, (2, 512, 512, 3), generator=gi, dtype=torch.uint8)
    meta = {"image": img.clone(), "offsets": torch.tensor([[512., 512, 0, 0, 0, 0, 0, 0, 0, 0]] * 2),
            "batch_ids": torch.arange(2)}
    with torch.no_grad():
        x = model.backbone(img)
        heads = model.head_forward(x)
        out = model(meta, mode="parsing", calc_loss=False)
        # MANOWrapper.forward minus the host cv2 PnP (cam_trans is section-8f scope and the
        # reference's INVALID_TRANS NameError fires for off-image joints, acr/utils.py:425,504)
        Ln, Rn = int(out["left_hand_num"]), int(out["right_hand_num"])
        pd = out["params_dict"]
        lv, lj, _ = mw.mano_layer["l"](pd["poses"][:Ln], th_betas=pd["betas"][:Ln])
        rv, rj, _ = mw.mano_layer["r"](pd["poses"][Ln:Ln + Rn], th_betas=pd["betas"][Ln:Ln + Rn])
        out["verts"], out["j3d"] = torch.cat([lv, rv]), torch.cat([lj, rj])
        pj = ref_utils.batch_orth_proj(out["j3d"], pd["cam"], mode="2d")[:, :, :2]
        out["pj2d_org"] = ref_utils.convert_kp2d_from_input_to_orgimg(pj, out["meta_data"]["offsets"])
    g = dict(backbone_mean=x.mean().item(), backbone_std=x.std().item(),
             backbone_crop=x[:, :, 60:68, 60:68].numpy(),
             segms_crop=heads["segms"][:, :, 100:108, 100:108].numpy(),
             l_center_map=heads["l_center_map"].numpy(), r_center_map=heads["r_center_map"].numpy(),
             l_params_crop=heads["l_params_maps"][:, :, 30:34, 30:34].numpy(),
             r_params_crop=heads["r_params_maps"][:, :, 30:34, 30:34].numpy(),
             l_prior_crop=heads["l_prior_maps"][:, :, 30:34, 30:34].numpy(),
             params_pred=out["params_pred"].numpy(), detection_flag=out["detection_flag"].numpy(),
             l_centers_pred=out["l_centers_pred"].numpy(), r_centers_pred=out["r_centers_pred"].numpy(),
             poses=out["params_dict"]["poses"].numpy(), betas=out["params_dict"]["betas"].numpy(),
             cam=out["params_dict"]["cam"].numpy(), verts=out["verts"].numpy(), j3d=out["j3d"].numpy(),
             pj2d_org=out["pj2d_org"].numpy(), reorganize_idx=out["reorganize_idx"].numpy())
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), **g)
    pack_golden(torch, ref_utils)
    for k, v in g.items():
        print(k, getattr(v, "shape", v))
    print("backbone mean/std", g["backbone_mean"], g["backbone_std"])


if __name__ == "__main__":
    main()
