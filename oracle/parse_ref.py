"""ORACLE (test infrastructure, never on the product path).

CPU/numpy restatement of the reference's result parser for ``K = 1`` (inference),
``prior_mode='cross'``, ``inter_prior=True``, ``Rot_type='6D'``.  Pinned against
the reference through tests/golden/parse_golden.npz.

Reference functions restated (in /root/reference/acr/result_parser.py):
  CenterMap.parse_centermap_heatmap_adaptive_scale_batch  :218-243
  nms                                                     :245-249
  ResultParser.parameter_sampling                         :49-57
  ResultParser.determine_coeff                            :42-47
  ResultParser.parse_maps                                 :85-190
  ResultParser.parse                                      :21-40
The batch>1 quirks are reproduced on purpose (hard part 3):
``determine_coeff`` looks only at the first left / first right detection of the
whole batch; the prior is applied only when both sides have at least one
detection; a side with no detection contributes one dummy row sampled at
(image 0, pixel 0) with flag False.
"""
import numpy as np

from .rotation_ref import rot6d_to_angular

F = np.float32
CONF_THRESH = 0.35   # acr/config.py:131
MAP = 64             # acr/config.py:130
PART_IDX = (3, 6, 90, 10)   # cam, global_orient(6D), hand_pose(15*6D), betas  (result_parser.py:12)


def nms5(det):
    """5x5 max-pool NMS: keep a pixel iff it equals the max of its (zero... -inf padded)
    5x5 neighbourhood; everything else becomes 0 (result_parser.py:245-249)."""
    B, _, H, W = det.shape
    pad = np.full((B, H + 4, W + 4), -np.inf, F)
    pad[:, 2:-2, 2:-2] = det[:, 0]
    mx = np.full((B, H, W), -np.inf, F)
    for dy in range(5):
        for dx in range(5):
            mx = np.maximum(mx, pad[:, dy:dy + H, dx:dx + W])
    keep = (mx == det[:, 0]).astype(F)
    return det[:, 0] * keep


def parse_centers(center_map, thresh=CONF_THRESH):
    """-> batch_ids (n,), flat_inds (n,), cyxs (n,2) [y,x], scores (n,)"""
    s = nms5(np.asarray(center_map, F)).reshape(center_map.shape[0], -1)
    flat = s.argmax(1)
    score = s[np.arange(s.shape[0]), flat]
    m = score > F(thresh)
    b = np.nonzero(m)[0]
    fi = flat[m]
    return b.astype(np.int64), fi.astype(np.int64), np.stack([fi // MAP, fi % MAP], 1).astype(F), score[m]


def _sample(maps, b, fi):
    B, C = maps.shape[:2]
    return maps.reshape(B, C, -1)[b, :, fi].astype(F).copy()


def parse_maps(maps, batch_ids_meta=None):
    """maps: dict with l/r_center_map (B,1,64,64), l/r_params_maps (B,109,64,64),
    l/r_prior_maps (B,106,64,64).  Returns the reference's ``outputs`` additions."""
    lb, lf, lyx, _ = parse_centers(maps["l_center_map"])
    rb, rf, ryx, _ = parse_centers(maps["r_center_map"])
    flags = []
    if len(lb):
        flags += [True] * len(lb)
    else:
        flags.append(False)
        lb, lf = np.zeros(1, np.int64), np.zeros(1, np.int64)
    lp = _sample(maps["l_params_maps"], lb, lf)
    if len(rb):
        flags += [True] * len(rb)
    else:
        flags.append(False)
        rb, rf = np.zeros(1, np.int64), np.zeros(1, np.int64)
    rp = _sample(maps["r_params_maps"], rb, rf)

    both = np.intersect1d(lb, rb)       # unique ids that appear on both sides (K=1 => count 2)
    if len(both) > 0 and all(flags):
        lc = np.array([np.nonzero(lb == v)[0][0] for v in both])
        rc = np.array([np.nonzero(rb == v)[0][0] for v in both])
        lpr = _sample(maps["l_prior_maps"], both, rf[rc])   # left prior read at the RIGHT centre
        rpr = _sample(maps["r_prior_maps"], both, lf[lc])
        # determine_coeff: first left vs first right detection of the batch, [y,x] in 64-grid
        d = np.sqrt((lyx[0, 0] - ryx[0, 0]) ** 2 + (lyx[0, 1] - ryx[0, 1]) ** 2)
        if not d > 32:
            lp[lc, 3:] += lpr
            rp[rc, 3:] += rpr

    out = {}
    out["detection_flag"] = np.asarray(flags, F)
    out["l_params_pred"], out["r_params_pred"] = lp, rp
    out["params_pred"] = np.concatenate([lp, rp])
    bi = np.concatenate([lb, rb])
    out["l_centers_pred"] = np.stack([lf % MAP, lf // MAP], 1)
    out["r_centers_pred"] = np.stack([rf % MAP, rf // MAP], 1)
    out["l_centers_conf"] = _sample(maps["l_center_map"], lb, lf)
    out["r_centers_conf"] = _sample(maps["r_center_map"], rb, rf)
    out["left_hand_num"] = np.array([len(lp)], np.int64)
    out["right_hand_num"] = np.array([len(rp)], np.int64)
    meta = np.arange(maps["l_center_map"].shape[0]) if batch_ids_meta is None else np.asarray(batch_ids_meta)
    out["reorganize_idx"] = meta[bi]
    out["batch_ids"] = bi
    out["detection_flag_cache"] = out["detection_flag"].astype(bool)
    return out


def parse(maps, batch_ids_meta=None):
    out = parse_maps(maps, batch_ids_meta)
    p = out["params_pred"]
    o = np.cumsum((0,) + PART_IDX)
    pd = dict(cam=p[:, o[0]:o[1]].copy(), global_orient=p[:, o[1]:o[2]].copy(),
              hand_pose=p[:, o[2]:o[3]].copy(), betas=p[:, o[3]:o[4]].copy())
    pd["hand_pose"] = rot6d_to_angular(pd["hand_pose"])
    pd["global_orient"] = rot6d_to_angular(pd["global_orient"])
    pd["poses"] = np.concatenate([pd["global_orient"], pd["hand_pose"]], 1)
    L, R = int(out["left_hand_num"][0]), int(out["right_hand_num"][0])
    out["output_hand_type"] = np.concatenate([np.zeros(L), np.ones(R)]).astype(np.int32)
    out["params_dict"] = pd
    return out
