"""CPU: size-independent properties of the oracle restatements (beyond the reference goldens): the same
properties the GPU tests use at sizes the oracle cannot reach."""
import numpy as np
import pytest

from oracle import mano_ref, parse_ref, rotation_ref
from acr_b200.synth import make_synthetic_mano


def test_rodrigues_is_a_rotation_and_inverts_with_the_angle():
    rng = np.random.default_rng(0)
    aa = rng.standard_normal((256, 3)).astype(np.float32)
    R = rotation_ref.batch_rodrigues(aa).reshape(-1, 3, 3).astype(np.float64)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 5e-6
    assert np.abs(np.linalg.det(R) - 1).max() < 5e-6
    Rn = rotation_ref.batch_rodrigues(-aa).reshape(-1, 3, 3).astype(np.float64)
    assert np.abs(Rn - R.transpose(0, 2, 1)).max() < 5e-6


def test_rot6d_round_trip_through_axis_angle():
    """6D -> Gram-Schmidt matrix -> quaternion -> axis-angle -> Rodrigues gives the Gram-Schmidt matrix back."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((512, 6)).astype(np.float32)
    R = rotation_ref.rot6d_to_rotmat(x)
    aa = rotation_ref.rot6d_to_angular(x)
    R2 = rotation_ref.batch_rodrigues(aa).reshape(-1, 3, 3)
    assert np.abs(R2 - R).max() < 2e-5
    assert (np.linalg.norm(aa, axis=1) <= np.pi + 1e-4).all()


def test_mano_rigid_motion_and_shape_linearity():
    """Centred outputs: a global rotation rotates vertices and joints rigidly; with the pose fixed the
    (uncentred) skinning is affine in the joint positions, so zero hand pose is linear in betas."""
    rng = np.random.default_rng(2)
    a = make_synthetic_mano("right")
    n = 6
    pose = (rng.standard_normal((n, 48)) * 0.3).astype(np.float32)
    betas = rng.standard_normal((n, 10)).astype(np.float32)
    v0, j0, _ = mano_ref.mano_forward(a, np.concatenate([np.zeros((n, 3), np.float32), pose[:, 3:]], 1), betas, "right")
    v1, j1, _ = mano_ref.mano_forward(a, pose, betas, "right")
    R = rotation_ref.batch_rodrigues(pose[:, :3]).reshape(n, 3, 3)
    assert np.abs(v1 - np.einsum("nab,nvb->nva", R, v0)).max() < 2e-5
    assert np.abs(j1 - np.einsum("nab,njb->nja", R, j0)).max() < 2e-5
    # flat hand (pose = -mean so that the full pose is zero), no centring: vertices are affine in betas
    flat = np.concatenate([np.zeros((1, 3), np.float32), -a["hands_mean"][None].astype(np.float32)], 1)
    f = lambda b: mano_ref.mano_forward(a, flat, b[None].astype(np.float32), "right", center_idx=None)[0][0].astype(np.float64)
    b1, b2 = betas[0], betas[1]
    assert np.abs(f(b1 + b2) - f(b1) - f(b2) + f(np.zeros(10))).max() < 5e-6
    # the left layer as ACR configures it mirrors the shape directions in x (acr/mano_wrapper.py:35)
    al = make_synthetic_mano("left")
    vl, _, _ = mano_ref.mano_forward(al, flat, b1[None], "left", center_idx=None)
    vl_noflip, _, _ = mano_ref.mano_forward(al, flat, b1[None], "left", center_idx=None, flip_shapedirs_x=False)
    assert np.abs(vl - vl_noflip).max() > 1e-4


def test_projection_round_trip():
    rng = np.random.default_rng(3)
    verts = rng.standard_normal((4, 778, 3)).astype(np.float32) * 0.1
    j3d = rng.standard_normal((4, 21, 3)).astype(np.float32) * 0.1
    cam = np.array([[1.5, 0.1, -0.2]] * 4, np.float32)
    offs = np.tile(np.array([512, 512, 0, 0, 0, 0, 0, 0, 0, 0], np.float32), (4, 1))
    out = mano_ref.project(verts, j3d, cam, offs)
    assert np.abs((out["pj2d"] - cam[:, None, 1:]) / cam[:, None, :1] - j3d[:, :, :2]).max() < 1e-6
    assert np.abs(out["pj2d_org"] - (out["pj2d"] + 1) * 256).max() < 1e-4
    assert (out["verts_camed"][:, :, 2] == verts[:, :, 2]).all()


def _maps(B, rng, peak=None):
    m = {}
    for s in "lr":
        c = (rng.random((B, 1, 64, 64)) * 0.2).astype(np.float32)
        if peak is not None:
            for b, (y, x, v) in enumerate(peak[s]):
                c[b, 0, y, x] = v
        m[f"{s}_center_map"] = c
        m[f"{s}_params_maps"] = rng.standard_normal((B, 109, 64, 64)).astype(np.float32)
        m[f"{s}_prior_maps"] = rng.standard_normal((B, 106, 64, 64)).astype(np.float32)
    return m


def test_parse_finds_the_planted_centres_and_samples_their_columns():
    rng = np.random.default_rng(4)
    B = 3
    peak = {"l": [(10, 20, 0.9), (5, 6, 0.1), (63, 0, 0.8)], "r": [(30, 31, 0.7), (40, 41, 0.95), (0, 63, 0.2)]}
    maps = _maps(B, rng, peak)
    out = parse_ref.parse(maps)
    # frames 0 and 2 have a left hand, frames 0 and 1 a right hand (threshold 0.35 of the centre map arg-max)
    assert int(out["left_hand_num"][0]) == 2 and int(out["right_hand_num"][0]) == 2
    # centres come back as (x, y) on the 64-grid (acr/result_parser.py:205-210)
    assert (np.asarray(out["l_centers_pred"]) == np.array([[20, 10], [0, 63]])).all()
    assert (np.asarray(out["r_centers_pred"]) == np.array([[31, 30], [41, 40]])).all()
    assert (np.asarray(out["reorganize_idx"]) == np.array([0, 2, 0, 1])).all()
    assert (np.asarray(out["output_hand_type"]) == np.array([0, 0, 1, 1])).all()
    # idempotence: parsing twice gives the same rows
    out2 = parse_ref.parse(maps)
    assert (np.asarray(out2["params_pred"]) == np.asarray(out["params_pred"])).all()


@pytest.mark.parametrize("B", [1, 5])
def test_parse_without_detections_falls_back_like_the_reference(B):
    rng = np.random.default_rng(6)
    out = parse_ref.parse(_maps(B, rng))
    # acr/result_parser.py:83-98: no centre above the threshold -> one dummy row per side, detection_flag False
    assert not np.asarray(out["detection_flag"]).any()
    assert np.asarray(out["params_pred"]).shape[1] == 109
