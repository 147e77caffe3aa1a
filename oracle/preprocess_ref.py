"""ORACLE (test infrastructure, never on the product path).

numpy restatement of the reference's frame pre-processing (acr/utils.py:1303-1337): BGR->RGB, white pad to a
square (imgaug compute_paddings_to_reach_aspect_ratio, pad_cval=255), cv2.resize(INTER_CUBIC) to 512x512, and
the offsets vector.  The resize follows OpenCV's generic 8-bit cubic path (resize.cpp, 11-bit fixed point);
it is pinned against cv2 itself in tests/test_oracle_golden.py (cv2 is a dependency of the reference, present
in this image): identical up to <=1 grey level on <1e-4 of the pixels with IPP off, <=1 grey level on ~5 % of
the pixels against the IPP-dispatched build.  imgaug is absent: its padding rule is restated from its
published source (imgaug 0.4.0) -- that split is the one "parity unpinned" item of this file."""
import numpy as np

F = np.float32


def paddings_to_square(h, w):
    top = right = bottom = left = 0
    if w > h:
        d = w - h
        top, bottom = d // 2, d - d // 2
    elif h > w:
        d = h - w
        left, right = d // 2, d - d // 2
    return top, right, bottom, left


def cubic_tables(n_src, n_dst):
    scale = np.float64(n_src) / n_dst
    A = F(-0.75)
    coef = np.zeros((n_dst, 4), np.int64)
    ofs = np.zeros(n_dst, np.int64)
    for d in range(n_dst):
        fx = F((d + 0.5) * scale - 0.5)
        s = int(np.floor(fx))
        x = F(fx - s)
        c = np.zeros(4, F)
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        c[3] = F(1.0) - c[0] - c[1] - c[2]
        ofs[d] = s
        coef[d] = np.rint(c * F(2048)).astype(np.int64)
    return coef, ofs


def resize_cubic_u8(img, dsize):
    """(H,W,C) uint8 -> (dsize,dsize,C) uint8, OpenCV generic fixed-point cubic."""
    H, W, _ = img.shape
    cx, ox = cubic_tables(W, dsize)
    cy, oy = cubic_tables(H, dsize)
    src = img.astype(np.int64)
    hor = np.zeros((H, dsize, img.shape[2]), np.int64)
    for k in range(4):
        hor += src[:, np.clip(ox + k - 1, 0, W - 1), :] * cx[:, k][None, :, None]
    acc = np.zeros((dsize, dsize, img.shape[2]), np.int64)
    for j in range(4):
        acc += hor[np.clip(oy + j - 1, 0, H - 1)] * cy[:, j][:, None, None]
    return np.clip((acc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)


def img_preprocess(frame_bgr, input_size=512):
    """-> (image (S,S,3) uint8 RGB, offsets (10,) float32)   [acr/utils.py:1315-1337]"""
    rgb = frame_bgr[:, :, ::-1]
    h, w = rgb.shape[:2]
    t, r, b, l = paddings_to_square(h, w)
    padded = np.full((h + t + b, w + l + r, 3), 255, np.uint8)
    padded[t:t + h, l:l + w] = rgb
    offsets = np.array([padded.shape[0], padded.shape[1], 0, 0, 0, 0, t, r, b, l], F)
    return resize_cubic_u8(padded, input_size), offsets
