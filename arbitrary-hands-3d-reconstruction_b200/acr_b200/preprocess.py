"""Host side of the device pre-processing (csrc/preprocess.cu): padding geometry, offsets vector and the
fixed-point cubic tables, built exactly like the third-party code the reference calls.

* padding split: imgaug 0.4.0 ``compute_paddings_to_reach_aspect_ratio`` (absent from this image; restated
  from its published source): a landscape frame gets floor(diff/2) rows on top and ceil(diff/2) below, a
  portrait frame floor(diff/2) columns left and ceil(diff/2) right.
* offsets vector: ``[padded_h, padded_w, crop t,r,b,l (= 0), pad t,r,b,l]`` (acr/utils.py:1301,1311).
* cubic tables: opencv-python ``resize.cpp`` (INTER_CUBIC, 8-bit): ``fx = (d+0.5)*scale-0.5`` in float32,
  ``A = -0.75``, coefficients ``round(c*2048)`` as int16.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

from . import lib as L


def paddings_to_square(h: int, w: int) -> Tuple[int, int, int, int]:
    """(top, right, bottom, left) that make an (h, w) frame square."""
    top = right = bottom = left = 0
    if w > h:
        d = w - h
        top, bottom = d // 2, d - d // 2
    elif h > w:
        d = h - w
        left, right = d // 2, d - d // 2
    return top, right, bottom, left


def offsets_vector(h: int, w: int) -> np.ndarray:
    t, r, b, l = paddings_to_square(h, w)
    return np.array([h + t + b, w + l + r, 0, 0, 0, 0, t, r, b, l], np.float32)


@lru_cache(maxsize=32)
def cubic_tables(n_src: int, n_dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """-> (coef (n_dst,4) int16, ofs (n_dst) int32): taps ofs-1 .. ofs+2 of the source axis."""
    scale = np.float64(n_src) / n_dst
    A = np.float32(-0.75)
    coef = np.zeros((n_dst, 4), np.int16)
    ofs = np.zeros(n_dst, np.int32)
    for d in range(n_dst):
        fx = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(fx))
        x = np.float32(fx - s)
        c = np.zeros(4, np.float32)
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        c[3] = np.float32(1.0) - c[0] - c[1] - c[2]
        ofs[d] = s
        coef[d] = np.rint(c * np.float32(2048)).astype(np.int16)
    return coef, ofs


_dev_tables = {}


def preprocess_frames(frames_bgr: torch.Tensor, input_size: int = 512):
    """(n,H,W,3) uint8 BGR CUDA tensor -> ((n,S,S,3) uint8 RGB CUDA tensor, (n,10) offsets)."""
    L.require_cuda(frames_bgr)
    assert frames_bgr.dtype == torch.uint8 and frames_bgr.dim() == 4 and frames_bgr.shape[-1] == 3
    frames_bgr = frames_bgr.contiguous()
    n, H, W, _ = frames_bgr.shape
    t, r, b, l = paddings_to_square(H, W)
    side = max(H, W)
    key = (side, input_size, str(frames_bgr.device))
    if key not in _dev_tables:
        coef, ofs = cubic_tables(side, input_size)
        _dev_tables[key] = (torch.from_numpy(coef).to(frames_bgr.device), torch.from_numpy(ofs).to(frames_bgr.device))
    coef, ofs = _dev_tables[key]
    out = torch.empty(n, input_size, input_size, 3, dtype=torch.uint8, device=frames_bgr.device)
    L.check(L.load().acr_b200_preprocess(L.ptr(frames_bgr), n, H, W, L.ptr(coef), L.ptr(ofs), L.ptr(coef), L.ptr(ofs),
                                         side, t, l, input_size, L.ptr(out), L.current_stream()), "preprocess")
    offsets = torch.from_numpy(np.tile(offsets_vector(H, W), (n, 1)))
    return out, offsets
