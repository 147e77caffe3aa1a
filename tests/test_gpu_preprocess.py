"""GPU parity: device pre-processing (BGR->RGB, white pad, bicubic 512x512) -- bit-exact vs the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(360, 640), (640, 360), (600, 600), (1080, 1920)])
def test_preprocess_bit_exact(hw):
    from acr_b200.preprocess import preprocess_frames
    from oracle import preprocess_ref
    rng = np.random.default_rng(hw[1])
    frames = rng.integers(0, 256, (2, hw[0], hw[1], 3), dtype=np.uint8)
    out, offs = preprocess_frames(torch.from_numpy(frames).cuda())
    out = out.cpu().numpy()
    for i in range(2):
        ref, o = preprocess_ref.img_preprocess(frames[i])
        assert np.array_equal(out[i], ref)            # integer work: bit-exact
        assert np.array_equal(offs[i].numpy(), o)
