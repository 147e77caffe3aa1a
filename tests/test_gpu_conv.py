"""GPU parity: the implicit-GEMM convolution (tcgen05) and the CUDA-core reference conv, one op at a
time through acr_b200_run_op, vs fp32 CPU convolution on identically rounded operands."""
import numpy as np
import pytest
import torch

from acr_b200 import lib as L
from tests.helpers import run_conv_case

pytestmark = pytest.mark.gpu

# (B, H, W, cin, cout, k, s, relu, residual, bias, bn, out_f32)  -- every distinct conv class of the net
CASES = [
    (2, 32, 32, 64, 64, 1, 1, True, False, False, True, False),     # 1x1, CK=64
    (2, 32, 32, 64, 64, 3, 1, True, True, False, True, False),      # BasicBlock conv2 @64ch
    (2, 64, 64, 32, 32, 3, 1, True, False, False, True, False),     # CK=32 (64B swizzle)
    (1, 32, 32, 128, 128, 3, 1, True, True, False, True, False),    # 2 channel chunks
    (2, 16, 16, 256, 256, 3, 1, True, True, False, True, False),    # N=256, smallest map
    (1, 64, 64, 64, 256, 1, 1, True, True, False, True, False),     # bottleneck expand + residual
    (1, 64, 64, 256, 64, 1, 1, True, False, False, True, False),
    (2, 64, 64, 32, 64, 3, 2, False, False, False, True, False),    # stride-2 fuse conv (no relu)
    (1, 64, 64, 256, 64, 3, 2, True, False, False, True, False),    # transition1[1]
    (2, 32, 32, 64, 128, 3, 2, True, False, False, True, False),
    (2, 64, 64, 34, 64, 3, 2, True, False, True, True, False),      # head stem: cin 34 -> CK=16 x3, bias
    (1, 32, 32, 34, 256, 3, 1, True, False, True, True, False),     # contact conv
    (1, 64, 64, 16, 64, 3, 1, True, False, True, True, False),      # CK=16 (32B swizzle)
    (1, 64, 64, 32, 16, 3, 1, True, False, True, True, False),      # N=16
    (1, 64, 64, 64, 33, 3, 1, True, False, True, True, False),      # cout 33 -> 48
    (1, 64, 64, 33, 33, 3, 1, False, False, True, False, False),    # segm logits conv (no BN, no act)
    (2, 32, 32, 64, 106, 1, 1, False, False, True, False, True),    # head final conv, fp32 out
    (2, 32, 32, 64, 1, 1, 1, False, False, True, False, True),      # centre head
    (2, 16, 16, 128, 32, 1, 1, False, False, False, True, False),   # fuse 1x1
]


def _check(kind, case, dt=L.DT_BF16):
    B, H, W, cin, cout, k, s, relu, res, bias, bn, f32 = case
    got, exp, pad_ok = run_conv_case(kind, B, H, W, cin, cout, k, s, relu, res, bias, bn, f32, dt=dt,
                                     seed=hash(case) % 1000)
    scale = exp.abs().max().item()
    err = (got - exp).abs().max().item()
    # fp32 accumulate on identical operands: only summation order + one output rounding differ
    tol = scale * (2e-5 if f32 else (2 ** -8 if dt == L.DT_BF16 else 2 ** -10)) + 1e-6
    assert pad_ok, "padding channels of the output are not zero"
    assert err <= tol, f"max err {err:.4g} > tol {tol:.4g} (scale {scale:.3g})"


@pytest.mark.parametrize("case", CASES)
def test_conv_ref(case):
    _check(L.OP_CONV_REF, case)


@pytest.mark.parametrize("case", CASES)
def test_conv_tc(case):
    _check(L.OP_CONV, case)


@pytest.mark.parametrize("case", CASES[:4])
def test_conv_tc_fp16(case):
    _check(L.OP_CONV, case, dt=L.DT_F16)


@pytest.mark.parametrize("cin,cout,k,s", [(34, 64, 3, 2), (34, 256, 3, 1), (33, 33, 3, 1), (34, 64, 1, 1)])
def test_conv_tc_k_padded_by_tma_oob(cin, cout, k, s):
    """The engine feeds 33/34-channel tensors (48-wide buffers) as ONE 64-channel K chunk: channels
    48..63 do not exist in memory and must come back as TMA out-of-bounds zeros."""
    got, exp, pad_ok = run_conv_case(L.OP_CONV, 2, 64, 64, cin, cout, k, s, True, False, True, True, False,
                                     seed=cin + cout, in_stride=48, cin_pad=64)
    assert pad_ok
    assert (got - exp).abs().max().item() <= exp.abs().max().item() * 2 ** -8 + 1e-6


def test_conv_tc_full_resolution_property():
    """BASELINE-size property check: 64->64 3x3 @128x128, B=8 -- linearity in the input
    (conv(a*x) == a*conv(x) without bias/relu) on the tensor-core path."""
    c1 = (8, 128, 128, 64, 64, 3, 1, False, False, False, False, True)
    g1, e1, _ = run_conv_case(L.OP_CONV, *c1, seed=11)
    assert (g1 - e1).abs().max().item() <= e1.abs().max().item() * 2e-5 + 1e-6
