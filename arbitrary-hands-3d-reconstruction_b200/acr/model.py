"""Drop-in ``acr.model.ACR`` (reference: /root/reference/acr/model.py:23-329, backbone :691-881).

Same constructor, ``forward(meta_data, **cfg)`` / ``head_forward(x)`` signatures, ``state_dict()``
keys (2067 tensors, checkpoint compatible) and output dict schema.  The module tree only *holds*
the parameters; the arithmetic is a precompiled CUDA launch plan (acr_b200.engine.Engine).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from acr.config import args
from acr.result_parser import ResultParser
from acr_b200.engine import Engine
from acr_b200.netspec import build_acr_spec

BN_MOMENTUM = 0.1
_MAP_KEYS = ('l_params_maps', 'r_params_maps', 'l_center_map', 'r_center_map', 'l_prior_maps', 'r_prior_maps', 'segms')


class _Node(nn.Module):
    """Anonymous container so that parameters get the reference's dotted names."""


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, is_param: bool) -> None:
    parts = dotted.split('.')
    m = root
    for p in parts[:-1]:
        if not hasattr(m, p) or getattr(m, p) is None:
            m.add_module(p, _Node())
        m = getattr(m, p)
    if is_param:
        m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))
    else:
        m.register_buffer(parts[-1], tensor)


class LazyOutputs(dict):
    """``outputs`` dict whose seven map entries are converted from the engine's NHWC arena to the
    reference's fp32 NCHW tensors on first access (nothing on the hot path reads them)."""

    def __init__(self, engine, *a, **k):
        super().__init__(*a, **k)
        self._engine = engine
        for key in _MAP_KEYS:
            dict.__setitem__(self, key, None)

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if v is None and key in _MAP_KEYS and self._engine is not None:
            v = self._engine.map_nchw(key)
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class ACR(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self._spec = build_acr_spec(args().input_size)
        g = torch.Generator().manual_seed(0)
        for key, (shape, kind) in self._spec.params.items():
            if kind == 'bn_nbt':
                t = torch.zeros(shape, dtype=torch.long)
            elif kind == 'bn_var' or kind == 'bn_w':
                t = torch.ones(shape)
            elif kind in ('bn_mean', 'bn_b', 'conv_b', 'lin_b'):
                t = torch.zeros(shape)
            else:
                t = torch.randn(shape, generator=g) * 0.01
            _attach(self, key, t, is_param=kind not in ('bn_mean', 'bn_var', 'bn_nbt'))
        self._result_parser = ResultParser()
        self.outmap_size = args().centermap_size
        self._engines = {}
        self.debug_ref_conv = bool(kwargs.get('debug_ref_conv', False))

    # ------------------------------------------------------------------ engine
    def invalidate_engine(self):
        self._engines = {}

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate_engine()
        return r

    def _act_dtype(self):
        p = args().model_precision
        if p in ('bf16',):
            return torch.bfloat16
        if p in ('fp16',):
            return torch.float16
        raise ValueError("model_precision must be 'bf16' or 'fp16' on the B200 path (the reference's fp32 "
                         "mode has no tensor-core equivalent here)")

    def engine(self, batch: int, device) -> Engine:
        key = (batch, str(device), self._act_dtype(), self.debug_ref_conv)
        if key not in self._engines:
            self._engines[key] = Engine(self.state_dict(), batch, device, self._act_dtype(), args().input_size,
                                        debug_ref_conv=self.debug_ref_conv)
        return self._engines[key]

    # ----------------------------------------------------------------- forward
    def _image(self, meta_data):
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("acr.model.ACR runs on a CUDA device only: call .cuda() first (no CPU fallback)")
        img = meta_data['image']
        if img.dtype != torch.uint8:      # the reference accepts float 0..255 too
            img = img.round().clamp(0, 255).to(torch.uint8)
        return img.to(dev, non_blocking=True).contiguous(), dev

    @torch.no_grad()
    def forward(self, meta_data, **cfg):
        img, dev = self._image(meta_data)
        eng = self.engine(img.shape[0], dev)
        eng.run(img)
        outputs = LazyOutputs(eng if args().return_maps else None)
        bufs = self._result_parser.launch(eng.parse_inputs(), img.shape[0], meta_data, dev)
        outputs, meta_data = self._result_parser.collect(bufs, outputs, meta_data)
        outputs['meta_data'] = meta_data
        return outputs

    @torch.no_grad()
    def forward_dense(self, meta_data):
        """Sync-free variant for the fused pipeline: runs backbone + heads + parse and returns the
        engine and the worst-case (2B rows) parse buffers; row validity lives in ``bufs.counts``."""
        img, dev = self._image(meta_data)
        eng = self.engine(img.shape[0], dev)
        eng.run(img)
        bufs = self._result_parser.launch(eng.parse_inputs(), img.shape[0], meta_data, dev)
        return eng, bufs

    @torch.no_grad()
    def head_forward(self, x, gt_segm=None):
        raise NotImplementedError("head_forward on an externally supplied backbone feature is not exposed; the "
                                  "launch plan runs backbone and heads as one schedule (use forward())")
