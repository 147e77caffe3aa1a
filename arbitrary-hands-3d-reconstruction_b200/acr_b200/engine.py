"""Launch-plan builder: NetSpec + state-dict -> packed weight blob, activation arena, C op list.

Host-side counterpart of csrc/plan.cu.  Built once per (weights, batch size, dtype); running it is a
single C call (``acr_b200_plan_run``) that issues every kernel of the backbone + heads.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import lib as L
from .netspec import BN_EPS, NetSpec, Op, Tensor, build_acr_spec, conv_flops_per_image

ALIGN = 1024          # arena / blob alignment (TMA global address needs 16 B; generous for swizzle)
POOL_CHUNKS = 16
POOL_PART_FLOATS = 256 * 32 + 64


def _pow2(v):
    return v > 0 and (v & (v - 1)) == 0


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Blob:
    """Append-only byte blob with aligned sub-allocations (the packed weights)."""

    def __init__(self):
        self.parts: List[bytes] = []
        self.size = 0

    def add(self, arr: np.ndarray) -> int:
        off = _rup(self.size, 256)
        if off > self.size:
            self.parts.append(b"\0" * (off - self.size))
        b = np.ascontiguousarray(arr).tobytes()
        self.parts.append(b)
        self.size = off + len(b)
        return off

    def tobytes(self) -> bytes:
        return b"".join(self.parts)


class _Arena:
    """First-fit interval allocator over op indices (tensor liveness) -> byte offsets."""

    def __init__(self):
        self.free: List[Tuple[int, int]] = []   # (offset, size) sorted by offset
        self.top = 0

    def alloc(self, size: int) -> int:
        size = _rup(size, ALIGN)
        for i, (off, sz) in enumerate(self.free):
            if sz >= size:
                if sz == size:
                    self.free.pop(i)
                else:
                    self.free[i] = (off + size, sz - size)
                return off
        off = self.top
        self.top += size
        return off

    def release(self, off: int, size: int) -> None:
        size = _rup(size, ALIGN)
        self.free.append((off, size))
        self.free.sort()
        merged: List[Tuple[int, int]] = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        # give the tail back to the bump pointer
        if merged and merged[-1][0] + merged[-1][1] == self.top:
            self.top = merged[-1][0]
            merged.pop()
        self.free = merged


class Engine:
    """The backbone + heads of ACR as one precompiled CUDA launch plan.

    ``state_dict`` uses the reference's key names (checkpoint compatible).  ``act_dtype`` is
    torch.bfloat16 or torch.float16 (storage of activations/weights; accumulation is fp32) -- the
    product path on the tensor cores -- or torch.float32: the VALIDATION plan (fp32 storage, fp64
    accumulation on the CUDA cores, csrc/validate_f32.cu), which is what the reference's default
    ``model_precision='fp32'`` maps to and what pins the whole pipeline at 1e-4.
    ``debug_ref_conv`` swaps the tcgen05 conv for the CUDA-core reference kernel (tests only).
    ``head_only`` builds the plan of ``ACR.head_forward`` (/root/reference/acr/model.py:47-65): the ops
    after the trunk, fed by an external (B,32,H/4,W/4) feature through ``run_heads``.
    ``weights`` re-uses the packed weight blob of another engine of the same dtype / flags (the blob does
    not depend on the batch size).
    """

    def __init__(self, state_dict: Dict[str, torch.Tensor], batch: int, device, act_dtype=torch.bfloat16,
                 input_size: int = 512, debug_ref_conv: bool = False, reuse_memory: bool = True,
                 dry_run: bool = False, keep_extra=(), stem_on_tensor_cores: bool = True,
                 head_only: bool = False, weights: Optional[torch.Tensor] = None, widths=None):
        self.keep_extra = tuple(keep_extra)   # extra tensor names kept alive after the run (tests)
        self.dry_run = dry_run      # layout only (arena size, op list); used by CPU tests
        if not dry_run and not torch.cuda.is_available():
            raise L.AcrB200Error("Engine needs a CUDA device; there is no CPU fallback on the product path")
        self.lib = L.load()
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.batch = int(batch)
        self.act_dtype = act_dtype
        self.dt = {torch.bfloat16: L.DT_BF16, torch.float16: L.DT_F16, torch.float32: L.DT_F32}[act_dtype]
        self.f32 = act_dtype == torch.float32
        self.esz = 4 if self.f32 else 2
        self.npw = np.float32 if self.f32 else np.uint16      # numpy type of one packed weight
        self.stem_on_tensor_cores = stem_on_tensor_cores and not self.f32
        self.head_only = head_only
        from .netspec import WIDTHS
        # Folding the coarse fuse sums into the producing conv (ACR_B200_FOLD_FUSE=1) is an opt-in: measured on B200 it
        # removes 0.97 ms of fuse kernels and adds 0.7-1.9 ms to the convs (the extra terms are latency-bound loads in
        # the direct epilogue), i.e. no gain -- profiles/r2_bench_ab_fold_fuse.json.  Tensor-core plans only.
        self.spec: NetSpec = build_acr_spec(input_size, merge_stems=os.environ.get("ACR_B200_MERGE_STEMS", "1") != "0",
                                            widths=tuple(widths) if widths else WIDTHS,
                                            fold_fuse=(os.environ.get("ACR_B200_FOLD_FUSE", "0") != "0"
                                                       and act_dtype != torch.float32 and not debug_ref_conv))
        self.input_size = input_size
        self.flops_per_image = conv_flops_per_image(self.spec)
        self.debug_ref_conv = debug_ref_conv or self.f32
        self.run_count = 0          # bumped by every run(): lazily read outputs check it (acr/model.py)
        sd = {k: v.detach().float().cpu().numpy() for k, v in (state_dict or {}).items()
              if v.dtype.is_floating_point}
        self._build(sd, reuse_memory, weights)

    # ------------------------------------------------------------------ weights
    def _pack_conv(self, sd, blob: _Blob, wkey: str, bnkey: Optional[str], has_bias: bool, cin_pad: int,
                   cout_pad: int, pair: bool = False, s2x: bool = False) -> Tuple[int, int]:
        w = np.ascontiguousarray(sd[wkey + ".weight"], np.float32)
        cb = np.ascontiguousarray(sd[wkey + ".bias"], np.float32) if has_bias else None
        bn = [np.ascontiguousarray(sd[f"{bnkey}.{n}"], np.float32) for n in
              ("weight", "bias", "running_mean", "running_var")] if bnkey else [None] * 4
        if s2x:
            # stride-2 conv of a dense 32-channel tensor read as x-pairs (row = even pixel's channels | odd pixel's): tap
            # (ky,kx) reads the even half for kx = 1 and the odd half for kx = 0 (pair ox-1) / kx = 2 (pair ox)
            co, ci = w.shape[:2]
            w2 = np.zeros((co, 2 * ci, 3, 3), np.float32)
            for kx in range(3):
                off = 0 if kx == 1 else ci
                w2[:, off:off + ci, :, kx] = w[:, :, :, kx]
            w = w2
        if pair:
            # x-paired grid: channel index = dx*32 + c.  Output pixel x_out = 2j+dxo reads input pixel
            # x_in = 2(j+pt-1)+dxi through the original tap kx = x_in - x_out + 1 (zero block if outside 0..2)
            co, ci = w.shape[:2]
            w2 = np.zeros((2 * co, 2 * ci, 3, 3), np.float32)
            for dxo in range(2):
                for dxi in range(2):
                    for pt in range(3):
                        kx = 2 * (pt - 1) + dxi - dxo + 1
                        if 0 <= kx <= 2:
                            w2[dxo * co:(dxo + 1) * co, dxi * ci:(dxi + 1) * ci, :, pt] = w[:, :, :, kx]
            w = w2
            cb = None if cb is None else np.tile(cb, 2)
            bn = [None if b is None else np.tile(b, 2) for b in bn]
        cout, cin, k, _ = w.shape
        wp = np.zeros((cout_pad, k * k, cin_pad), self.npw)
        bias = np.zeros(cout_pad, np.float32)
        p = lambda a: None if a is None else a.ctypes.data
        L.check(self.lib.acr_b200_pack_conv(p(w), p(cb), p(bn[0]), p(bn[1]), p(bn[2]), p(bn[3]), BN_EPS, cout, cin, k,
                                            cout_pad, cin_pad, self.dt, wp.ctypes.data, bias.ctypes.data),
                "pack_conv " + wkey)
        return blob.add(wp), blob.add(bias)

    def _pack_raw(self, blob: _Blob, w: np.ndarray, cin_pad: int, cout_pad: int) -> int:
        cout, cin, k, _ = w.shape
        wp = np.zeros((cout_pad, k * k, cin_pad), self.npw)
        bias = np.zeros(cout_pad, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        L.check(self.lib.acr_b200_pack_conv(w.ctypes.data, None, None, None, None, None, BN_EPS, cout, cin, k,
                                            cout_pad, cin_pad, self.dt, wp.ctypes.data, bias.ctypes.data), "pack_conv")
        return blob.add(wp)

    # --------------------------------------------------------------------- plan
    def _build(self, sd, reuse_memory: bool, shared_weights: Optional[torch.Tensor] = None) -> None:
        spec, B = self.spec, self.batch
        blob = _Blob()
        ops = spec.ops
        if self.head_only:   # ACR.head_forward: everything after the trunk; feat32 (channels 0..31 of xcat) is external
            first = next(i for i, op in enumerate(ops) if op.kind == "coordcat")
            ops = ops[first:]

        # ---- memory geometry of every tensor
        geo: Dict[str, dict] = {}

        def geom(t: Tensor) -> dict:
            root = t.base or t
            if root.name not in geo:
                if root.dtype == "u8":
                    g = dict(stride=root.C, esz=1, dt=L.DT_U8)
                elif root.dtype == "f32":
                    g = dict(stride=_rup(root.C, 16), esz=4, dt=L.DT_F32)
                else:
                    g = dict(stride=_rup(root.C, 16), esz=self.esz, dt=self.dt)
                g["bytes"] = B * root.H * root.W * g["stride"] * g["esz"]
                g["offset"] = None
                geo[root.name] = g
            return geo[root.name]

        # flat fp32 scratch tensors of the part branch
        part = Tensor("pool_part", POOL_CHUNKS * POOL_PART_FLOATS, 1, 1, "f32")
        pooled = spec.tensors["pooled"]
        pooled.C, pooled.H, pooled.W = 256 * 32, 1, 1
        bias_img = {s: Tensor(f"{s}_bias_img", 112, 1, 1, "f32") for s in "lr"}
        pare = {s: Tensor(f"{s}_pare", 106, 1, 1, "f32") for s in "lr"}
        for t in [part] + list(bias_img.values()) + list(pare.values()):
            spec.tensors[t.name] = t

        # ---- expand the spec ops into launch records (python dicts first)
        recs: List[dict] = []
        for op in ops:
            if op.kind == "pool":
                recs.append(dict(kind=L.OP_POOL, out=part, ins=[op.ins[0], op.ins[1]]))
            elif op.kind == "parthead":
                if op.attrs["side"] == "l":   # one launch serves both sides
                    recs.append(dict(kind=L.OP_PARTHEAD, out=pooled, ins=[part],
                                     aux=[bias_img["l"], bias_img["r"], pare["l"], pare["r"]]))
                s = op.attrs["side"]   # folded contact_layers[4|5]: 1x1 conv 128 -> 109 with a per-image bias
                recs.append(dict(kind=L.OP_CONV_REF if self.debug_ref_conv else L.OP_CONV, out=op.out,
                                 ins=[op.ins[1]], aux=[bias_img[s]],
                                 attrs=dict(k=1, s=1, relu=False, residual=False, pow11=False, fold_side=s)))
            elif op.kind == "stem" and self.stem_on_tensor_cores and not self.debug_ref_conv \
                    and os.environ.get("ACR_B200_STEM_FUSED", "1") != "0" and _pow2(op.out.W // 16) and _pow2(op.out.H // 16):
                # conv1 + bn1 + relu as ONE tcgen05 GEMM whose im2col operand is built in shared memory (csrc/stem_tc.cu)
                recs.append(dict(kind=L.OP_STEM_TC, out=op.out, ins=[op.ins[0]], attrs=dict(stem=op.attrs)))
            elif op.kind == "stem" and self.stem_on_tensor_cores:
                # conv1 + bn1 + relu as im2col (27 normalised taps -> 32 channels) + a 1x1 tcgen05 conv
                cols = Tensor("stem_im2col", 32, op.out.H, op.out.W, "act")
                spec.tensors[cols.name] = cols
                recs.append(dict(kind=L.OP_IM2COL_STEM, out=cols, ins=[op.ins[0]]))
                recs.append(dict(kind=L.OP_CONV_REF if self.debug_ref_conv else L.OP_CONV, out=op.out, ins=[cols],
                                 attrs=dict(k=1, s=1, relu=True, residual=False, pow11=False, stem=op.attrs)))
            elif op.kind == "coordcat":
                recs.append(dict(kind=L.OP_COORD, out=op.out, ins=[op.ins[0]]))
            else:
                kind = {"stem": L.OP_STEM, "conv": L.OP_CONV_REF if self.debug_ref_conv else L.OP_CONV,
                        "fuse": L.OP_FUSE, "bilinear2x": L.OP_BILINEAR2X}[op.kind]
                recs.append(dict(kind=kind, out=op.out, ins=list(op.ins), attrs=op.attrs))
        self.recs = recs

        # ---- liveness (by root tensor) and arena offsets
        keep = {"segms", "l_center_map", "r_center_map", "l_prior_maps", "r_prior_maps", "l_params_maps",
                "r_params_maps", "pooled", "l_pare", "r_pare"}
        keep |= set(self.keep_extra)
        keep_roots = {(spec.tensors[k].base or spec.tensors[k]).name for k in keep}
        last_use: Dict[str, int] = {}
        for i, r in enumerate(recs):
            for t in r["ins"] + [r["out"]] + r.get("aux", []):
                last_use[(t.base or t).name] = i
        arena = _Arena()
        if self.head_only:      # the external feature is copied into the coord-concat buffer before the run
            xcat = spec.tensors["feat32"].base
            geom(xcat)["offset"] = arena.alloc(geom(xcat)["bytes"])
            keep_roots.add(xcat.name)
        for i, r in enumerate(recs):
            for t in [r["out"]] + r.get("aux", []):
                g = geom(t)
                if g["offset"] is None:
                    g["offset"] = arena.alloc(g["bytes"])
            for t in r["ins"]:
                g = geom(t)
                if t.dtype != "u8":
                    assert g["offset"] is not None, f"{t.name} read before written"
            if reuse_memory:
                for name, lu in last_use.items():
                    if lu == i and name not in keep_roots and name != "image" and geo[name]["offset"] is not None \
                            and not geo[name].get("freed"):
                        arena.release(geo[name]["offset"], geo[name]["bytes"])
                        geo[name]["freed"] = True
        self.arena_bytes = max(arena.top, ALIGN)
        for g in geo.values():
            if g["offset"] is not None:
                self.arena_bytes = max(self.arena_bytes, g["offset"] + _rup(g["bytes"], ALIGN))
        self.geo = geo
        self.n_ops = len(recs)
        if self.dry_run:
            return

        def ctensor(t: Tensor, pair: bool = False) -> L.Tensor:
            g = geom(t)
            ct = L.Tensor()
            ext = t.dtype == "u8"
            ct.offset = 0 if ext else g["offset"] + (t.c_off * g["esz"] if t.base is not None else 0)
            ct.C, ct.H, ct.W = t.C, t.H, t.W
            ct.pix_stride, ct.dtype, ct.external = g["stride"], g["dt"], int(ext)
            if pair:   # dense 32-channel NHWC seen as (H, W/2, 64): two x-adjacent pixels form one 128-byte row
                assert t.base is None and g["stride"] == t.C == 32 and t.W % 2 == 0 and not self.f32
                ct.C, ct.W, ct.pix_stride = 64, t.W // 2, 64
            return ct

        def s2x_able(r) -> bool:
            """3x3 stride-2 convs of a DENSE 32-channel tensor read it as x-pairs (H, W/2, 64): two boxes per tile with
            128-byte rows instead of nine 64-byte-row boxes of four parity views (csrc/conv_tc.cu MODE_S2X)."""
            a = r.get("attrs", {})
            if self.f32 or r["kind"] != L.OP_CONV or a.get("k") != 3 or a.get("s") != 2 or a.get("merged") or "stem" in a:
                return False
            x, y = r["ins"][0], r["out"]
            return (x.C == 32 and x.base is None and x.dtype == "act" and x.W % 32 == 0 and y.H % 16 == 0 and y.W % 16 == 0
                    and os.environ.get("ACR_B200_S2X", "1") != "0")

        def pairable(r) -> bool:
            """3x3 stride-1 32->32 convs on dense tensors run as 64->64 convs on the x-paired grid: same
            bytes in memory, but 128-byte operand rows (SWIZZLE_128B) instead of 64-byte ones."""
            a = r.get("attrs", {})
            if self.f32 or r["kind"] not in (L.OP_CONV, L.OP_CONV_REF) or "fold_side" in a or a.get("k") != 3 or a.get("s") != 1:
                return False
            ts = r["ins"] + [r["out"]]
            return all(t.C == 32 and t.base is None and t.dtype == "act" and t.W % 32 == 0 for t in ts)

        # ---- weights + C op records
        f32 = lambda k: np.ascontiguousarray(sd[k], np.float32)
        cops = (L.Op * len(recs))()
        for i, r in enumerate(recs):
            o = cops[i]
            o.kind = r["kind"]
            pair = pairable(r)
            s2x = s2x_able(r)
            o.out = ctensor(r["out"], pair)
            o.n_in = len(r["ins"])
            for j, t in enumerate(r["ins"]):
                o.in_[j] = ctensor(t, pair or (s2x and j == 0))
            for j, t in enumerate(r.get("aux", [])):
                o.aux[j] = ctensor(t)
            a = r.get("attrs", {})
            if r["kind"] in (L.OP_CONV, L.OP_CONV_REF):
                x = r["ins"][0]
                o.k, o.stride, o.relu, o.has_residual = a["k"], a["s"], int(a["relu"]), int(a["residual"])
                # K per tap: 33/34-channel inputs are padded to ONE 64-channel chunk (TMA zero-fills the
                # channels beyond the 48-wide buffer) instead of three 16-channel chunks: a TMA box costs
                # ~620 clk whatever its size, so fewer, fatter boxes win (tools/tma_bench.cu)
                # (same rule for the wider trunks: 48 -> 64, 96 -> 128 with zero-filled tails)
                o.cin_pad = _rup(x.C, 64) if x.C > 32 else _rup(x.C, 16)
                o.cout_pad = _rup(r["out"].C, 16)
                if a.get("extra"):
                    o.shift[0] |= 16    # ACR_CONV_EXTRA: in_[1..] are further terms, nearest-upsampled by 2**shift[j]
                    for q, (_, sh) in enumerate(a["extra"]):
                        o.shift[1 + q] = sh
                if "stem" in a:
                    # weights (64,3,3,3) OIHW -> (64, 32, 1, 1) with input channel (ky*3+kx)*3+ci; BN folded by pack_conv
                    w = f32(a["stem"]["w"] + ".weight")
                    w1 = np.zeros((64, 32, 1, 1), np.float32)
                    w1[:, :27, 0, 0] = w.transpose(0, 2, 3, 1).reshape(64, 27)
                    sd_stem = {"stem.weight": w1}
                    for nme in ("weight", "bias", "running_mean", "running_var"):
                        sd_stem[f"stembn.{nme}"] = f32(f"{a['stem']['bn']}.{nme}")
                    o.cin_pad, o.cout_pad = 32, 64
                    o.w_offset[0], o.w_offset[1] = self._pack_conv(sd_stem, blob, "stem", "stembn", False, 32, 64)
                elif "fold_side" in a:
                    # out = W[:, :109].pm + W[:, 109:112].pm[:3] + (b + W[:, 112:].pare),  pm = [cam3 | params106]
                    # (acr/model.py:158-164); input channel order of the 128-wide tensor: params at 0..105, cam at 112..114
                    W = f32(f"contact_layers.{4 if a['fold_side'] == 'l' else 5}.weight").reshape(109, 218)
                    weff = np.zeros((109, 128, 1, 1), np.float32)
                    weff[:, :106, 0, 0] = W[:, 3:109]
                    weff[:, 112:115, 0, 0] = W[:, 0:3] + W[:, 109:112]
                    o.cin_pad = 128
                    o.w_offset[0] = self._pack_raw(blob, weff, o.cin_pad, o.cout_pad)
                    o.shift[0] |= 1     # ACR_CONV_BIAS_PER_IMAGE (aux[0] = bias_img from the part head)
                elif s2x:
                    o.cin_pad = 64
                    o.w_offset[0], o.w_offset[1] = self._pack_conv(sd, blob, a["w"], a["bn"], a["bias"], 64, o.cout_pad, s2x=True)
                    o.shift[0] |= 8     # ACR_CONV_S2X
                elif pair:
                    o.cin_pad = o.cout_pad = 64
                    o.w_offset[0], o.w_offset[1] = self._pack_conv(sd, blob, a["w"], a["bn"], a["bias"], 64, 64, pair=True)
                    o.shift[0] |= 4     # ACR_CONV_XPAIR: side taps are 32x32 corners of the 64x64 block
                elif a.get("merged"):
                    # convs of identical geometry on the same input: weights / biases concatenated along cout
                    each = a["merged"]
                    wps, bs = [], []
                    for wk, bk in zip(a["w"], a["bn"]):
                        tmp = _Blob()
                        self._pack_conv(sd, tmp, wk, bk, a["bias"], o.cin_pad, each)
                        raw = tmp.tobytes()
                        nw = each * a["k"] * a["k"] * o.cin_pad * np.dtype(self.npw).itemsize
                        wps.append(np.frombuffer(raw[:nw], self.npw))
                        boff = _rup(nw, 256)
                        bs.append(np.frombuffer(raw[boff: boff + each * 4], np.float32))
                    o.w_offset[0] = blob.add(np.concatenate(wps))
                    o.w_offset[1] = blob.add(np.concatenate(bs))
                else:
                    o.w_offset[0], o.w_offset[1] = self._pack_conv(sd, blob, a["w"], a["bn"], a["bias"], o.cin_pad, o.cout_pad)
                    if a.get("pow11"):
                        o.shift[0] |= 2  # ACR_CONV_POW11_CH0
            elif r["kind"] == L.OP_STEM_TC:
                # weights (64,3,3,3) OIHW -> (64, 32, 1, 1) with input channel (ky*3+kx)*3+ci; BN folded by pack_conv
                w = f32(a["stem"]["w"] + ".weight")
                w1 = np.zeros((64, 32, 1, 1), np.float32)
                w1[:, :27, 0, 0] = w.transpose(0, 2, 3, 1).reshape(64, 27)
                sd_stem = {"stem.weight": w1}
                for nme in ("weight", "bias", "running_mean", "running_var"):
                    sd_stem[f"stembn.{nme}"] = f32(f"{a['stem']['bn']}.{nme}")
                o.w_offset[0], o.w_offset[1] = self._pack_conv(sd_stem, blob, "stem", "stembn", False, 32, 64)
            elif r["kind"] == L.OP_STEM:
                w = f32(a["w"] + ".weight")                                   # (64,3,3,3) OIHW
                g_, b_, m_, v_ = (f32(f"{a['bn']}.{n}") for n in ("weight", "bias", "running_mean", "running_var"))
                sc = g_ / np.sqrt(v_ + BN_EPS)
                wt = (w * sc[:, None, None, None]).transpose(2, 3, 1, 0).reshape(27, 64)  # [(ky,kx,ci)][co]
                o.w_offset[0] = blob.add(wt.astype(np.float32))
                o.w_offset[1] = blob.add((b_ - m_ * sc).astype(np.float32))
            elif r["kind"] == L.OP_FUSE:
                o.relu = int(a["relu"])
                for j, sh in enumerate(a["shifts"]):
                    o.shift[j] = sh
            elif r["kind"] == L.OP_PARTHEAD:
                keys = ["contact_layers.2.weight", "contact_layers.3.weight", "cam_shape_layers.1.0.weight",
                        "cam_shape_layers.1.0.bias", "cam_shape_layers.2.weight", "cam_shape_layers.3.weight",
                        "cam_shape_layers.2.bias", "cam_shape_layers.3.bias", "contact_layers.4.weight",
                        "contact_layers.5.weight", "contact_layers.4.bias", "contact_layers.5.bias"]
                for j, k in enumerate(keys):
                    o.w_offset[j] = blob.add(f32(k).reshape(-1))
        self.n_ops = len(recs)
        self._cops = cops
        if shared_weights is not None:     # same dtype / flags => byte-identical blob (packing is deterministic)
            if shared_weights.numel() != blob.size or shared_weights.device != self.device:
                raise L.AcrB200Error("Engine: the shared weight blob does not match this plan")
            self.weights = shared_weights
        else:
            self.weights = torch.frombuffer(bytearray(blob.tobytes()), dtype=torch.uint8).to(self.device)
        with torch.cuda.device(self.device):
            self.arena = torch.zeros(self.arena_bytes, dtype=torch.uint8, device=self.device)
            plan = C.c_void_p()
            L.check(self.lib.acr_b200_plan_create(cops, len(recs), B, self.arena.data_ptr(), self.arena_bytes,
                                                  self.weights.data_ptr(), self.weights.numel(), self.dt, C.byref(plan)),
                    "plan_create")
        self.plan = plan

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.acr_b200_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # ---------------------------------------------------------------------- run
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def run(self, image: torch.Tensor) -> None:
        """image: uint8 CUDA tensor (B, S, S, 3) RGB on this engine's device.  Asynchronous on that device's
        current stream.  Outputs (``view`` / ``map_nchw`` / ``parse_inputs``) alias the arena: they are valid
        until the next ``run`` on this engine."""
        if self.head_only:
            raise L.AcrB200Error("Engine.run: this is a heads-only plan, use run_heads(x)")
        if not (image.is_cuda and image.dtype == torch.uint8 and image.is_contiguous()):
            raise L.AcrB200Error("Engine.run expects a contiguous uint8 CUDA tensor (B,H,W,3)")
        if image.device != self.device:
            raise L.AcrB200Error(f"Engine.run: image lives on {image.device}, the plan on {self.device}")
        if tuple(image.shape) != (self.batch, self.input_size, self.input_size, 3):
            raise L.AcrB200Error(f"Engine.run: expected {(self.batch, self.input_size, self.input_size, 3)}, got {tuple(image.shape)}")
        with torch.cuda.device(self.device):
            self.run_count += 1
            L.check(self.lib.acr_b200_plan_run(self.plan, image.data_ptr(), self._stream()), "plan_run")

    def run_heads(self, x: torch.Tensor) -> None:
        """x: (B,32,S/4,S/4) backbone feature (any float dtype, NCHW like the reference's head_forward input).
        Copied (and rounded to the storage type) into channels 0..31 of the coord-concat buffer, then the
        heads-only plan runs."""
        if not self.head_only:
            raise L.AcrB200Error("Engine.run_heads needs an engine built with head_only=True")
        F = self.input_size // 4
        if tuple(x.shape) != (self.batch, 32, F, F) or not x.is_cuda or x.device != self.device:
            raise L.AcrB200Error(f"Engine.run_heads: expected a CUDA tensor {(self.batch, 32, F, F)} on {self.device}")
        with torch.cuda.device(self.device):
            self.run_count += 1
            self.view("feat32")[..., :32].copy_(x.permute(0, 2, 3, 1))
            L.check(self.lib.acr_b200_plan_run(self.plan, None, self._stream()), "plan_run")

    def profile(self, image: torch.Tensor) -> Dict[int, Tuple[float, int]]:
        """One serialised, event-bracketed pass: {op kind: (device ms, launches)}."""
        ms = np.zeros(16, np.float32)
        cnt = np.zeros(16, np.int32)
        with torch.cuda.device(self.device):
            L.check(self.lib.acr_b200_plan_profile(self.plan, image.data_ptr(), self._stream(), ms.ctypes.data,
                                                   cnt.ctypes.data), "plan_profile")
        return {k: (float(ms[k]), int(cnt[k])) for k in range(16) if cnt[k]}

    @property
    def num_launches(self) -> int:
        return int(self.lib.acr_b200_plan_num_launches(self.plan))

    # ------------------------------------------------------------------ outputs
    def view(self, name) -> torch.Tensor:
        """NHWC view of a tensor (name or netspec.Tensor) inside the arena, no copy: (B,H,W,stride - c_off)
        starting at the tensor's first channel (channel slices of a wider buffer start at their c_off)."""
        t = self.spec.tensors[name] if isinstance(name, str) else name
        g = self.geo[(t.base or t).name]
        if g["offset"] is None:
            raise L.AcrB200Error(f"tensor {t.name} is not part of this plan")
        tdt = {L.DT_F32: torch.float32, L.DT_BF16: torch.bfloat16, L.DT_F16: torch.float16}[g["dt"]]
        n = self.batch * t.H * t.W * g["stride"]
        off = g["offset"]
        flat = self.arena[off: off + n * g["esz"]].view(tdt)
        v = flat.view(self.batch, t.H, t.W, g["stride"])
        return v[..., t.c_off:] if t.base is not None and t.c_off else v

    def map_nchw(self, name) -> torch.Tensor:
        """fp32 NCHW copy of a tensor with its logical channel count (the reference's layout)."""
        t = self.spec.tensors[name] if isinstance(name, str) else name
        return self.view(t)[..., : t.C].permute(0, 3, 1, 2).float().contiguous()

    def parse_inputs(self) -> Dict[str, tuple]:
        out = {}
        for s in "lr":
            out[f"{s}_center"] = (self.view(f"{s}_center_map"), 16)
            out[f"{s}_params"] = (self.view(f"{s}_params_maps"), 112)
            out[f"{s}_prior"] = (self.view(f"{s}_prior_maps"), 112)
        return out
