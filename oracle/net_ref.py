"""ORACLE (test infrastructure, never on the product path).

CPU restatement (torch fp32 functional ops, driven directly by a state-dict with
the reference's key names) of the ACR network forward: HRNet-W32 trunk, SegmNet,
global heads and the part/attention branch.  Floating-point convolution is the one
place where the oracle keeps a torch fp32 reference instead of numpy.  Pinned
against the reference through tests/golden/net_golden.npz.

Reference functions restated (in /root/reference/acr/model.py):
  HigherResolutionNet.forward :831-865   BasicBlock :470-499   Bottleneck :501-539
  HighResolutionModule.forward :668-686  (+ fuse layer construction :620-663)
  SegmNet / Up / DoubleConv :374-463     ACR.head_forward :47-65
  ACR.global_forward :68-101             ACR.part_forward :116-166
  ACR.Hadamard_product :103-113          LocallyConnected2d.forward :559-569
  get_coord_maps :340-369                BHWC_to_BCHW acr/utils.py:226-231

``act_dtype`` (None | torch.bfloat16 | torch.float16) optionally rounds every
conv/fuse output to that storage type, which is where the B200 path rounds; the
arithmetic itself stays fp32.  With ``act_dtype=None`` this is the reference's
fp32 path.

``fold_round=True`` (needs ``act_dtype``) additionally restates the ROUNDING POINTS of
the 16-bit product path -- not its kernels: eval-mode BatchNorm is folded into the conv
weights and the folded weights are rounded to the storage type (bias stays fp32), the
normalised stem input is rounded (the im2col taps are stored 16-bit), the raw params /
cam head outputs are rounded (they are stored 16-bit before the final 1x1 conv), the
softmax weights of the attention pooling are rounded, and contact_layers[4|5] is
evaluated in its algebraically folded form (one 128->109 1x1 conv on the stored
[params106 | cam3] tensor with rounded weights + a per-image fp32 bias).  The result is
what an exact-arithmetic machine with the product path's storage types would produce:
the whole-network bf16 comparison against it isolates wiring errors from storage
round-off (tests/test_gpu_network.py).
"""
import torch
import torch.nn.functional as Fn

EPS = 1e-5
WIDTHS = (32, 64, 128, 256)


class _Net:
    def __init__(self, sd, act_dtype=None, fold_round=False):
        self.sd = sd
        self.q = act_dtype
        self.fold = bool(fold_round)
        assert not (self.fold and act_dtype is None), "fold_round needs a storage dtype"

    # -- helpers ------------------------------------------------------------
    def rnd(self, x):
        return x.to(self.q).float() if self.q is not None else x

    def conv(self, x, key, stride=1, pad=None):
        w = self.sd[key + ".weight"].float()
        b = self.sd.get(key + ".bias")
        if pad is None:
            pad = w.shape[-1] // 2
        return Fn.conv2d(x, w, None if b is None else b.float(), stride, pad)

    def bn(self, x, key):
        g, b = self.sd[key + ".weight"].float(), self.sd[key + ".bias"].float()
        m, v = self.sd[key + ".running_mean"].float(), self.sd[key + ".running_var"].float()
        s = g / torch.sqrt(v + EPS)
        return x * s.view(1, -1, 1, 1) + (b - m * s).view(1, -1, 1, 1)

    def convbn(self, x, ckey, bkey=None, stride=1):
        """conv (+bias) (+eval BatchNorm).  fold mode: y = conv(x, rnd(w * s)) + (beta - mean*s + cb*s), the
        weight/bias split of acr_b200_pack_conv."""
        if not self.fold:
            y = self.conv(x, ckey, stride)
            return self.bn(y, bkey) if bkey else y
        w = self.sd[ckey + ".weight"].float()
        cb = self.sd.get(ckey + ".bias")
        if bkey:
            g, b = self.sd[bkey + ".weight"].float(), self.sd[bkey + ".bias"].float()
            m, v = self.sd[bkey + ".running_mean"].float(), self.sd[bkey + ".running_var"].float()
            sc = g / torch.sqrt(v + EPS)
            sh = b - m * sc
        else:
            sc, sh = torch.ones(w.shape[0]), torch.zeros(w.shape[0])
        if cb is not None:
            sh = sh + cb.float() * sc
        return Fn.conv2d(x, self.rnd(w * sc.view(-1, 1, 1, 1)), sh, stride, w.shape[-1] // 2)

    def cbr(self, x, ckey, bkey, stride=1, relu=True):
        y = self.convbn(x, ckey, bkey, stride)
        return self.rnd(torch.relu(y) if relu else y)

    def basic(self, x, p):
        y = self.cbr(x, p + ".conv1", p + ".bn1")
        y = self.convbn(y, p + ".conv2", p + ".bn2") + x
        return self.rnd(torch.relu(y))

    def bottleneck(self, x, p, down):
        res = self.cbr(x, p + ".downsample.0", p + ".downsample.1", relu=False) if down else x
        y = self.cbr(x, p + ".conv1", p + ".bn1")
        y = self.cbr(y, p + ".conv2", p + ".bn2")
        y = self.convbn(y, p + ".conv3", p + ".bn3") + res
        return self.rnd(torch.relu(y))

    def hr_module(self, xs, prefix, multi):
        nb = len(xs)
        xs = list(xs)
        for b in range(nb):
            for k in range(4):
                xs[b] = self.basic(xs[b], f"{prefix}.branches.{b}.{k}")
        outs = []
        for i in range(nb if multi else 1):
            y = None
            for j in range(nb):
                if j == i:
                    t = xs[j]
                elif j > i:
                    p = f"{prefix}.fuse_layers.{i}.{j}"
                    t = self.cbr(xs[j], p + ".0", p + ".1", relu=False)
                    t = Fn.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
                else:
                    t = xs[j]
                    for k in range(i - j):
                        p = f"{prefix}.fuse_layers.{i}.{j}.{k}"
                        t = self.cbr(t, p + ".0", p + ".1", stride=2, relu=(k != i - j - 1))
                y = t if y is None else y + t
            outs.append(self.rnd(torch.relu(y)))
        return outs

    # -- network ------------------------------------------------------------
    def backbone(self, image_bhwc):
        x = image_bhwc.float().permute(0, 3, 1, 2)
        x = (x / 255.0) * 2.0 - 1.0
        if self.fold:
            x = self.rnd(x)          # the im2col taps of the tensor-core stem are stored in the 16-bit type
        x = self.cbr(x, "backbone.conv1", "backbone.bn1", stride=2)
        x = self.cbr(x, "backbone.conv2", "backbone.bn2", stride=2)
        for i in range(4):
            x = self.bottleneck(x, f"backbone.layer1.{i}", i == 0)
        xs = [self.cbr(x, "backbone.transition1.0.0", "backbone.transition1.0.1"),
              self.cbr(x, "backbone.transition1.1.0.0", "backbone.transition1.1.0.1", stride=2)]
        xs = self.hr_module(xs, "backbone.stage2.0", True)
        xs.append(self.cbr(xs[-1], "backbone.transition2.2.0.0", "backbone.transition2.2.0.1", stride=2))
        for m in range(4):
            xs = self.hr_module(xs, f"backbone.stage3.{m}", True)
        xs.append(self.cbr(xs[-1], "backbone.transition3.3.0.0", "backbone.transition3.3.0.1", stride=2))
        for m in range(3):
            xs = self.hr_module(xs, f"backbone.stage4.{m}", m != 2)
        return xs[0]

    def segm(self, x):
        up = self.rnd(Fn.interpolate(x, scale_factor=(2, 2), mode="bilinear", align_corners=True))
        p = "backbone.hand_segm.segm_head.upsampler.up1.conv.double_conv"
        y = self.cbr(up, p + ".0", p + ".1")
        y = self.cbr(y, p + ".3", p + ".4")
        p = "backbone.hand_segm.segm_head.segm_net.double_conv"
        y = self.cbr(y, p + ".0", p + ".1")
        return self.rnd(self.convbn(y, p + ".3"))

    def head_stack(self, x, p):
        y = self.cbr(x, p + ".0.0", p + ".0.1", stride=2)
        for k in range(2):
            y = self.basic(y, f"{p}.1.{k}.0")
        return self.convbn(y, p + ".2")

    def heads(self, x):
        B, _, H, W = x.shape
        seg = self.segm(x)
        lin = torch.arange(H, dtype=torch.float32) / (H - 1) * 2 - 1
        coords = torch.stack([lin.view(1, W).expand(H, W), lin.view(H, 1).expand(H, W)])  # ch0=x, ch1=y
        xc = torch.cat([x, self.rnd(coords)[None].expand(B, -1, -1, -1)], 1)
        out = {}
        raw = {}
        for s in ("l", "r"):
            prm = self.head_stack(xc, f"{s}_final_layers.1")
            out[f"{s}_center_map"] = self.head_stack(xc, f"{s}_final_layers.2")
            cam = self.head_stack(xc, f"{s}_final_layers.3")
            out[f"{s}_prior_maps"] = self.head_stack(xc, f"{s}_final_layers.4")
            cam = torch.cat([torch.pow(1.1, cam[:, :1]), cam[:, 1:]], 1)
            if self.fold:            # both heads are stored 16-bit (slices of the 128-channel input of the folded conv)
                cam, prm = self.rnd(cam), self.rnd(prm)
            raw[s] = torch.cat([cam, prm], 1)                      # (B,109,64,64)
        # ---- part branch
        att = seg[:, 1:, ::2, ::2]                                 # nearest 1/2, drop background
        contact = self.cbr(xc, "contact_layers.1.0", "contact_layers.1.1")
        shape_f = self.conv(contact, "cam_shape_layers.1.0")
        if self.fold:                # split softmax with weights rounded to the storage type; they still sum to one
            lg = att.reshape(B, 32, -1)
            e = self.rnd(torch.exp(lg - lg.max(-1, keepdim=True).values))
            a = e / e.sum(-1, keepdim=True)
        else:
            a = torch.softmax(att.reshape(B, 32, -1), -1)
        wc = torch.matmul(a, contact.reshape(B, 256, -1).transpose(1, 2)).transpose(1, 2)   # (B,256,32)
        ws = torch.matmul(a, shape_f.reshape(B, 64, -1).transpose(1, 2)).transpose(1, 2)    # (B,64,32)
        for s, sl, li, ci in (("l", slice(16, 32), 2, 4), ("r", slice(0, 16), 3, 5)):
            lw = self.sd[f"contact_layers.{li}.weight"].float()[0, :, :, :, 0, 0]           # (6,256,16)
            off = torch.einsum("bcj,ocj->boj", wc[:, :, sl], lw)                            # (B,6,16)
            off = off.transpose(1, 2).reshape(B, 96)                                         # joint-major
            sh = Fn.linear(ws[:, :, sl].reshape(B, -1), self.sd[f"cam_shape_layers.{li}.weight"].float(),
                           self.sd[f"cam_shape_layers.{li}.bias"].float())
            pare = torch.cat([off, sh], 1)[:, :, None, None].expand(-1, -1, 64, 64)
            if self.fold:
                # out = W[:, :109].pm + W[:, 109:112].pm[:3] + (b + W[:, 112:].pare),  pm = [cam3 | params106]
                Wf = self.sd[f"contact_layers.{ci}.weight"].float().reshape(109, 218)
                bf = self.sd[f"contact_layers.{ci}.bias"].float()
                w_prm, w_cam = self.rnd(Wf[:, 3:109]), self.rnd(Wf[:, 0:3] + Wf[:, 109:112])
                bias_img = bf[None] + torch.cat([off, sh], 1) @ Wf[:, 112:].t()            # (B,109) fp32
                out[f"{s}_params_maps"] = (torch.einsum("oc,bchw->bohw", w_prm, raw[s][:, 3:])
                                           + torch.einsum("oc,bchw->bohw", w_cam, raw[s][:, :3])
                                           + bias_img[:, :, None, None])
                continue
            inp = torch.cat([raw[s], raw[s][:, :3], pare], 1)                               # (B,218,64,64)
            out[f"{s}_params_maps"] = self.conv(inp, f"contact_layers.{ci}")
        out["segms"] = seg
        out["pooled"] = wc
        return out


def head_forward(sd, x, act_dtype=None, fold_round=False):
    """ACR.head_forward (acr/model.py:47-65): backbone feature (B,32,128,128) -> the 7 maps."""
    with torch.no_grad():
        return _Net(sd, act_dtype, fold_round).heads(x.float())


def net_forward(sd, image_bhwc, act_dtype=None, return_backbone=False, fold_round=False):
    """image uint8/float (B,512,512,3) RGB 0..255 -> the 7 maps of ACR.head_forward."""
    with torch.no_grad():
        n = _Net(sd, act_dtype, fold_round)
        x = n.backbone(image_bhwc)
        out = n.heads(x)
        if return_backbone:
            out["backbone"] = x
        return out
