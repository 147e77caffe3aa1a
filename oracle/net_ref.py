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
"""
import torch
import torch.nn.functional as Fn

EPS = 1e-5
WIDTHS = (32, 64, 128, 256)


class _Net:
    def __init__(self, sd, act_dtype=None):
        self.sd = sd
        self.q = act_dtype

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

    def cbr(self, x, ckey, bkey, stride=1, relu=True):
        y = self.bn(self.conv(x, ckey, stride), bkey)
        return self.rnd(torch.relu(y) if relu else y)

    def basic(self, x, p):
        y = self.cbr(x, p + ".conv1", p + ".bn1")
        y = self.bn(self.conv(y, p + ".conv2"), p + ".bn2") + x
        return self.rnd(torch.relu(y))

    def bottleneck(self, x, p, down):
        res = self.cbr(x, p + ".downsample.0", p + ".downsample.1", relu=False) if down else x
        y = self.cbr(x, p + ".conv1", p + ".bn1")
        y = self.cbr(y, p + ".conv2", p + ".bn2")
        y = self.bn(self.conv(y, p + ".conv3"), p + ".bn3") + res
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
        return self.rnd(self.conv(y, p + ".3"))

    def head_stack(self, x, p):
        y = self.cbr(x, p + ".0.0", p + ".0.1", stride=2)
        for k in range(2):
            y = self.basic(y, f"{p}.1.{k}.0")
        return self.conv(y, p + ".2")

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
            raw[s] = torch.cat([cam, prm], 1)                      # (B,109,64,64)
        # ---- part branch
        att = seg[:, 1:, ::2, ::2]                                 # nearest 1/2, drop background
        contact = self.cbr(xc, "contact_layers.1.0", "contact_layers.1.1")
        shape_f = self.conv(contact, "cam_shape_layers.1.0")
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
            inp = torch.cat([raw[s], raw[s][:, :3], pare], 1)                               # (B,218,64,64)
            out[f"{s}_params_maps"] = self.conv(inp, f"contact_layers.{ci}")
        out["segms"] = seg
        out["pooled"] = wc
        return out


def net_forward(sd, image_bhwc, act_dtype=None, return_backbone=False):
    """image uint8/float (B,512,512,3) RGB 0..255 -> the 7 maps of ACR.head_forward."""
    with torch.no_grad():
        n = _Net(sd, act_dtype)
        x = n.backbone(image_bhwc)
        out = n.heads(x)
        if return_backbone:
            out["backbone"] = x
        return out
