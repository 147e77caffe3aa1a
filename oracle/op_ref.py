"""ORACLE (test infrastructure, never on the product path).

Per-op CPU restatement (torch fp32) of every kind of launch in the network plan, used by the
teacher-forced sweep (tests/test_gpu_teacher_forced.py): each op of the CUDA plan is re-computed here
from the plan's OWN stored inputs and the unrounded fp32 parameters of the state dict, so every one of
the ~375 launches is pinned individually (no error accumulates from op to op and no chaotic
amplification through the seeded random network).  The whole-network restatement in net_ref.py is what
is pinned against the reference goldens; the functions below are the same arithmetic cut at op
boundaries, each citing the reference lines it follows (/root/reference/acr/model.py).
"""
import torch
import torch.nn.functional as Fn

EPS = 1e-5


def conv_bn_act(x, sd, wkey, bnkey=None, stride=1, relu=False, residual=None, pow11=False):
    """nn.Conv2d (+bias) -> eval BatchNorm2d -> (+residual) -> (ReLU); BasicBlock :470-499, Bottleneck
    :501-539, fuse / transition convs :620-663, 703-736, head stacks :288-313, SegmNet :374-463.
    `pow11`: channel 0 -> 1.1**x after the conv (cam scale, :95-96; the cam head has no BN / act)."""
    w = sd[wkey + ".weight"].float()
    b = sd.get(wkey + ".bias")
    y = Fn.conv2d(x, w, None if b is None else b.float(), stride, w.shape[-1] // 2)
    if bnkey:
        g, be = sd[bnkey + ".weight"].float(), sd[bnkey + ".bias"].float()
        m, v = sd[bnkey + ".running_mean"].float(), sd[bnkey + ".running_var"].float()
        s = g / torch.sqrt(v + EPS)
        y = y * s.view(1, -1, 1, 1) + (be - m * s).view(1, -1, 1, 1)
    if pow11:
        y = torch.cat([torch.pow(1.1, y[:, :1]), y[:, 1:]], 1)
    if residual is not None:
        y = y + residual
    return torch.relu(y) if relu else y


def stem(image_bhwc, sd):
    """HigherResolutionNet.forward :832-835: x/255*2-1, conv1 3x3 s2 + bn1 + relu."""
    x = image_bhwc.float().permute(0, 3, 1, 2)
    x = (x / 255.0) * 2.0 - 1.0
    return conv_bn_act(x, sd, "backbone.conv1", "backbone.bn1", stride=2, relu=True)


def im2col_stem(image_bhwc):
    """The 27 normalised taps of every stride-2 output pixel, channel (ky*3+kx)*3+ci, zero in the conv
    padding (the tensor-core form of the stem: an exact re-indexing of :832-835's input)."""
    x = image_bhwc.float().permute(0, 3, 1, 2)
    x = (x / 255.0) * 2.0 - 1.0
    cols = Fn.unfold(x, 3, padding=1, stride=2)                         # (B, ci*9 + ky*3+kx, Ho*Wo)
    B, _, Hh, Ww = x.shape
    cols = cols.view(B, 3, 9, Hh // 2, Ww // 2).permute(0, 2, 1, 3, 4)  # (B, tap, ci, Ho, Wo)
    return cols.reshape(B, 27, Hh // 2, Ww // 2)


def stem_from_cols(cols27, sd):
    """conv1 + bn1 + relu applied to the im2col tensor (1x1 contraction over the 27 taps)."""
    w = sd["backbone.conv1.weight"].float()                              # (64, 3, 3, 3) OIHW
    w1 = w.permute(0, 2, 3, 1).reshape(64, 27, 1, 1)                     # channel (ky*3+kx)*3+ci
    y = Fn.conv2d(cols27, w1)
    g, be = sd["backbone.bn1.weight"].float(), sd["backbone.bn1.bias"].float()
    m, v = sd["backbone.bn1.running_mean"].float(), sd["backbone.bn1.running_var"].float()
    s = g / torch.sqrt(v + EPS)
    return torch.relu(y * s.view(1, -1, 1, 1) + (be - m * s).view(1, -1, 1, 1))


def fuse(terms, shifts, relu=True):
    """HighResolutionModule.forward :677-684: sum in the order j = 0..nb-1, nearest upsample by 2**shift."""
    y = None
    for t, sh in zip(terms, shifts):
        if sh:
            t = Fn.interpolate(t, scale_factor=2 ** sh, mode="nearest")
        y = t if y is None else y + t
    return torch.relu(y) if relu else y


def bilinear2x(x):
    """Up.forward :432."""
    return Fn.interpolate(x, scale_factor=(2, 2), mode="bilinear", align_corners=True)


def coord(H, W):
    """get_coord_maps :340-369 (+ cat :52): ch0 = x in [-1,1] along W, ch1 = y along H."""
    lx = torch.arange(W, dtype=torch.float32) / (W - 1) * 2 - 1
    ly = torch.arange(H, dtype=torch.float32) / (H - 1) * 2 - 1
    return torch.stack([lx.view(1, W).expand(H, W), ly.view(H, 1).expand(H, W)])


def attention_pool(contact, segm):
    """part_forward :126-136 + Hadamard_product :103-113: nearest 1/2 of the logits, drop the background,
    softmax over HW, (B,32,HW) @ (B,HW,256) -> (B,256,32)."""
    B = contact.shape[0]
    att = segm[:, 1:33, ::2, ::2]
    a = torch.softmax(att.reshape(B, 32, -1), -1)
    return torch.matmul(a, contact.reshape(B, 256, -1).transpose(1, 2)).transpose(1, 2)


def part_offsets(pooled, sd, side):
    """part_forward :139-156 after the pooling, for one hand: LocallyConnected2d (:559-569) on its 16 parts
    (left = parts 16..31, right = 0..15; output joint-major), the 256->64 1x1 conv of cam_shape_layers[1]
    applied to the pooled feature (conv-then-pool == pool-then-conv: softmax weights sum to 1) and
    Linear 1024->10.  -> (B,106) = [96 contact offsets | 10 shape offsets]."""
    B = pooled.shape[0]
    sl, li = (slice(16, 32), 2) if side == "l" else (slice(0, 16), 3)
    lw = sd[f"contact_layers.{li}.weight"].float()[0, :, :, :, 0, 0]            # (6,256,16)
    off = torch.einsum("bcj,ocj->boj", pooled[:, :, sl], lw).transpose(1, 2).reshape(B, 96)
    ws = torch.einsum("oc,bcj->boj", sd["cam_shape_layers.1.0.weight"].float()[:, :, 0, 0], pooled) \
        + sd["cam_shape_layers.1.0.bias"].float().view(1, -1, 1)                  # (B,64,32)
    sh = Fn.linear(ws[:, :, sl].reshape(B, -1), sd[f"cam_shape_layers.{li}.weight"].float(),
                   sd[f"cam_shape_layers.{li}.bias"].float())
    return torch.cat([off, sh], 1)


def final_params(params106, cam3, pare, sd, side):
    """part_forward :158-164: cat[cam3 | params106 | cam3 | pare(106, spatially constant)] = 218 channels ->
    contact_layers[4|5] 1x1 conv -> (B,109,64,64).  `cam3` is after the 1.1** of :95-96."""
    ci = 4 if side == "l" else 5
    B, _, H, W = params106.shape
    inp = torch.cat([cam3, params106, cam3, pare[:, :, None, None].expand(-1, -1, H, W)], 1)
    return Fn.conv2d(inp, sd[f"contact_layers.{ci}.weight"].float(), sd[f"contact_layers.{ci}.bias"].float())
