"""Oracle restatement of the reference 3D UNet / ResUNet as pure functions over a state_dict.

Follows (reference file:line):
  UNet.forward                     model/dim3/unet.py:50-64
  inconv / down_block / up_block   model/dim3/unet_utils.py:7-21, 24-46, 48-75
  ConvNormAct / SingleConv / BasicBlock   model/dim3/conv_layers.py:16-53, 56-68, 71-94
  InstanceNorm3d(eps=1e-4), affine=False  conv_layers.py:40,42
TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's baseline legs may import this.
"""
import torch
import torch.nn.functional as F

EPS = 1e-4


def _k3(k):
    return [k] * 3 if isinstance(k, int) else list(k)


def _pad(k):
    return [i // 2 for i in k]


_MARGIN_PROBE = None      # list collecting min |xhat| at every ReLU input while relu_margin_probe() is active


class relu_margin_probe:
    """Context manager: records, for every ReLU input of a forward pass, the smallest |normalised value|.  A test
    uses it to PROVE a case is mask-flip-free (every pre-activation is further from 0 than the implementation's
    rounding error), so gradients can be compared at 1e-3 with no noise-floor allowance."""

    def __enter__(self):
        global _MARGIN_PROBE
        _MARGIN_PROBE = []
        return _MARGIN_PROBE

    def __exit__(self, *a):
        global _MARGIN_PROBE
        _MARGIN_PROBE = None


def _in_relu(x):
    h = F.instance_norm(x, eps=EPS)
    if _MARGIN_PROBE is not None:
        _MARGIN_PROBE.append(h.detach().abs().min().item())
    return F.relu(h)


def single_conv(sd, pre, x, k):
    """act(norm(conv(x))): conv_layers.py:50-51 via SingleConv (preact=False)."""
    return _in_relu(F.conv3d(x, sd[pre + "conv.conv.weight"], padding=_pad(k)))


def basic_block(sd, pre, x, k):
    """conv_layers.py:86-94: two pre-activation convs + (conv) shortcut."""
    out = F.conv3d(_in_relu(x), sd[pre + "conv1.conv.weight"], padding=_pad(k))
    out = F.conv3d(_in_relu(out), sd[pre + "conv2.conv.weight"], padding=_pad(k))
    key = pre + "shortcut.conv.weight"
    res = F.conv3d(_in_relu(x), sd[key], padding=_pad(k)) if key in sd else x
    return out + res


_BLOCK = {"SingleConv": single_conv, "BasicBlock": basic_block}


def unet_forward(sd, x, scale, kernel_size, block="BasicBlock"):
    """x: [B,in_ch,D,H,W] -> logits [B,classes,D,H,W]."""
    blk = _BLOCK[block]
    ks = [_k3(k) for k in kernel_size]
    sc = [_k3(s) for s in scale]
    # inc (unet_utils.py:17-21)
    x1 = F.conv3d(x, sd["inc.conv1.weight"], padding=_pad(ks[0]))
    x1 = blk(sd, "inc.conv2.", x1, ks[0])
    feats = [x1]
    cur = x1
    # down1..4 (unet_utils.py:35-46): MaxPool3d(scale) then 2 blocks with kernel_size[i+1]
    for i in range(4):
        cur = F.max_pool3d(cur, sc[i])
        cur = blk(sd, "down%d.conv.1." % (i + 1), cur, ks[i + 1])
        cur = blk(sd, "down%d.conv.2." % (i + 1), cur, ks[i + 1])
        feats.append(cur)
    # up1..4 (unet_utils.py:68-75; kernel_size[3],[2],[1],[0], unet.py:42-45)
    for j in range(4):
        skip = feats[3 - j]
        up = F.interpolate(cur, size=skip.shape[2:], mode="trilinear", align_corners=True)
        cur = torch.cat([skip, up], dim=1)
        cur = blk(sd, "up%d.conv.0." % (j + 1), cur, ks[3 - j])
        cur = blk(sd, "up%d.conv.1." % (j + 1), cur, ks[3 - j])
    return F.conv3d(cur, sd["outc.weight"], sd["outc.bias"])


def unet_param_shapes(in_ch, base, classes, kernel_size, block="BasicBlock"):
    """state_dict key -> shape, in the reference's registration order (for layout/host tests)."""
    ks = [_k3(k) for k in kernel_size]
    out = {}

    def add_block(pre, ci, co, k):
        if block == "SingleConv":
            out[pre + "conv.conv.weight"] = (co, ci, *k)
        else:
            out[pre + "conv1.conv.weight"] = (co, ci, *k)
            out[pre + "conv2.conv.weight"] = (co, co, *k)
            if ci != co:
                out[pre + "shortcut.conv.weight"] = (co, ci, *k)

    ch = [base, 2 * base, 4 * base, 8 * base, 10 * base]
    out["inc.conv1.weight"] = (base, in_ch, *ks[0])
    add_block("inc.conv2.", base, base, ks[0])
    for i in range(4):
        add_block("down%d.conv.1." % (i + 1), ch[i], ch[i + 1], ks[i + 1])
        add_block("down%d.conv.2." % (i + 1), ch[i + 1], ch[i + 1], ks[i + 1])
    for j in range(4):
        ci, co = ch[4 - j], ch[3 - j]
        add_block("up%d.conv.0." % (j + 1), ci + co, co, ks[3 - j])
        add_block("up%d.conv.1." % (j + 1), co, co, ks[3 - j])
    out["outc.weight"] = (classes, base, 1, 1, 1)
    out["outc.bias"] = (classes,)
    return out


def make_state_dict(shapes, seed=0, dtype=torch.float32):
    """Deterministic synthetic weights (kaiming-like scale) keyed by name — independent of module
    construction order and of torch's default-init RNG stream."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shp in shapes.items():
        fan_in = 1
        for s in shp[1:]:
            fan_in *= s
        bound = (1.0 / max(fan_in, 1)) ** 0.5
        sd[name] = ((torch.rand(*shp, generator=g, dtype=torch.float64) * 2 - 1) * bound * 1.7320508).to(dtype)
    return sd
