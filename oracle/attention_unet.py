"""TEST INFRASTRUCTURE — oracle restatement of the reference Attention-UNet as pure functions over a state_dict.

Follows (reference file:line):
  AttentionUNet.forward        model/dim3/attention_unet.py:31-46
  AttentionBlock.forward       model/dim3/attention_unet_utils.py:26-37  (InstanceNorm3d default eps 1e-5, :13,17,21)
  attention_up_block.forward   model/dim3/attention_unet_utils.py:55-64  (`conv_ch` is constructed but never called)
Encoder / blocks are oracle/unet3d.py's.  Pinned by oracle/make_golden_attention_unet.py."""
import torch
import torch.nn.functional as F

from .unet3d import _BLOCK, _k3, _pad, unet_param_shapes

GATE_EPS = 1e-5


def attention_block(sd, pre, g, x):
    """g: upsampled low-resolution feature, x: encoder skip (attention_unet_utils.py:26-37)."""
    g1 = F.instance_norm(F.conv3d(g, sd[pre + "W_g.0.weight"]), eps=GATE_EPS)
    x1 = F.instance_norm(F.conv3d(x, sd[pre + "W_x.0.weight"]), eps=GATE_EPS)
    psi = F.relu(g1 + x1)
    psi = torch.sigmoid(F.instance_norm(F.conv3d(psi, sd[pre + "psi.0.weight"]), eps=GATE_EPS))
    return x * psi


def attention_unet_forward(sd, x, scale, kernel_size, block="BasicBlock"):
    blk = _BLOCK[block]
    ks = [_k3(k) for k in kernel_size]
    sc = [_k3(s) for s in scale]
    x1 = F.conv3d(x, sd["inc.conv1.weight"], padding=_pad(ks[0]))
    x1 = blk(sd, "inc.conv2.", x1, ks[0])
    feats = [x1]
    cur = x1
    for i in range(4):
        cur = F.max_pool3d(cur, sc[i])
        cur = blk(sd, "down%d.conv.1." % (i + 1), cur, ks[i + 1])
        cur = blk(sd, "down%d.conv.2." % (i + 1), cur, ks[i + 1])
        feats.append(cur)
    for j in range(4):
        skip = feats[3 - j]
        up = F.interpolate(cur, size=skip.shape[2:], mode="trilinear", align_corners=True)
        gated = attention_block(sd, "up%d.attn." % (j + 1), up, skip)
        cur = torch.cat([gated, up], dim=1)
        cur = blk(sd, "up%d.conv.0." % (j + 1), cur, ks[3 - j])
        cur = blk(sd, "up%d.conv.1." % (j + 1), cur, ks[3 - j])
    return F.conv3d(cur, sd["outc.weight"], sd["outc.bias"])


def attention_unet_param_shapes(in_ch, base, classes, kernel_size, block="BasicBlock"):
    """state_dict key -> shape in the reference's registration order: the UNet's, with each up level prefixed by
    conv_ch.{weight,bias} and attn.{W_g,W_x,psi}.0.weight (attention_unet_utils.py:43-47)."""
    base_shapes = unet_param_shapes(in_ch, base, classes, kernel_size, block)
    ch = [base, 2 * base, 4 * base, 8 * base, 10 * base]
    out = {}
    seen = set()
    for k, v in base_shapes.items():
        if k.startswith("up"):
            j = int(k[2]) - 1
            if j not in seen:
                seen.add(j)
                ci, co = ch[4 - j], ch[3 - j]
                pre = "up%d." % (j + 1)
                out[pre + "conv_ch.weight"] = (co, ci, 1, 1, 1)
                out[pre + "conv_ch.bias"] = (co,)
                out[pre + "attn.W_g.0.weight"] = (co // 2, ci, 1, 1, 1)
                out[pre + "attn.W_x.0.weight"] = (co // 2, co, 1, 1, 1)
                out[pre + "attn.psi.0.weight"] = (1, co // 2, 1, 1, 1)
        out[k] = v
    return out
