"""Oracle restatement of the reference MedFormer-3D as pure functions over a state_dict (NCDHW, plain torch).

Follows (reference file:line):
  MedFormer.forward                      model/dim3/medformer.py:68-101
  inconv / down_block / up_block         model/dim3/medformer_utils.py:271-284, 288-327, 329-383
  PatchMerging                           medformer_utils.py:140-177
  SemanticMapGeneration / Fusion         medformer_utils.py:204-228, 231-268
  BidirectionAttention(+Block)           medformer_utils.py:11-97, 102-138
  MBConv / SEBlock / DepthwiseSeparable  model/dim3/conv_layers.py:197-238, 159-174, 126-157
  TransformerBlock / Attention / Mlp     model/dim3/trans_layers.py:16-118
Pinned against the unmodified reference by oracle/make_golden_medformer.py (tests/golden/medformer_*.pt).
Configuration covered: norm 'in', act 'relu', conv_block 'BasicBlock', proj_type 'depthwise', dropout 0.
TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's baseline legs may import this.
"""
import torch
import torch.nn.functional as F

from .medformer_ops import bidirection_attention_core, depthwise_conv3d
from .unet3d import basic_block

EPS_BLOCK = 1e-4     # ConvNormAct: norm(in_ch, eps=1e-4), conv_layers.py:40
EPS_PLAIN = 1e-5     # norm(dim) with the default eps: medformer_utils.py:107-108,158


def _k3(k):
    return [k] * 3 if isinstance(k, int) else list(k)


def _in(x, eps):
    return F.instance_norm(x, eps=eps)


def dwsep(sd, pre, x):
    """DepthwiseSeparableConv.forward, conv_layers.py:153-157."""
    return F.conv3d(depthwise_conv3d(x, sd[pre + "depthwise.weight"]), sd[pre + "pointwise.weight"])


def se_block(sd, pre, x):
    """conv_layers.py:169-174."""
    s = x.mean(dim=(2, 3, 4), keepdim=True)
    s = F.relu(F.conv3d(s, sd[pre + "excitation.0.weight"], sd[pre + "excitation.0.bias"]))
    s = torch.sigmoid(F.conv3d(s, sd[pre + "excitation.2.weight"], sd[pre + "excitation.2.bias"]))
    return x * s


def mbconv(sd, pre, x):
    """conv_layers.py:221-238 with identity shortcut and DropPath(p=0)."""
    e = F.conv3d(F.relu(_in(x, EPS_BLOCK)), sd[pre + "expand_proj.conv.weight"])
    d = depthwise_conv3d(F.relu(_in(e, EPS_BLOCK)), sd[pre + "depthwise.conv.weight"])
    d = se_block(sd, pre + "se.", d)
    return F.conv3d(_in(d, EPS_BLOCK), sd[pre + "pointwise.conv.weight"]) + x


def biattn_block(sd, pre, x, smap, heads):
    """BidirectionAttentionBlock.forward, medformer_utils.py:124-138."""
    feat, mapp = _in(x, EPS_PLAIN), _in(smap, EPS_PLAIN)
    fq, fv = dwsep(sd, pre + "attn.feat_qv.", feat).chunk(2, dim=1)
    mq, mv = F.conv3d(mapp, sd[pre + "attn.map_qv.weight"]).chunk(2, dim=1)
    fo, mo = bidirection_attention_core(fq, fv, mq, mv, heads)
    out = dwsep(sd, pre + "attn.feat_out.", fo)
    if pre + "attn.map_out.weight" in sd:
        mo = F.conv3d(mo, sd[pre + "attn.map_out.weight"])
    sc = pre + "shortcut.conv.weight"
    out = out + (F.conv3d(F.relu(_in(x, EPS_BLOCK)), sd[sc]) if sc in sd else x)
    out = mbconv(sd, pre + "feedforward.", out)
    return out, mo + smap


def basic_layer(sd, pre, x, smap, num_blocks, heads):
    for i in range(num_blocks):
        x, smap = biattn_block(sd, "%sblocks.%d." % (pre, i), x, smap, heads)
    return x, smap


def patch_merging(sd, pre, x, scale):
    """medformer_utils.py:160-177."""
    parts = []
    for i in range(scale[0]):
        for j in range(scale[1]):
            for k in range(scale[2]):
                parts.append(x[:, :, i::scale[0], j::scale[1], k::scale[2]])
    return dwsep(sd, pre + "reduction.", _in(torch.cat(parts, 1), EPS_PLAIN))


def map_generation(sd, pre, x, map_size):
    """medformer_utils.py:216-228."""
    B = x.shape[0]
    feat = F.conv3d(x, sd[pre + "base_proj.weight"], padding=1)
    wm = F.conv3d(x, sd[pre + "semantic_proj.weight"], padding=1)
    wm = F.softmax(wm.reshape(B, wm.shape[1], -1), dim=2)
    smap = torch.einsum("bij,bkj->bik", feat.reshape(B, feat.shape[1], -1), wm)
    return smap.reshape(B, feat.shape[1], *map_size)


def token_attention(sd, pre, x, heads):
    """Attention.forward, trans_layers.py:78-100 ('(heads dim_head)' channel order both ways)."""
    B, L, _ = x.shape
    qkv = F.linear(x, sd[pre + "to_qkv.weight"])
    q, k, v = (t.reshape(B, L, heads, -1).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    attn = F.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * q.shape[-1] ** -0.5, dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(B, L, -1)
    return F.linear(out, sd[pre + "to_out.weight"], sd[pre + "to_out.bias"])


def transformer_block(sd, pre, x, depth, heads):
    """trans_layers.py:103-118 with PreNorm :35-41 and Mlp :16-33."""
    for i in range(depth):
        p = "%slayers.%d." % (pre, i)
        xn = F.layer_norm(x, x.shape[-1:], sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])
        x = token_attention(sd, p + "0.fn.", xn, heads) + x
        xn = F.layer_norm(x, x.shape[-1:], sd[p + "1.norm.weight"], sd[p + "1.norm.bias"])
        h = F.gelu(F.linear(xn, sd[p + "1.fn.fc1.weight"], sd[p + "1.fn.fc1.bias"]))
        x = F.linear(h, sd[p + "1.fn.fc2.weight"], sd[p + "1.fn.fc2.bias"]) + x
    return x


def map_fusion(sd, maps, depth, heads):
    """SemanticMapFusion.forward, medformer_utils.py:250-268."""
    B = maps[0].shape[0]
    sp = maps[0].shape[2:]
    toks = [F.conv3d(m, sd["map_fusion.in_proj.%d.weight" % i]).flatten(2).permute(0, 2, 1) for i, m in enumerate(maps)]
    x = transformer_block(sd, "map_fusion.fusion.", torch.cat(toks, dim=1), depth, heads)
    outs = []
    for i, t in enumerate(x.chunk(len(maps), dim=1)):
        outs.append(F.conv3d(t.permute(0, 2, 1).reshape(B, -1, *sp), sd["map_fusion.out_proj.%d.weight" % i]))
    return outs


def medformer_forward(sd, x, cfg):
    """cfg: dict(map_size, conv_num, trans_num, num_heads, fusion_depth, fusion_heads, kernel_size, scale, aux_loss).
    Returns [logits, aux] (aux_loss) or logits — medformer.py:68-101."""
    ks = [_k3(k) for k in cfg["kernel_size"]]
    sc = [_k3(s) for s in cfg["scale"]]
    cn, tn, nh, ms = cfg["conv_num"], cfg["trans_num"], cfg["num_heads"], list(cfg["map_size"])

    def down(name, x, lvl):
        out = patch_merging(sd, name + ".patch_merging.", x, sc[lvl])
        for i in range(cn[lvl]):
            out = basic_block(sd, "%s.conv_blocks.%d." % (name, i), out, ks[lvl + 1])
        smap = map_generation(sd, name + ".map_gen.", out, ms) if (name + ".map_gen.base_proj.weight") in sd else None
        return basic_layer(sd, name + ".trans_blocks.", out, smap, tn[lvl], nh[lvl])

    def up(name, x1, x2, map1, map2, lvl, k):
        x1 = F.interpolate(x1, size=x2.shape[-3:], mode="trilinear", align_corners=True)
        feat = torch.cat([x1, x2], dim=1)
        if (name + ".map_reduction.weight") in sd and map2 is not None:
            smap = F.conv3d(torch.cat([map1, map2], dim=1), sd[name + ".map_reduction.weight"])
        else:
            smap = map1
        out, smap = basic_layer(sd, name + ".trans_blocks.", feat, smap, tn[lvl], nh[lvl])
        for i in range(cn[lvl]):
            out = basic_block(sd, "%s.conv_blocks.%d." % (name, i), out, k)
        return out, smap

    x0 = F.conv3d(x, sd["inc.conv1.weight"], padding=[i // 2 for i in ks[0]])
    x0 = basic_block(sd, "inc.conv2.", x0, ks[0])
    x1, _ = down("down1", x0, 0)
    x2, m2 = down("down2", x1, 1)
    x3, m3 = down("down3", x2, 2)
    x4, m4 = down("down4", x3, 3)
    maps = map_fusion(sd, [m2, m3, m4], cfg["fusion_depth"], cfg["fusion_heads"])
    out, smap = up("up1", x4, x3, maps[2], maps[1], 4, ks[3])
    out, smap = up("up2", out, x2, smap, maps[0], 5, ks[2])
    aux = None
    if cfg.get("aux_loss"):
        aux = F.conv3d(out, sd["aux_out.weight"], sd["aux_out.bias"])
        aux = F.interpolate(aux, size=x.shape[-3:], mode="trilinear", align_corners=True)
    out, smap = up("up3", out, x1, smap, None, 6, ks[1])
    out, smap = up("up4", out, x0, smap, None, 7, ks[0])
    out = F.conv3d(out, sd["outc.weight"], sd["outc.bias"])
    return [out, aux] if cfg.get("aux_loss") else out
