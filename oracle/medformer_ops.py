"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the MedFormer operators that libb200seg implements natively so far:
  * the bidirectional attention core, BidirectionAttention.forward medformer_utils.py:63-97 (between the q/v
    projections :67-68 and the output projections :95-96);
  * the depthwise 3-D convolution inside DepthwiseSeparableConv, conv_layers.py:135-143,153-155.
Pinned by oracle/make_golden_medformer.py against the unmodified reference modules (tests/golden/biattn_*.pt,
dwconv_*.pt).  Tensors are NCDHW like the reference's.
"""
import torch
import torch.nn.functional as F


def _split_heads(x, heads):
    """'b (dim_head heads) d h w -> b heads (d h w) dim_head'  (rearrange1, medformer_utils.py:43-51)."""
    b, l = x.shape[:2]
    return x.reshape(b, l // heads, heads, -1).permute(0, 2, 3, 1)


def _merge_heads(x, spatial):
    """'b heads n dim_head -> b (dim_head heads) *spatial'  (rearrange2, medformer_utils.py:52-59)."""
    b, heads, n, dh = x.shape
    return x.permute(0, 3, 1, 2).reshape(b, dh * heads, *spatial)


def bidirection_attention_core(feat_q, feat_v, map_q, map_v, heads):
    """Returns (feat_out [B,inner,D,H,W], map_out [B,inner,*map_size]) — medformer_utils.py:70-91, dropout p=0."""
    dim_head = feat_q.shape[1] // heads
    fs, ms = feat_q.shape[2:], map_q.shape[2:]
    fq, fv, mq, mv = (_split_heads(t, heads) for t in (feat_q, feat_v, map_q, map_v))
    attn = torch.einsum("bhid,bhjd->bhij", fq, mq) * dim_head ** -0.5          # :77-78
    a_row = F.softmax(attn, dim=-1)                                            # :80
    a_col = F.softmax(attn, dim=-2)                                            # :82
    feat_out = torch.einsum("bhij,bhjd->bhid", a_row, mv)                      # :84
    map_out = torch.einsum("bhji,bhjd->bhid", a_col, fv)                       # :89
    return _merge_heads(feat_out, fs), _merge_heads(map_out, ms)


def depthwise_conv3d(x, w, bias=None):
    """nn.Conv3d(C, C, k, stride=1, padding=k//2, groups=C)  (conv_layers.py:135-143). w: [C,1,kd,kh,kw]."""
    pad = [k // 2 for k in w.shape[2:]]
    return F.conv3d(x, w, bias, stride=1, padding=pad, groups=x.shape[1])
