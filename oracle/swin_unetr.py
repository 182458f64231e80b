"""ORACLE (test infrastructure — never imported by the product path): SwinUNETR as one pure function of a state_dict.

The transformer half follows the reference's VENDORED code and is pinned to it (oracle/swin_ops.py,
oracle/make_golden_swin.py, and whole-model by oracle/make_golden_swin_unetr.py):
  SwinUNETR.forward            model/dim3/swin_unetr.py:279-292
  SwinTransformer.forward      :985-1000   (proj_out = affine-free LayerNorm over channels, :970-983)
  BasicLayer.forward           :886-905    (mask built once per stage, blocks, then the v0.9 PatchMerging)
  SwinTransformerBlock         :554-606, :608-609, :645-657
The seven symbols the file imports from `monai` (1.1.0, requirements.txt:41) are NOT under /root/reference; they are
restated here from MONAI 1.1.0's published semantics — "PARITY UNPINNED" for exactly these pieces:
  PatchEmbed        Conv3d(in, embed, kernel=patch, stride=patch) (+bias), no norm (patch_norm=False, :931-936)
  MLPBlock          linear1 -> GELU -> linear2 (dropout 0)                                     (:552, names :640-643)
  UnetrBasicBlock(res_block=True) = UnetResBlock: conv k3 -> IN -> LeakyReLU(0.01) -> conv k3 -> IN,
                    residual = conv1x1 -> IN when Cin != Cout, add, LeakyReLU                    (call sites :129-177)
  UnetrUpBlock      ConvTranspose3d(k=2, s=2, no bias) -> cat([up, skip]) -> UnetResBlock(2*Cout -> Cout)  (:179-226)
  UnetOutBlock      Conv3d 1x1 with bias                                                          (:228)
InstanceNorm3d there is torch's default (eps 1e-5, affine=False), all convs of the blocks have bias=False.
"""
import torch
import torch.nn.functional as F

from . import swin_ops as so

IN_EPS = 1e-5
LRELU = 0.01


def _in(x):
    return F.instance_norm(x, eps=IN_EPS)


def res_block(sd, pre, x):
    """monai UnetResBlock (norm 'instance', act leakyrelu 0.01, kernel 3, stride 1)."""
    out = F.conv3d(x, sd[pre + "conv1.conv.weight"], padding=1)
    out = F.leaky_relu(_in(out), LRELU)
    out = _in(F.conv3d(out, sd[pre + "conv2.conv.weight"], padding=1))
    res = x
    if pre + "conv3.conv.weight" in sd:
        res = _in(F.conv3d(x, sd[pre + "conv3.conv.weight"]))
    return F.leaky_relu(out + res, LRELU)


def up_block(sd, pre, x, skip):
    up = F.conv_transpose3d(x, sd[pre + "transp_conv.conv.weight"], stride=2)
    return res_block(sd, pre + "conv_block.", torch.cat([up, skip], dim=1))


def proj_out(x):
    """affine-free LayerNorm over channels of an NCDHW tensor — swin_unetr.py:970-983."""
    c = x.shape[1]
    return F.layer_norm(x.permute(0, 2, 3, 4, 1), (c,)).permute(0, 4, 1, 2, 3)


def swin_block(sd, pre, x, heads, window_size, shift_size, mask):
    """SwinTransformerBlock.forward (drop_path 0) on x [b,d,h,w,c] — swin_unetr.py:645-657."""
    p = {"norm1_w": sd[pre + "norm1.weight"], "norm1_b": sd[pre + "norm1.bias"],
         "qkv_w": sd[pre + "attn.qkv.weight"], "qkv_b": sd.get(pre + "attn.qkv.bias"),
         "proj_w": sd[pre + "attn.proj.weight"], "proj_b": sd[pre + "attn.proj.bias"],
         "bias_table": sd[pre + "attn.relative_position_bias_table"]}
    x = x + so.swin_block_part1(x, p, heads, window_size, shift_size, mask)
    c = x.shape[-1]
    h = F.layer_norm(x, (c,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    h = F.linear(F.gelu(F.linear(h, sd[pre + "mlp.linear1.weight"], sd[pre + "mlp.linear1.bias"])),
                 sd[pre + "mlp.linear2.weight"], sd[pre + "mlp.linear2.bias"])
    return x + h


def basic_layer(sd, pre, x, depth, heads, window_size):
    """BasicLayer.forward on NCDHW x — swin_unetr.py:886-905."""
    b, c, d, h, w = x.shape
    shift = tuple(i // 2 for i in window_size)
    ws, ss = so.get_window_size((d, h, w), window_size, shift)
    x = x.permute(0, 2, 3, 4, 1)
    dp, hp, wp = [-(-s // k) * k for s, k in zip((d, h, w), ws)]
    mask = so.compute_mask([dp, hp, wp], ws, ss).to(device=x.device, dtype=x.dtype) if any(s > 0 for s in ss) else None
    for i in range(depth):
        x = swin_block(sd, "%sblocks.%d." % (pre, i), x, heads, window_size, (0, 0, 0) if i % 2 == 0 else shift, mask)
    x = x.reshape(b, d, h, w, -1)
    x = so.patch_merging(x, sd[pre + "downsample.norm.weight"], sd[pre + "downsample.norm.bias"],
                         sd[pre + "downsample.reduction.weight"])
    return x.permute(0, 4, 1, 2, 3)


def swin_unetr_forward(sd, x, depths=(2, 2, 2, 0), num_heads=(3, 6, 12, 24), window_size=(7, 7, 7), normalize=True):
    """x [B, in_ch, D, H, W] -> logits [B, classes, D, H, W]."""
    x0 = F.conv3d(x, sd["swinViT.patch_embed.proj.weight"], sd["swinViT.patch_embed.proj.bias"], stride=2)
    hs = [x0]
    cur = x0
    for i in range(4):
        cur = basic_layer(sd, "swinViT.layers%d.0." % (i + 1), cur.contiguous(), depths[i], num_heads[i], window_size)
        hs.append(cur)
    hs = [proj_out(t) for t in hs] if normalize else hs
    enc0 = res_block(sd, "encoder1.layer.", x)
    enc1 = res_block(sd, "encoder2.layer.", hs[0])
    enc2 = res_block(sd, "encoder3.layer.", hs[1])
    enc3 = res_block(sd, "encoder4.layer.", hs[2])
    dec4 = res_block(sd, "encoder10.layer.", hs[4])
    dec3 = up_block(sd, "decoder5.", dec4, hs[3])
    dec2 = up_block(sd, "decoder4.", dec3, enc3)
    dec1 = up_block(sd, "decoder3.", dec2, enc2)
    dec0 = up_block(sd, "decoder2.", dec1, enc1)
    out = up_block(sd, "decoder1.", dec0, enc0)
    return F.conv3d(out, sd["out.conv.conv.weight"], sd["out.conv.conv.bias"])


def swin_unetr_param_shapes(in_ch, classes, fs, depths=(2, 2, 2, 0), num_heads=(3, 6, 12, 24), window_size=(7, 7, 7)):
    """state_dict key -> shape in registration order (swinViT, encoder1..4, encoder10, decoder5..1, out)."""
    out = {}
    out["swinViT.patch_embed.proj.weight"] = (fs, in_ch, 2, 2, 2)
    out["swinViT.patch_embed.proj.bias"] = (fs,)
    T = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
    for i in range(4):
        dim = fs * 2 ** i
        pre = "swinViT.layers%d.0." % (i + 1)
        for j in range(depths[i]):
            b = "%sblocks.%d." % (pre, j)
            out[b + "norm1.weight"] = (dim,); out[b + "norm1.bias"] = (dim,)
            out[b + "attn.relative_position_bias_table"] = (T, num_heads[i])
            out[b + "attn.qkv.weight"] = (3 * dim, dim); out[b + "attn.qkv.bias"] = (3 * dim,)
            out[b + "attn.proj.weight"] = (dim, dim); out[b + "attn.proj.bias"] = (dim,)
            out[b + "norm2.weight"] = (dim,); out[b + "norm2.bias"] = (dim,)
            out[b + "mlp.linear1.weight"] = (4 * dim, dim); out[b + "mlp.linear1.bias"] = (4 * dim,)
            out[b + "mlp.linear2.weight"] = (dim, 4 * dim); out[b + "mlp.linear2.bias"] = (dim,)
        out[pre + "downsample.reduction.weight"] = (2 * dim, 8 * dim)
        out[pre + "downsample.norm.weight"] = (8 * dim,); out[pre + "downsample.norm.bias"] = (8 * dim,)

    def res(pre, ci, co):
        out[pre + "conv1.conv.weight"] = (co, ci, 3, 3, 3)
        out[pre + "conv2.conv.weight"] = (co, co, 3, 3, 3)
        if ci != co:
            out[pre + "conv3.conv.weight"] = (co, ci, 1, 1, 1)
    res("encoder1.layer.", in_ch, fs)
    res("encoder2.layer.", fs, fs)
    res("encoder3.layer.", 2 * fs, 2 * fs)
    res("encoder4.layer.", 4 * fs, 4 * fs)
    res("encoder10.layer.", 16 * fs, 16 * fs)
    for name, ci, co in (("decoder5.", 16 * fs, 8 * fs), ("decoder4.", 8 * fs, 4 * fs), ("decoder3.", 4 * fs, 2 * fs),
                         ("decoder2.", 2 * fs, fs), ("decoder1.", fs, fs)):
        out[name + "transp_conv.conv.weight"] = (ci, co, 2, 2, 2)
        res(name + "conv_block.", 2 * co, co)
    out["out.conv.conv.weight"] = (classes, fs, 1, 1, 1)
    out["out.conv.conv.bias"] = (classes,)
    return out
