"""SwinUNETR whose every op runs in libb200seg.so — drop-in for the reference's ``model/dim3/swin_unetr.py`` (the
class ``get_model`` builds at model/utils.py:109-111): same constructor signature, same module tree, therefore the same
``state_dict`` keys / shapes / registration order (131 entries for the default depths) — including the modules the
reference takes from ``monai`` 1.1.0 (PatchEmbed, MLPBlock, UnetrBasicBlock, UnetrUpBlock, UnetOutBlock), which are
re-created here under their MONAI attribute names (``proj``, ``linear1/2``, ``layer.conv1.conv`` ..., cross-checked
against ``load_from``, swin_unetr.py:230-277,629-643).  MONAI's source is not under /root/reference, so those blocks
follow its published 1.1.0 semantics ("parity unpinned"); everything the reference file itself defines
is pinned (oracle/make_golden_swin*.py).

Underneath nothing of the reference's dataflow survives: activations are channels-last fp16/fp32 buffers; window
partition / shift / padding / masks / relative-position bias are index arithmetic inside one attention kernel; every
Linear is a 1x1x1 tcgen05 GEMM with bias / residual epilogues; the UnetResBlocks run as raw-output convs whose
InstanceNorm + LeakyReLU are applied by the NEXT conv's loader; ConvTranspose3d(k2,s2) and the k2s2 patch embedding
are GEMMs around a depth<->space shuffle."""
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_LRELU, ACT_NONE, call
from .medformer_ops import CatFn, ConvFn, GeluFn, LayerNormFn
from .ops import OutConvFn, PackedWeights, _dt, _need_cuda, _stream

IN_EPS = 1e-5         # nn.InstanceNorm3d default (monai get_norm_layer("instance"))
LN_EPS = 1e-5


# ----------------------------------------------------------------------------- autograd Functions
class DepthSpaceFn(torch.autograd.Function):
    """space-to-depth (to_depth=True: [B,D,H,W,C] -> [B,D/2,H/2,W/2,8C], channel q*C+c, q=(i*2+j)*2+k) or its inverse."""

    @staticmethod
    def forward(ctx, x, to_depth):
        _need_cuda(x)
        x = x.contiguous()
        B, D, H, W, C = x.shape
        if to_depth:
            y = torch.empty(B, D // 2, H // 2, W // 2, 8 * C, dtype=x.dtype, device=x.device)
            call("b200seg_space_to_depth", x.data_ptr(), y.data_ptr(), B, D // 2, H // 2, W // 2, C, 2, 2, 2, 0, _dt(x), _stream())
        else:
            y = torch.empty(B, 2 * D, 2 * H, 2 * W, C // 8, dtype=x.dtype, device=x.device)
            call("b200seg_space_to_depth", y.data_ptr(), x.data_ptr(), B, D, H, W, C // 8, 2, 2, 2, 1, _dt(x), _stream())
        ctx.to_depth = to_depth
        return y

    @staticmethod
    def backward(ctx, dy):
        return DepthSpaceFn.apply(dy, not ctx.to_depth), None


class DeriveWeightFn(torch.autograd.Function):
    """A GEMM weight that is a permutation of a parameter (ConvTranspose3d / strided patch-embedding kernels): written
    into a persistent buffer (stable address for the packer) and differentiated back through the permutation."""

    @staticmethod
    def forward(ctx, param, perm, shape, buf):
        buf.copy_(param.detach().permute(*perm).reshape(shape))
        ctx.meta = (perm, tuple(param.permute(*perm).shape))
        return buf.view(shape)

    @staticmethod
    def backward(ctx, d):
        perm, pshape = ctx.meta
        inv = [perm.index(i) for i in range(len(perm))]
        return d.reshape(pshape).permute(*inv).contiguous(), None, None, None


class ResOutFn(torch.autograd.Function):
    """y = act(IN(r2) + res): the output stage of monai's UnetResBlock (act = LeakyReLU 0.01); res = IN(r3) (1x1
    projection branch) or the block input itself.  With act = ReLU it is also the `relu(g1 + x1)` of Attention-UNet's
    gate (attention_unet_utils.py:31-34)."""

    @staticmethod
    def forward(ctx, r2, st2, r3, st3, act=ACT_LRELU):
        _need_cuda(r2)
        B, D, H, W, C = r2.shape
        y = torch.empty(B, D, H, W, C, dtype=r2.dtype, device=r2.device)
        r3 = r3.contiguous()
        call("b200seg_resblock_out_fwd", r2.data_ptr(), C, st2.data_ptr(), r3.data_ptr(), r3.shape[-1], 0,
             None if st3 is None else st3.data_ptr(), IN_EPS, act, y.data_ptr(), C, B, D * H * W, C, _dt(r2), _stream())
        ctx.save_for_backward(r2, st2, r3, st3, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        r2, st2, r3, st3, y = ctx.saved_tensors
        dy = dy.contiguous()
        B, D, H, W, C = r2.shape
        g = torch.empty_like(r2)
        sums = ops.zeros_scratch((B, C, 3), torch.float64, r2.device)
        call("b200seg_resblock_out_bwd_reduce", dy.data_ptr(), C, y.data_ptr(), C, r2.data_ptr(), C, st2.data_ptr(),
             r3.data_ptr(), r3.shape[-1], 0, None if st3 is None else st3.data_ptr(), IN_EPS, ctx.act, g.data_ptr(),
             sums.data_ptr(), B, D * H * W, C, _dt(r2), _stream())
        b2 = sums[:, :, :2].contiguous()
        dr2 = ops.in_bwd_apply(g, r2, 0, C, st2, b2, eps=IN_EPS)
        if st3 is not None:
            b3 = sums[:, :, [0, 2]].contiguous()
            dr3 = ops.in_bwd_apply(g, r3, 0, C, st3, b3, eps=IN_EPS)
        else:
            dr3 = g
        return dr2, None, dr3, None, None


class WindowAttnFn(torch.autograd.Function):
    """qkv [B,D,H,W,3C] -> attention output [B,D,H,W,C] (WindowAttention.forward between its two Linears, with the
    padding / shift / partition / mask / bias plumbing of forward_part1 inside the kernel)."""

    @staticmethod
    def forward(ctx, qkv, qkv_bias, table, heads, window, shift):
        _need_cuda(qkv)
        qkv = qkv.contiguous()
        B, D, H, W, C3 = qkv.shape
        C = C3 // 3
        dh = C // heads
        import ctypes
        win = (ctypes.c_int * 3)(*window)
        sft = (ctypes.c_int * 3)(*shift)
        lib = ops._lib.load()
        nbytes = lib.b200seg_window_attn_workspace(B, D, H, W, heads, win)
        lse = torch.empty(nbytes // 4, dtype=torch.float32, device=qkv.device)
        out = torch.empty(B, D, H, W, C, dtype=qkv.dtype, device=qkv.device)
        tb = table.detach().float().contiguous()
        qb = None if qkv_bias is None else qkv_bias.detach().float().contiguous()
        call("b200seg_window_attn_fwd", qkv.data_ptr(), None if qb is None else qb.data_ptr(), tb.data_ptr(), out.data_ptr(),
             lse.data_ptr(), B, D, H, W, heads, dh, win, sft, _dt(qkv), _stream())
        ctx.save_for_backward(qkv, tb, out, lse, qb if qb is not None else torch.empty(0, device=qkv.device))
        ctx.meta = (heads, dh, tuple(window), tuple(shift), qb is not None, table.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes
        qkv, tb, out, lse, qb = ctx.saved_tensors
        heads, dh, window, shift, has_b, tdtype = ctx.meta
        dout = dout.contiguous()
        B, D, H, W, C3 = qkv.shape
        win = (ctypes.c_int * 3)(*window)
        sft = (ctypes.c_int * 3)(*shift)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        dtable = torch.zeros_like(tb)
        dbias = torch.zeros(C3, dtype=torch.float32, device=qkv.device) if has_b else None
        call("b200seg_window_attn_bwd", qkv.data_ptr(), qb.data_ptr() if has_b else None, tb.data_ptr(), out.data_ptr(),
             dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), dtable.data_ptr(),
             dbias.data_ptr() if has_b else None, B, D, H, W, heads, dh, win, sft, _dt(qkv), _stream())
        return dqkv, dbias, dtable.to(tdtype), None, None, None


class SwinMergeFn(torch.autograd.Function):
    """PatchMerging's slice-gather + cat (v0.9 list with its duplicated slices, or V2's product order)."""

    @staticmethod
    def forward(ctx, x, v2):
        _need_cuda(x)
        x = x.contiguous()
        B, D, H, W, C = x.shape
        y = torch.empty(B, (D + 1) // 2, (H + 1) // 2, (W + 1) // 2, 8 * C, dtype=x.dtype, device=x.device)
        call("b200seg_swin_merge", x.data_ptr(), y.data_ptr(), B, D, H, W, C, 0, 1 if v2 else 0, _dt(x), _stream())
        ctx.meta = (x.shape, v2)
        return y

    @staticmethod
    def backward(ctx, dy):
        (B, D, H, W, C), v2 = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty(B, D, H, W, C, dtype=dy.dtype, device=dy.device)
        call("b200seg_swin_merge", dy.data_ptr(), dx.data_ptr(), B, D, H, W, C, 1, 1 if v2 else 0, _dt(dy), _stream())
        return dx, None


def _linear(pack, x, lin, residual=None):
    """nn.Linear on the channel axis of a channels-last tensor as a 1x1x1 conv GEMM (bias / residual in the epilogue)."""
    w = lin.weight.view(lin.weight.shape[0], lin.weight.shape[1], 1, 1, 1)
    packs = pack.get([w], x.dtype, x.shape[0], 0)
    y, _ = ConvFn.apply(x, None, residual, lin.bias, packs, (1, 1, 1), ACT_NONE, 0, IN_EPS, False, w)
    return y


def _ln(x, norm):
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)


# ----------------------------------------------------------------------------- MONAI-named blocks
class MLPBlock(nn.Module):
    """monai.networks.blocks.MLPBlock(hidden, mlp_dim, act='GELU', dropout 0): linear1 -> GELU -> linear2."""

    def __init__(self, hidden_size, mlp_dim):
        super().__init__()
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self._p1, self._p2 = PackedWeights(), PackedWeights()

    def forward(self, x, residual):
        return _linear(self._p2, GeluFn.apply(_linear(self._p1, x, self.linear1)), self.linear2, residual=residual)


class PatchEmbed(nn.Module):
    """monai PatchEmbed(patch_size=2, norm_layer=None): Conv3d(in, embed, kernel=2, stride=2) — swin_unetr.py:931-936."""

    def __init__(self, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=2, stride=2)
        self._free = {"pack": PackedWeights(), "buf": None}

    def forward(self, x):
        co, ci = self.proj.weight.shape[:2]
        xs = DepthSpaceFn.apply(x, True)                                  # [B, D/2, H/2, W/2, 8*ci], channel q*ci + c
        if self._free["buf"] is None or self._free["buf"].device != x.device:
            self._free["buf"] = torch.empty(co, 8 * ci, 1, 1, 1, dtype=torch.float32, device=x.device)
        w = DeriveWeightFn.apply(self.proj.weight, (0, 2, 3, 4, 1), (co, 8 * ci, 1, 1, 1), self._free["buf"])
        packs = self._free["pack"].get([w], x.dtype, x.shape[0], 0)
        y, _ = ConvFn.apply(xs, None, None, self.proj.bias, packs, (1, 1, 1), ACT_NONE, 0, IN_EPS, False, w)
        return y


class Convolution(nn.Module):
    """monai Convolution(conv_only=True): the parameter holder `conv` (never called)."""

    def __init__(self, ci, co, k, transposed=False, bias=False):
        super().__init__()
        if transposed:
            self.conv = nn.ConvTranspose3d(ci, co, kernel_size=k, stride=k, bias=bias)
        else:
            self.conv = nn.Conv3d(ci, co, kernel_size=k, padding=(k - 1) // 2, bias=bias)


class UnetResBlock(nn.Module):
    """monai UnetResBlock(kernel 3, stride 1, norm 'instance', act leakyrelu(0.01)): registration order conv1, conv2,
    lrelu, norm1, norm2, [conv3, norm3] — only the convs carry state."""

    def __init__(self, ci, co):
        super().__init__()
        self.conv1 = Convolution(ci, co, 3)
        self.conv2 = Convolution(co, co, 3)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        self.norm1 = nn.InstanceNorm3d(co)
        self.norm2 = nn.InstanceNorm3d(co)
        self.downsample = ci != co
        if self.downsample:
            self.conv3 = Convolution(ci, co, 1)
            self.norm3 = nn.InstanceNorm3d(co)
        self._p1, self._p2, self._p3 = PackedWeights(), PackedWeights(), PackedWeights()

    def forward(self, x):
        w1, w2 = self.conv1.conv.weight, self.conv2.conv.weight
        r1, st1 = ConvFn.apply(x, None, None, None, self._p1.get([w1], x.dtype, x.shape[0], 0), (3, 3, 3), ACT_NONE, 0, IN_EPS, True, w1)
        r2, st2 = ConvFn.apply(r1, st1, None, None, self._p2.get([w2], x.dtype, x.shape[0], 0), (3, 3, 3), ACT_LRELU, 0, IN_EPS, True, w2)
        if self.downsample:
            w3 = self.conv3.conv.weight
            r3, st3 = ConvFn.apply(x, None, None, None, self._p3.get([w3], x.dtype, x.shape[0], 0), (1, 1, 1), ACT_NONE, 0, IN_EPS, True, w3)
            return ResOutFn.apply(r2, st2, r3, st3, ACT_LRELU)
        return ResOutFn.apply(r2, st2, x, None, ACT_LRELU)


class UnetrBasicBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, res_block=False):
        super().__init__()
        if spatial_dims != 3 or kernel_size != 3 or stride != 1 or not res_block:
            raise ValueError("the B200 path implements the UnetrBasicBlock configuration SwinUNETR uses (3D, k3, s1, res_block)")
        _check_norm(norm_name)
        self.layer = UnetResBlock(in_channels, out_channels)

    def forward(self, x):
        return self.layer(x)


class UnetrUpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size, norm_name, res_block=False):
        super().__init__()
        if spatial_dims != 3 or kernel_size != 3 or upsample_kernel_size != 2 or not res_block:
            raise ValueError("the B200 path implements the UnetrUpBlock configuration SwinUNETR uses (3D, k3, up 2, res_block)")
        _check_norm(norm_name)
        self.transp_conv = Convolution(in_channels, out_channels, 2, transposed=True)
        self.conv_block = UnetResBlock(out_channels + out_channels, out_channels)
        self._free = {"pack": PackedWeights(), "buf": None}

    def forward(self, inp, skip):
        wt = self.transp_conv.conv.weight                                  # [Cin, Cout, 2, 2, 2]
        ci, co = wt.shape[:2]
        if self._free["buf"] is None or self._free["buf"].device != inp.device:
            self._free["buf"] = torch.empty(8 * co, ci, 1, 1, 1, dtype=torch.float32, device=inp.device)
        w = DeriveWeightFn.apply(wt, (2, 3, 4, 1, 0), (8 * co, ci, 1, 1, 1), self._free["buf"])   # row q*Cout + co
        packs = self._free["pack"].get([w], inp.dtype, inp.shape[0], 0)
        y8, _ = ConvFn.apply(inp, None, None, None, packs, (1, 1, 1), ACT_NONE, 0, IN_EPS, False, w)
        up = DepthSpaceFn.apply(y8, False)
        return self.conv_block(CatFn.apply(up, skip))


class UnetOutBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels):
        super().__init__()
        self.conv = Convolution(in_channels, out_channels, 1, bias=True)
        self._pack = PackedWeights()

    def forward(self, x):
        w, b = self.conv.conv.weight, self.conv.conv.bias
        return OutConvFn.apply(x, w, b, self._pack.get([w], x.dtype, x.shape[0]))


def _check_norm(norm_name):
    name = norm_name[0] if isinstance(norm_name, (tuple, list)) else norm_name
    if str(name).lower() != "instance":
        raise ValueError("the B200 path implements norm_name='instance' only, got %r" % (norm_name,))


# ----------------------------------------------------------------------------- the vendored Swin classes
class WindowAttention(nn.Module):
    """swin_unetr.py:384-490."""

    def __init__(self, dim, num_heads, window_size, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if attn_drop or proj_drop:
            raise ValueError("dropout is not implemented by the B200 path (the reference trains SwinUNETR with 0)")
        if len(window_size) != 3:
            raise ValueError("3D windows only")
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        ws = self.window_size
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * ws[0] - 1) * (2 * ws[1] - 1) * (2 * ws[2] - 1), num_heads))
        coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2]), indexing="ij"))
        flat = torch.flatten(coords, 1)
        rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws[0] - 1
        rel[:, :, 1] += ws[1] - 1
        rel[:, :, 2] += ws[2] - 1
        rel[:, :, 0] *= (2 * ws[1] - 1) * (2 * ws[2] - 1)
        rel[:, :, 1] *= 2 * ws[2] - 1
        self.register_buffer("relative_position_index", rel.sum(-1))        # state_dict contract; the kernel derives it
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        self._pq, self._pp = PackedWeights(), PackedWeights()

    def forward(self, xn, shortcut, shift):
        """xn = norm1(x) [B,D,H,W,C]; returns shortcut + proj(attention(...))."""
        qkv = _linear(self._pq, xn, self.qkv)
        att = WindowAttnFn.apply(qkv, self.qkv.bias, self.relative_position_bias_table, self.num_heads, self.window_size, shift)
        return _linear(self._pp, att, self.proj, residual=shortcut)


class SwinTransformerBlock(nn.Module):
    """swin_unetr.py:493-657 (drop_path 0; checkpointing is a memory knob the B200 path does not need)."""

    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio=4.0, qkv_bias=True, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer="GELU", norm_layer=nn.LayerNorm, use_checkpoint=False):
        super().__init__()
        if drop or drop_path:
            raise ValueError("dropout / drop_path are not implemented by the B200 path (the reference uses 0)")
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, tuple(window_size), tuple(shift_size)
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=self.window_size, num_heads=num_heads, qkv_bias=qkv_bias)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = self.attn(_ln(x, self.norm1), x, self.shift_size)              # x + attention      (:645-651)
        return self.mlp(_ln(x, self.norm2), x)                             # x + mlp(norm2(x))  (:652-656)


class PatchMergingV2(nn.Module):
    """swin_unetr.py:660-704."""
    V2 = True

    def __init__(self, dim, norm_layer=nn.LayerNorm, spatial_dims=3):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("3D only")
        self.dim = dim
        self.reduction = nn.Linear(8 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(8 * dim)
        self._pack = PackedWeights()

    def forward(self, x):
        return _linear(self._pack, _ln(SwinMergeFn.apply(x, self.V2), self.norm), self.reduction)


class PatchMerging(PatchMergingV2):
    """The v0.9 variant the reference instantiates by default, duplicated slices included (swin_unetr.py:707-731)."""
    V2 = False


MERGING_MODE = {"merging": PatchMerging, "mergingv2": PatchMergingV2}


class BasicLayer(nn.Module):
    """swin_unetr.py:776-907."""

    def __init__(self, dim, depth, num_heads, window_size, drop_path, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0,
                 norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.window_size = tuple(window_size)
        self.shift_size = tuple(i // 2 for i in window_size)
        self.no_shift = tuple(0 for _ in window_size)
        self.depth = depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, num_heads=num_heads, window_size=self.window_size,
                                 shift_size=self.no_shift if (i % 2 == 0) else self.shift_size, mlp_ratio=mlp_ratio,
                                 qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path, norm_layer=norm_layer)
            for i in range(depth)])
        self.downsample = downsample
        if callable(self.downsample):
            self.downsample = downsample(dim=dim, norm_layer=norm_layer, spatial_dims=len(self.window_size))

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        if self.downsample is not None:
            x = self.downsample(x)
        return x


class SwinTransformer(nn.Module):
    """swin_unetr.py:910-1000."""

    def __init__(self, in_chans, embed_dim, window_size, patch_size, depths, num_heads, mlp_ratio=4.0, qkv_bias=True,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm, patch_norm=False,
                 use_checkpoint=False, spatial_dims=3, downsample="merging"):
        super().__init__()
        if spatial_dims != 3 or tuple(patch_size) != (2, 2, 2) or patch_norm:
            raise ValueError("the B200 path implements the 3D, patch 2, patch_norm=False configuration SwinUNETR uses")
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise ValueError("dropout / drop_path are not implemented by the B200 path (the reference uses 0)")
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.window_size = tuple(window_size)
        self.patch_embed = PatchEmbed(in_chans, embed_dim)
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.layers1, self.layers2, self.layers3, self.layers4 = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        down = MERGING_MODE[downsample] if isinstance(downsample, str) else downsample
        for i_layer in range(self.num_layers):
            layer = BasicLayer(dim=int(embed_dim * 2 ** i_layer), depth=depths[i_layer], num_heads=num_heads[i_layer],
                               window_size=self.window_size, drop_path=0.0, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                               norm_layer=norm_layer, downsample=down)
            [self.layers1, self.layers2, self.layers3, self.layers4][i_layer].append(layer)
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self._ones = {}

    def proj_out(self, x, normalize):
        """affine-free LayerNorm over channels (swin_unetr.py:970-983)."""
        if not normalize:
            return x
        C = x.shape[-1]
        key = (C, x.device)
        if key not in self._ones:
            self._ones[key] = (torch.ones(C, device=x.device), torch.zeros(C, device=x.device))
        g, b = self._ones[key]
        return LayerNormFn.apply(x, g, b, LN_EPS)

    def forward(self, x, normalize=True):
        x0 = self.patch_embed(x)
        x1 = self.layers1[0](x0)
        x2 = self.layers2[0](x1)
        x3 = self.layers3[0](x2)
        x4 = self.layers4[0](x3)
        return [self.proj_out(t, normalize) for t in (x0, x1, x2, x3, x4)]


class SwinUNETR(nn.Module):
    """swin_unetr.py:32-292."""

    def __init__(self, img_size, in_channels, out_channels, depths=(2, 2, 2, 0), num_heads=(3, 6, 12, 24), feature_size=24,
                 norm_name="instance", drop_rate=0.0, attn_drop_rate=0.0, dropout_path_rate=0.0, normalize=True,
                 use_checkpoint=False, spatial_dims=3, downsample="merging"):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("the B200 path implements the 3D SwinUNETR")
        img_size = tuple(img_size) if isinstance(img_size, (list, tuple)) else (img_size,) * 3
        for m in img_size:
            if m % 32 != 0:
                raise ValueError("input image size (img_size) should be divisible by stage-wise image resolution.")
        if feature_size % 12 != 0:
            raise ValueError("feature_size should be divisible by 12.")
        self.normalize = normalize
        fs = feature_size
        self.swinViT = SwinTransformer(in_chans=in_channels, embed_dim=fs, window_size=(7, 7, 7), patch_size=(2, 2, 2),
                                       depths=depths, num_heads=num_heads, mlp_ratio=4.0, qkv_bias=True, drop_rate=drop_rate,
                                       attn_drop_rate=attn_drop_rate, drop_path_rate=dropout_path_rate, norm_layer=nn.LayerNorm,
                                       use_checkpoint=use_checkpoint, spatial_dims=3, downsample=downsample)
        kw = dict(spatial_dims=3, kernel_size=3, stride=1, norm_name=norm_name, res_block=True)
        self.encoder1 = UnetrBasicBlock(in_channels=in_channels, out_channels=fs, **kw)
        self.encoder2 = UnetrBasicBlock(in_channels=fs, out_channels=fs, **kw)
        self.encoder3 = UnetrBasicBlock(in_channels=2 * fs, out_channels=2 * fs, **kw)
        self.encoder4 = UnetrBasicBlock(in_channels=4 * fs, out_channels=4 * fs, **kw)
        self.encoder10 = UnetrBasicBlock(in_channels=16 * fs, out_channels=16 * fs, **kw)
        up = dict(spatial_dims=3, kernel_size=3, upsample_kernel_size=2, norm_name=norm_name, res_block=True)
        self.decoder5 = UnetrUpBlock(in_channels=16 * fs, out_channels=8 * fs, **up)
        self.decoder4 = UnetrUpBlock(in_channels=8 * fs, out_channels=4 * fs, **up)
        self.decoder3 = UnetrUpBlock(in_channels=4 * fs, out_channels=2 * fs, **up)
        self.decoder2 = UnetrUpBlock(in_channels=2 * fs, out_channels=fs, **up)
        self.decoder1 = UnetrUpBlock(in_channels=fs, out_channels=fs, **up)
        self.out = UnetOutBlock(spatial_dims=3, in_channels=fs, out_channels=out_channels)
        self._packs = ops.PackRegistry(self)

    def forward(self, x_in):
        if not x_in.is_cuda:
            raise ops._lib.B200SegError("b200seg.SwinUNETR runs on a B200 only — there is no CPU fallback")
        with ops.on_device(x_in):
            return self._forward(x_in)

    def _forward(self, x_in):
        dt = ops.compute_dtype()
        self._packs.refresh()
        x = x_in.permute(0, 2, 3, 4, 1).to(dt).contiguous()                # NDHWC working layout
        hs = self.swinViT(x, self.normalize)
        enc0 = self.encoder1(x)
        enc1 = self.encoder2(hs[0])
        enc2 = self.encoder3(hs[1])
        enc3 = self.encoder4(hs[2])
        dec4 = self.encoder10(hs[4])
        dec3 = self.decoder5(dec4, hs[3])
        dec2 = self.decoder4(dec3, enc3)
        dec1 = self.decoder3(dec2, enc2)
        dec0 = self.decoder2(dec1, enc1)
        out = self.decoder1(dec0, enc0)
        return self.out(out).permute(0, 4, 1, 2, 3)                         # logical NCDHW over the NDHWC buffer
