"""Attention-UNet on the b200seg blocks — drop-in for the reference's ``model/dim3/attention_unet.py:8-46`` (the class
``get_model`` builds for ``args.model == 'attention_unet'``, model/utils.py:88-90) and
``attention_unet_utils.py:7-64``: same constructor signature, same module tree (``inc``, ``down1..4``,
``up1..4.{conv_ch, attn.{W_g, W_x, psi}, conv}``, ``outc``), therefore the same ``state_dict`` keys / shapes /
registration order and default initialisation.

Encoder, residual blocks, pooling and the head are the 3D UNet path unchanged.  The additive gate
runs as: two 1x1x1 tcgen05 GEMMs (W_g, W_x; their InstanceNorm sums come out of the conv epilogues), one fused
``relu(IN(.) + IN(.))`` pass, and the ``csrc/attn_gate.cu`` kernels (psi dot product + its norm sums; sigmoid gate +
the IN sums of the gated skip for the next block's loader)."""
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU, call
from .medformer_ops import CatFn, ConvFn
from .ops import OutConvFn, PackedWeights, _dt, _need_cuda, _stream
from .swin_unetr import ResOutFn
from .unet3d import _triple, down_block, get_block, inconv

GATE_EPS = 1e-5      # nn.InstanceNorm3d default eps (attention_unet_utils.py:13,17,21); ResOutFn uses the same value


class UpsampleStatsFn(torch.autograd.Function):
    """F.interpolate(x1, size, 'trilinear', align_corners=True) (attention_unet_utils.py:56) as its own tensor — the
    gate reads it (W_g) before it is concatenated — with the IN sums of the result."""

    @staticmethod
    def forward(ctx, x, size):
        _need_cuda(x)
        x = x.contiguous()
        B, Di, Hi, Wi, C = x.shape
        Do, Ho, Wo = size
        y = torch.empty(B, Do, Ho, Wo, C, dtype=x.dtype, device=x.device)
        st = ops.new_stats(B, C, x.device)
        call("b200seg_upsample_trilinear_fwd", x.data_ptr(), C, 0, y.data_ptr(), C, 0, st.data_ptr(),
             B, Di, Hi, Wi, Do, Ho, Wo, C, _dt(x), _stream())
        ctx.meta = (x.shape, size)
        ctx.mark_non_differentiable(st)
        return y, st

    @staticmethod
    def backward(ctx, dy, _):
        (B, Di, Hi, Wi, C), (Do, Ho, Wo) = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty(B, Di, Hi, Wi, C, dtype=dy.dtype, device=dy.device)
        call("b200seg_upsample_trilinear_bwd", dy.data_ptr(), C, 0, dx.data_ptr(), C, 0, 0,
             B, Di, Hi, Wi, Do, Ho, Wo, C, _dt(dy), _stream())
        return dx, None


class AttnGateFn(torch.autograd.Function):
    """out = x * sigmoid(IN(psi_conv(t))) (attention_unet_utils.py:20-22,35-37) and the IN sums of out."""

    @staticmethod
    def forward(ctx, x, t, w_psi):
        _need_cuda(x)
        x, t = x.contiguous(), t.contiguous()
        B, D, H, W, Cx = x.shape
        Ct = t.shape[-1]
        V = D * H * W
        w = w_psi.detach().reshape(-1).float().contiguous()
        p = torch.empty(B, V, dtype=torch.float32, device=x.device)
        pst = ops.zeros_scratch((B, 2), torch.float64, x.device)
        out = torch.empty_like(x)
        ost = ops.new_stats(B, Cx, x.device)
        call("b200seg_attn_gate_fwd", t.data_ptr(), Ct, w.data_ptr(), x.data_ptr(), Cx, 0, GATE_EPS, p.data_ptr(), pst.data_ptr(),
             out.data_ptr(), Cx, 0, ost.data_ptr(), B, V, Ct, Cx, _dt(x), _stream())
        ctx.save_for_backward(x, t, w, p, pst)
        ctx.wmeta = (w_psi.shape, w_psi.dtype)
        ctx.mark_non_differentiable(ost)
        return out, ost

    @staticmethod
    def backward(ctx, dout, _):
        x, t, w, p, pst = ctx.saved_tensors
        dout = dout.contiguous()
        B, D, H, W, Cx = x.shape
        Ct = t.shape[-1]
        V = D * H * W
        dx, dt = torch.empty_like(x), torch.empty_like(t)
        dw = torch.zeros(Ct, dtype=torch.float32, device=x.device)
        dz = torch.empty(B, V, dtype=torch.float32, device=x.device)
        bs = ops.zeros_scratch((B, 2), torch.float64, x.device)
        call("b200seg_attn_gate_bwd", dout.data_ptr(), Cx, 0, x.data_ptr(), Cx, 0, t.data_ptr(), Ct, w.data_ptr(), p.data_ptr(),
             pst.data_ptr(), GATE_EPS, dx.data_ptr(), dt.data_ptr(), dw.data_ptr(), dz.data_ptr(), bs.data_ptr(), B, V, Ct, Cx,
             _dt(x), _stream())
        shape, wdtype = ctx.wmeta
        return dx, dt, dw.reshape(shape).to(wdtype)


def _conv1(pack, x, w):
    """raw 1x1x1 conv (no norm / act on the input, no bias) + the IN sums of its output"""
    packs = pack.get([w], x.dtype, x.shape[0], 0)
    return ConvFn.apply(x, None, None, None, packs, (1, 1, 1), ACT_NONE, 0, GATE_EPS, True, w)


class AttentionBlock(nn.Module):
    """attention_unet_utils.py:7-37.  The Sequentials only hold the parameters (keys ``W_g.0.weight`` ...); the
    InstanceNorm / Sigmoid / ReLU members are stateless."""

    def __init__(self, g_ch, l_ch, int_ch):
        super().__init__()
        if int_ch % 8 or g_ch % 8 or l_ch % 8:
            raise ValueError("the B200 attention gate needs channel counts that are multiples of 8 (got %d, %d, %d)"
                             % (g_ch, l_ch, int_ch))
        self.W_g = nn.Sequential(nn.Conv3d(g_ch, int_ch, kernel_size=1, stride=1, padding=0, bias=False), nn.InstanceNorm3d(int_ch))
        self.W_x = nn.Sequential(nn.Conv3d(l_ch, int_ch, kernel_size=1, stride=1, padding=0, bias=False), nn.InstanceNorm3d(int_ch))
        self.psi = nn.Sequential(nn.Conv3d(int_ch, 1, kernel_size=1, stride=1, padding=0, bias=False), nn.InstanceNorm3d(1),
                                 nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)
        self._pack_g = PackedWeights()
        self._pack_x = PackedWeights()

    def forward(self, g, x):
        g1, sg = _conv1(self._pack_g, g, self.W_g[0].weight)            # :29
        x1, sx = _conv1(self._pack_x, x, self.W_x[0].weight)            # :30
        t = ResOutFn.apply(g1, sg, x1, sx, ACT_RELU)                    # relu(IN(g1) + IN(x1)), :13,17,32
        return AttnGateFn.apply(x, t, self.psi[0].weight)               # :33-37


class attention_up_block(nn.Module):
    """attention_unet_utils.py:39-64.  ``conv_ch`` exists in the reference but is never called; it is kept so the
    state_dict matches (and, as there, gets no gradient: wrap with find_unused_parameters=True under DDP,
    train_ddp.py:353)."""

    def __init__(self, in_ch, out_ch, num_block, block, kernel_size=[3, 3, 3], up_scale=[2, 2, 2]):
        super().__init__()
        self.conv_ch = nn.Conv3d(in_ch, out_ch, kernel_size=1)
        self.up_scale = _triple(up_scale)
        self.attn = AttentionBlock(in_ch, out_ch, out_ch // 2)
        layers = [block(in_ch + out_ch, out_ch, kernel_size=_triple(kernel_size))]
        for _ in range(num_block - 1):
            layers.append(block(out_ch, out_ch, kernel_size=_triple(kernel_size)))
        self.conv = nn.Sequential(*layers)

    def forward(self, a1, a2):
        low, _ = a1
        skip, _ = a2
        up, up_st = UpsampleStatsFn.apply(low, tuple(skip.shape[1:4]))                   # :56
        gated, g_st = self.attn(up, skip)                                                # :58
        cat = CatFn.apply(gated, up)                                                     # :61 cat([x2, x1])
        return self.conv((cat, torch.cat([g_st, up_st], dim=1).contiguous()))


class AttentionUNet(nn.Module):
    """model/dim3/attention_unet.py:8-46.  `norm` must be 'in'."""

    def __init__(self, in_ch, base_ch, scale, kernel_size, num_classes=1, block='SingleConv', pool=True, norm='bn'):
        super().__init__()
        if norm not in ('in', nn.InstanceNorm3d):
            raise ValueError("the B200 path implements InstanceNorm ('in') only, got norm=%r" % (norm,))
        num_block = 2
        blk = get_block(block)
        c = base_ch
        self.inc = inconv(in_ch, c, block=blk, kernel_size=kernel_size[0])
        self.down1 = down_block(c, 2 * c, num_block, blk, kernel_size[1], scale[0], pool)
        self.down2 = down_block(2 * c, 4 * c, num_block, blk, kernel_size[2], scale[1], pool)
        self.down3 = down_block(4 * c, 8 * c, num_block, blk, kernel_size[3], scale[2], pool)
        self.down4 = down_block(8 * c, 10 * c, num_block, blk, kernel_size[4], scale[3], pool)
        self.up1 = attention_up_block(10 * c, 8 * c, num_block, blk, kernel_size[3], scale[3])
        self.up2 = attention_up_block(8 * c, 4 * c, num_block, blk, kernel_size[2], scale[2])
        self.up3 = attention_up_block(4 * c, 2 * c, num_block, blk, kernel_size[1], scale[1])
        self.up4 = attention_up_block(2 * c, c, num_block, blk, kernel_size[0], scale[0])
        self.outc = nn.Conv3d(c, num_classes, kernel_size=1)
        self._pack_out = PackedWeights()
        self._packs = ops.PackRegistry(self)

    def forward(self, x):
        if not x.is_cuda:
            raise ops._lib.B200SegError("b200seg.AttentionUNet runs on a B200 only — there is no CPU fallback")
        with ops.on_device(x):
            return self._forward(x)

    def _forward(self, x):
        dt = ops.compute_dtype()
        self._packs.refresh()
        xin = x.permute(0, 2, 3, 4, 1).to(dt).contiguous()
        x1 = self.inc(xin)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.down3(x3)
        x5 = self.down4(x4)
        out = self.up1(x5, x4)
        out = self.up2(out, x3)
        out = self.up3(out, x2)
        out = self.up4(out, x1)
        w, b = self.outc.weight, self.outc.bias
        logits = OutConvFn.apply(out[0], w, b, self._pack_out.get([w], dt, x.shape[0]))
        return logits.permute(0, 4, 1, 2, 3)
