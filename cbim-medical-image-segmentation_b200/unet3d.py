"""3D UNet / ResUNet whose every op runs in libb200seg.so — drop-in for the reference's
``model/dim3/unet.py:12-64`` (+ ``unet_utils.py``, ``conv_layers.py``): same constructor signature, same
module tree and therefore the same ``state_dict`` keys/shapes, parameter registration order (EMA zips
parameters, training/utils.py:101) and default initialisation (the parameter holders are plain
``nn.Conv3d`` modules that are never called, so the same seed draws the same weights).

What differs is everything underneath: activations are NDHWC fp16/fp32 buffers that never leave the
kernels' layout; InstanceNorm+ReLU never exist as tensors (they are applied in the consuming conv's
loader from sums produced in the producing kernel's epilogue); conv1+shortcut of a BasicBlock are one
GEMM; upsample+concat is one kernel.
"""
import torch
import torch.nn as nn

from . import ops
from .ops import (BasicBlockFn, MaxPoolFn, OutConvFn, PackedWeights, SingleConvFn, StemConvFn, UpCatFn)


def _triple(k):
    return [k] * 3 if isinstance(k, int) else list(k)


def _check_kernel(k):
    if any(int(v) % 2 == 0 for v in k):
        raise ValueError("b200seg conv3d supports odd kernel sizes only, got %s" % (k,))


class ConvNormAct(nn.Module):
    """Parameter holder mirroring conv_layers.py:16-53 (`conv` is the only stateful child; the
    InstanceNorm has affine=False and no running stats, so it contributes no state)."""

    def __init__(self, in_ch, out_ch, kernel_size=3, padding=1, bias=False):
        super().__init__()
        self.conv = nn.Conv3d(in_ch, out_ch, kernel_size=kernel_size, padding=padding, bias=bias)


class SingleConv(nn.Module):
    """conv -> IN -> ReLU (conv_layers.py:56-68)."""

    def __init__(self, in_ch, out_ch, kernel_size=[3, 3, 3], stride=1):
        super().__init__()
        if stride != 1:
            raise ValueError("strided conv blocks (pool=False) are not supported by the B200 path")
        ks = _triple(kernel_size)
        _check_kernel(ks)
        self.ksize = tuple(ks)
        self.conv = ConvNormAct(in_ch, out_ch, ks, padding=[i // 2 for i in ks])
        self._pack = PackedWeights()

    def forward(self, a):
        x, _ = a
        w = self.conv.conv.weight
        y = SingleConvFn.apply(x, w, self._pack.get([w], x.dtype, x.shape[0]), self.ksize, 0, w.shape[1])
        return (y, None)


class BasicBlock(nn.Module):
    """Pre-activation residual block (conv_layers.py:71-94)."""

    def __init__(self, in_ch, out_ch, kernel_size=[3, 3, 3], stride=1):
        super().__init__()
        if stride != 1:
            raise ValueError("strided conv blocks (pool=False) are not supported by the B200 path")
        ks = _triple(kernel_size)
        _check_kernel(ks)
        self.ksize = tuple(ks)
        pad = [i // 2 for i in ks]
        self.conv1 = ConvNormAct(in_ch, out_ch, ks, padding=pad)
        self.conv2 = ConvNormAct(out_ch, out_ch, ks, padding=pad)
        self.shortcut = nn.Sequential()
        if in_ch != out_ch:
            self.shortcut = ConvNormAct(in_ch, out_ch, ks, padding=pad)
        self._pack_f = PackedWeights()
        self._pack_2 = PackedWeights()

    def forward(self, a):
        x, st = a
        if st is None or st.numel() == 0:
            st = ops.instnorm_stats(x, 0, x.shape[-1])
        w1, w2 = self.conv1.conv.weight, self.conv2.conv.weight
        wsc = self.shortcut.conv.weight if isinstance(self.shortcut, ConvNormAct) else None
        fused = [w1] if wsc is None else [w1, wsc]
        packs = (self._pack_f.get(fused, x.dtype, x.shape[0]), self._pack_2.get([w2], x.dtype, x.shape[0]))
        out, out_st = BasicBlockFn.apply(x, st, w1, w2, wsc, packs, self.ksize, 0, w1.shape[1])
        return (out, out_st)


_BLOCKS = {"SingleConv": SingleConv, "BasicBlock": BasicBlock}


def get_block(name):
    """model/dim3/utils.py:7-13 (Bottleneck is not on the BASELINE path)."""
    if name not in _BLOCKS:
        raise ValueError("block %r is not supported by the B200 path (have %s)" % (name, sorted(_BLOCKS)))
    return _BLOCKS[name]


class inconv(nn.Module):
    """unet_utils.py:7-21: raw Conv3d stem, then one block."""

    def __init__(self, in_ch, out_ch, kernel_size=[3, 3, 3], block=BasicBlock):
        super().__init__()
        ks = _triple(kernel_size)
        _check_kernel(ks)
        self.ksize = tuple(ks)
        self.conv1 = nn.Conv3d(in_ch, out_ch, kernel_size=ks, padding=[i // 2 for i in ks], bias=False)
        self.conv2 = block(out_ch, out_ch, kernel_size=ks)
        self._pack = PackedWeights()

    def forward(self, x):
        w = self.conv1.weight
        wf, _ = self._pack.get([w], x.dtype, x.shape[0])
        y, st = StemConvFn.apply(x, w, wf, self.ksize)
        return self.conv2((y, st))


class _Pool(nn.Module):
    """Stands in index 0 of down_block.conv so block indices (…conv.1…, …conv.2…) match the reference."""

    def __init__(self, scale):
        super().__init__()
        self.scale = tuple(int(s) for s in scale)

    def forward(self, a):
        x, _ = a
        y, st = MaxPoolFn.apply(x, self.scale, True)
        return (y, st)


class down_block(nn.Module):
    """unet_utils.py:24-46."""

    def __init__(self, in_ch, out_ch, num_block, block=BasicBlock, kernel_size=[3, 3, 3], down_scale=[2, 2, 2], pool=True):
        super().__init__()
        if not pool:
            raise ValueError("pool=False (strided conv downsampling) is not supported by the B200 path")
        layers = [_Pool(_triple(down_scale)), block(in_ch, out_ch, kernel_size=_triple(kernel_size))]
        for _ in range(num_block - 1):
            layers.append(block(out_ch, out_ch, kernel_size=_triple(kernel_size)))
        self.conv = nn.Sequential(*layers)

    def forward(self, a):
        return self.conv(a)


class up_block(nn.Module):
    """unet_utils.py:48-75: trilinear upsample to the skip's size, cat([skip, up]), blocks."""

    def __init__(self, in_ch, out_ch, num_block, block=BasicBlock, kernel_size=[3, 3, 3], up_scale=[2, 2, 2]):
        super().__init__()
        self.up_scale = _triple(up_scale)
        layers = [block(in_ch + out_ch, out_ch, kernel_size=_triple(kernel_size))]
        for _ in range(num_block - 1):
            layers.append(block(out_ch, out_ch, kernel_size=_triple(kernel_size)))
        self.conv = nn.Sequential(*layers)

    def forward(self, a1, a2):
        low, _ = a1
        skip, skip_st = a2
        cat, cat_st = UpCatFn.apply(low, skip, skip_st, True)
        return self.conv((cat, cat_st))


class UNet(nn.Module):
    """model/dim3/unet.py:12-64.  `norm` must be 'in' (every 3D BASELINE config)."""

    def __init__(self, in_ch, base_ch, scale=[2, 2, 2, 2], kernel_size=[3, 3, 3, 3], num_classes=1,
                 block='BasicBlock', pool=True, norm='in'):
        super().__init__()
        if norm not in ('in', nn.InstanceNorm3d):
            raise ValueError("the B200 path implements InstanceNorm ('in') only, got norm=%r" % (norm,))
        if len(kernel_size) < 5 or len(scale) < 4:
            raise ValueError("kernel_size needs 5 entries and scale 4 (unet.py:35-45)")
        num_block = 2
        blk = get_block(block)
        self.in_ch = in_ch
        self.inc = inconv(in_ch, base_ch, block=blk, kernel_size=kernel_size[0])
        self.down1 = down_block(base_ch, 2 * base_ch, num_block, blk, kernel_size[1], scale[0], pool)
        self.down2 = down_block(2 * base_ch, 4 * base_ch, num_block, blk, kernel_size[2], scale[1], pool)
        self.down3 = down_block(4 * base_ch, 8 * base_ch, num_block, blk, kernel_size[3], scale[2], pool)
        self.down4 = down_block(8 * base_ch, 10 * base_ch, num_block, blk, kernel_size[4], scale[3], pool)
        self.up1 = up_block(10 * base_ch, 8 * base_ch, num_block, blk, kernel_size[3], scale[3])
        self.up2 = up_block(8 * base_ch, 4 * base_ch, num_block, blk, kernel_size[2], scale[2])
        self.up3 = up_block(4 * base_ch, 2 * base_ch, num_block, blk, kernel_size[1], scale[1])
        self.up4 = up_block(2 * base_ch, base_ch, num_block, blk, kernel_size[0], scale[0])
        self.outc = nn.Conv3d(base_ch, num_classes, kernel_size=1)
        self._pack_out = PackedWeights()
        self._packs = ops.PackRegistry(self)

    def forward(self, x):
        if not x.is_cuda:
            raise ops._lib.B200SegError("b200seg.UNet runs on a B200 only — there is no CPU fallback")
        with ops.on_device(x):
            return self._forward(x)

    def _forward(self, x):
        dt = ops.compute_dtype()
        self._packs.refresh()           # every packed weight image is rebuilt from the live parameters (one launch)
        # boundary: NCDHW-shaped input -> NDHWC working layout (free when in_ch == 1)
        xin = x.permute(0, 2, 3, 4, 1)
        xin = xin.to(dt).contiguous()
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
        # boundary: logical NCDHW shape over the NDHWC buffer (channels_last_3d strides, zero-copy)
        return logits.permute(0, 4, 1, 2, 3)
