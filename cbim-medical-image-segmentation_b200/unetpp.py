"""UNet++ (nested U-Net) on the b200seg blocks — drop-in for the reference's ``model/dim3/unetpp.py:8-90`` (the class
``get_model`` builds for ``args.model == 'unet++'``, model/utils.py:85-87): same constructor signature, same module
tree (``conv{i}_{j}`` Sequentials of blocks, ``output``), therefore the same ``state_dict`` keys / shapes / registration
order.  It reuses the kernels of the 3D UNet path unchanged: the dense skip concatenations are built
in place by the upsample+concat kernel, InstanceNorm+ReLU live in the conv loaders."""
import torch
import torch.nn as nn

from . import ops
from .medformer_ops import CatFn
from .ops import MaxPoolFn, OutConvFn, PackedWeights, UpCatFn
from .unet3d import _triple, get_block


def _cat(feats):
    """torch.cat([...], 1) of (tensor, IN-sums) features; the sums of a concatenation are the concatenated sums."""
    x, st = feats[0]
    for y, sy in feats[1:]:
        if st is None or st.numel() == 0:
            st = ops.instnorm_stats(x, 0, x.shape[-1])
        if sy is None or sy.numel() == 0:
            sy = ops.instnorm_stats(y, 0, y.shape[-1])
        x = CatFn.apply(x, y)
        st = torch.cat([st, sy], dim=1).contiguous()
    return x, st


class UNetPlusPlus(nn.Module):
    """model/dim3/unetpp.py:8-90.  `norm` must be 'in' (every 3D config of the reference uses InstanceNorm)."""

    def __init__(self, in_ch, base_ch, scale, kernel_size, num_classes=1, block='SingleConv', norm='bn'):
        super().__init__()
        if norm not in ('in', nn.InstanceNorm3d):
            raise ValueError("the B200 path implements InstanceNorm ('in') only, got norm=%r" % (norm,))
        num_block = 2
        blk = get_block(block)
        n = [base_ch, base_ch * 2, base_ch * 4, base_ch * 8, base_ch * 10]
        self.scales = [tuple(int(v) for v in _triple(s)) for s in scale[:4]]
        k = kernel_size
        self.conv0_0 = self.make_layer(in_ch, n[0], num_block, blk, k[0])
        self.conv1_0 = self.make_layer(n[0], n[1], num_block, blk, k[1])
        self.conv2_0 = self.make_layer(n[1], n[2], num_block, blk, k[2])
        self.conv3_0 = self.make_layer(n[2], n[3], num_block, blk, k[3])
        self.conv4_0 = self.make_layer(n[3], n[4], num_block, blk, k[4])
        self.conv0_1 = self.make_layer(n[0] + n[1], n[0], num_block, blk, k[0])
        self.conv1_1 = self.make_layer(n[1] + n[2], n[1], num_block, blk, k[1])
        self.conv2_1 = self.make_layer(n[2] + n[3], n[2], num_block, blk, k[2])
        self.conv3_1 = self.make_layer(n[3] + n[4], n[3], num_block, blk, k[3])
        self.conv0_2 = self.make_layer(n[0] * 2 + n[1], n[0], num_block, blk, k[0])
        self.conv1_2 = self.make_layer(n[1] * 2 + n[2], n[1], num_block, blk, k[1])
        self.conv2_2 = self.make_layer(n[2] * 2 + n[3], n[2], num_block, blk, k[2])
        self.conv0_3 = self.make_layer(n[0] * 3 + n[1], n[0], num_block, blk, k[0])
        self.conv1_3 = self.make_layer(n[1] * 3 + n[2], n[1], num_block, blk, k[1])
        self.conv0_4 = self.make_layer(n[0] * 4 + n[1], n[0], num_block, blk, k[0])
        self.output = nn.Conv3d(n[0], num_classes, kernel_size=1)
        self._pack_out = PackedWeights()
        self._packs = ops.PackRegistry(self)

    @staticmethod
    def make_layer(in_ch, out_ch, num_block, block, kernel_size):
        blocks = [block(in_ch, out_ch, kernel_size=_triple(kernel_size))]
        for _ in range(num_block - 1):
            blocks.append(block(out_ch, out_ch, kernel_size=_triple(kernel_size)))
        return nn.Sequential(*blocks)

    def _pool(self, f, i):
        y, st = MaxPoolFn.apply(f[0], self.scales[i], True)
        return (y, st)

    def _up_cat(self, skips, low):
        """cat([*skips, upsample(low)], 1): nn.Upsample(scale_factor, trilinear, align_corners=True) + torch.cat
        (unetpp.py:17-24,58-76) written straight into one buffer."""
        sx, sst = _cat(skips)
        return UpCatFn.apply(low[0], sx, sst, True)

    def forward(self, x):
        if not x.is_cuda:
            raise ops._lib.B200SegError("b200seg.UNetPlusPlus runs on a B200 only — there is no CPU fallback")
        with ops.on_device(x):
            return self._forward(x)

    def _forward(self, x):
        dt = ops.compute_dtype()
        self._packs.refresh()
        xin = x.permute(0, 2, 3, 4, 1).to(dt).contiguous()
        x0_0 = self.conv0_0((xin, None))
        x1_0 = self.conv1_0(self._pool(x0_0, 0))
        x0_1 = self.conv0_1(self._up_cat([x0_0], x1_0))
        x2_0 = self.conv2_0(self._pool(x1_0, 1))
        x1_1 = self.conv1_1(self._up_cat([x1_0], x2_0))
        x0_2 = self.conv0_2(self._up_cat([x0_0, x0_1], x1_1))
        x3_0 = self.conv3_0(self._pool(x2_0, 2))
        x2_1 = self.conv2_1(self._up_cat([x2_0], x3_0))
        x1_2 = self.conv1_2(self._up_cat([x1_0, x1_1], x2_1))
        x0_3 = self.conv0_3(self._up_cat([x0_0, x0_1, x0_2], x1_2))
        x4_0 = self.conv4_0(self._pool(x3_0, 3))
        x3_1 = self.conv3_1(self._up_cat([x3_0], x4_0))
        x2_2 = self.conv2_2(self._up_cat([x2_0, x2_1], x3_1))
        x1_3 = self.conv1_3(self._up_cat([x1_0, x1_1, x1_2], x2_2))
        x0_4 = self.conv0_4(self._up_cat([x0_0, x0_1, x0_2, x0_3], x1_3))
        w, b = self.output.weight, self.output.bias
        logits = OutConvFn.apply(x0_4[0], w, b, self._pack_out.get([w], dt, x.shape[0]))
        return logits.permute(0, 4, 1, 2, 3)
