"""Pin oracle/swin_unetr.py (whole-model functional restatement) against the reference's VENDORED SwinUNETR class and
write tests/golden/swin_unetr_*.pt.

model/dim3/swin_unetr.py is imported UNMODIFIED; the seven `monai` symbols it pulls in (monai 1.1.0 is not in this
image and its source is not under /root/reference) are provided by working stand-in modules written from MONAI
1.1.0's published semantics (attribute names cross-checked against `load_from`, swin_unetr.py:230-277,629-643).
So this run pins everything the reference file itself defines — SwinUNETR.forward's wiring, SwinTransformer /
BasicLayer / SwinTransformerBlock / WindowAttention / PatchMerging / compute_mask — bit-for-bit, while the monai
blocks stay "parity unpinned" (the stand-in and the oracle are two statements of the same published semantics).
Runs only where /root/reference exists.  Usage:  python oracle/make_golden_swin_unetr.py
"""
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import digest, import_reference     # noqa: E402
from oracle import losses as olosses                         # noqa: E402
from oracle import swin_unetr as osw                         # noqa: E402
from oracle import unet3d as ounet                           # noqa: E402
from oracle.synth import make_volume                         # noqa: E402


def install_monai_standin():
    class Convolution(nn.Module):            # monai.networks.blocks.Convolution(conv_only=True): child `conv`
        def __init__(self, ci, co, k, stride=1, transposed=False, bias=False):
            super().__init__()
            if transposed:
                self.conv = nn.ConvTranspose3d(ci, co, kernel_size=k, stride=stride, bias=bias)
            else:
                self.conv = nn.Conv3d(ci, co, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=bias)

        def forward(self, x):
            return self.conv(x)

    class UnetResBlock(nn.Module):
        def __init__(self, ci, co):
            super().__init__()
            self.conv1 = Convolution(ci, co, 3)
            self.conv2 = Convolution(co, co, 3)
            self.lrelu = nn.LeakyReLU(0.01, inplace=True)
            self.norm1 = nn.InstanceNorm3d(co)
            self.norm2 = nn.InstanceNorm3d(co)
            if ci != co:
                self.conv3 = Convolution(ci, co, 1)
                self.norm3 = nn.InstanceNorm3d(co)

        def forward(self, inp):
            residual = inp
            out = self.lrelu(self.norm1(self.conv1(inp)))
            out = self.norm2(self.conv2(out))
            if hasattr(self, "conv3"):
                residual = self.norm3(self.conv3(residual))
            out += residual
            return self.lrelu(out)

    class UnetrBasicBlock(nn.Module):
        def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, res_block=False):
            super().__init__()
            assert res_block and kernel_size == 3 and stride == 1 and spatial_dims == 3
            self.layer = UnetResBlock(in_channels, out_channels)

        def forward(self, x):
            return self.layer(x)

    class UnetrUpBlock(nn.Module):
        def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size, norm_name, res_block=False):
            super().__init__()
            assert res_block and upsample_kernel_size == 2
            self.transp_conv = Convolution(in_channels, out_channels, 2, stride=2, transposed=True)
            self.conv_block = UnetResBlock(out_channels + out_channels, out_channels)

        def forward(self, inp, skip):
            return self.conv_block(torch.cat((self.transp_conv(inp), skip), dim=1))

    class UnetOutBlock(nn.Module):
        def __init__(self, spatial_dims, in_channels, out_channels):
            super().__init__()
            self.conv = Convolution(in_channels, out_channels, 1, bias=True)

        def forward(self, x):
            return self.conv(x)

    class PatchEmbed(nn.Module):
        def __init__(self, patch_size, in_chans, embed_dim, norm_layer=None, spatial_dims=3):
            super().__init__()
            assert norm_layer is None
            self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

        def forward(self, x):
            return self.proj(x)

    class MLPBlock(nn.Module):
        def __init__(self, hidden_size, mlp_dim, act="GELU", dropout_rate=0.0, dropout_mode="swin"):
            super().__init__()
            self.linear1, self.linear2 = nn.Linear(hidden_size, mlp_dim), nn.Linear(mlp_dim, hidden_size)

        def forward(self, x):
            return self.linear2(F.gelu(self.linear1(x)))

    def optional_import(module, name=""):
        import importlib
        try:
            m = importlib.import_module(module)
            return (getattr(m, name) if name else m), True
        except Exception:
            return None, False

    mods = {n: types.ModuleType(n) for n in ("monai", "monai.networks", "monai.networks.blocks", "monai.networks.layers", "monai.utils")}
    b, l, u = mods["monai.networks.blocks"], mods["monai.networks.layers"], mods["monai.utils"]
    b.MLPBlock, b.PatchEmbed, b.UnetOutBlock, b.UnetrBasicBlock, b.UnetrUpBlock = MLPBlock, PatchEmbed, UnetOutBlock, UnetrBasicBlock, UnetrUpBlock
    l.DropPath, l.trunc_normal_ = nn.Identity, torch.nn.init.trunc_normal_
    u.ensure_tuple_rep = lambda v, n: tuple(v) if isinstance(v, (list, tuple)) else (v,) * n
    u.look_up_option = lambda k, opts, default=None: opts[k] if isinstance(opts, dict) else k
    u.optional_import = optional_import
    sys.modules.update(mods)


CASES = {  # name: (img size, in_ch, classes, feature_size, ce weight, seeds)
    "swin_unetr_small": ((64, 64, 64), 1, 3, 24, [0.5, 1.0, 2.0], (71, 72)),
}


def main():
    torch.set_num_threads(8)
    import_reference()
    install_monai_standin()
    from model.dim3 import swin_unetr as ref
    out = os.path.join(ROOT, "tests", "golden")
    for name, (size, in_ch, classes, fs, w, (sseed, dseed)) in CASES.items():
        net = ref.SwinUNETR(size, in_ch, classes, feature_size=fs)
        keys = [k for k in net.state_dict() if not k.endswith("relative_position_index")]
        shapes = osw.swin_unetr_param_shapes(in_ch, classes, fs)
        assert keys == list(shapes), [k for k in keys if k not in shapes][:5] + [k for k in shapes if k not in keys][:5]
        for k in keys:
            assert tuple(net.state_dict()[k].shape) == shapes[k], k
        sd = ounet.make_state_dict(shapes, seed=sseed)
        for k in sd:      # LayerNorm weights around 1, rel-pos tables with some spread
            if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k.endswith("norm.weight"):
                sd[k] = 1.0 + 0.1 * sd[k] / sd[k].abs().max()
            if k.endswith("relative_position_bias_table"):
                sd[k] = sd[k] * 3.0
        missing = net.load_state_dict(sd, strict=False)
        assert all(k.endswith("relative_position_index") for k in missing.missing_keys) and not missing.unexpected_keys
        img, lab = make_volume(1, *size, classes, seed=dseed, in_ch=in_ch)
        weight = torch.tensor(w)
        logits = net(img)
        loss = nn.CrossEntropyLoss(weight=weight)(logits, lab.squeeze(1)) + olosses.dice_loss(logits, lab)
        loss.backward()
        ref_grads = {k: p.grad.clone() for k, p in net.named_parameters()}
        so = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lo = osw.swin_unetr_forward(so, img)
        lo_loss = olosses.total_loss(lo, lab, weight)
        lo_loss.backward()
        e = (lo - logits).abs().max().item() / logits.abs().max().item()
        print(name, "logits rel diff oracle vs reference class: %.2e, loss %.6f vs %.6f" % (e, lo_loss.item(), loss.item()))
        assert e < 1e-5 and abs(lo_loss.item() - loss.item()) < 1e-5
        worst = max(((so[k].grad - ref_grads[k]).abs().max() / (ref_grads[k].abs().max() + 1e-30)).item() for k in ref_grads)
        print(name, "worst grad rel diff %.2e over %d tensors (%d params)" % (worst, len(ref_grads), sum(v.numel() for v in sd.values())))
        assert worst < 2e-3
        torch.save({"cfg": dict(size=size, in_ch=in_ch, classes=classes, feature_size=fs, ce_weight=w, state_seed=sseed, data_seed=dseed),
                    "shapes": shapes, "logits": logits.detach().half(), "argmax": logits.argmax(1).to(torch.uint8), "loss": loss.item(),
                    "grad_digest": {k: digest(v) for k, v in ref_grads.items()}}, os.path.join(out, name + ".pt"))


if __name__ == "__main__":
    main()
