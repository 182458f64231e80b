"""Pin oracle/swin_ops.py against the reference's VENDORED Swin code and write tests/golden/swin_*.pt.

model/dim3/swin_unetr.py imports seven symbols from `monai` (not in this image, source not under /root/reference).
None of them is exercised by the classes pinned here (WindowAttention, SwinTransformerBlock.forward_part1,
PatchMerging, compute_mask, window_partition/reverse), so a throw-away stand-in module is registered only to let the
file import; the stand-ins raise if anything actually calls into them, except `trunc_normal_` (torch's own) and the
`MLPBlock` placeholder that SwinTransformerBlock.__init__ constructs but forward_part1 never runs.
Runs only where /root/reference exists.  Usage:  python oracle/make_golden_swin.py
"""
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference      # noqa: E402
from oracle import swin_ops as so                    # noqa: E402


def _stub_monai():
    class _Refuse(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            raise RuntimeError("monai stand-in called: this part of SwinUNETR is not pinned")

    class _Mlp(_Refuse):      # constructed by SwinTransformerBlock.__init__, used only by forward_part2
        def __init__(self, hidden_size, mlp_dim, act="GELU", dropout_rate=0.0, dropout_mode="swin"):
            super().__init__()
            self.linear1, self.linear2 = nn.Linear(hidden_size, mlp_dim), nn.Linear(mlp_dim, hidden_size)

    def optional_import(module, name=""):
        import importlib
        try:
            m = importlib.import_module(module)
            return (getattr(m, name) if name else m), True
        except Exception:
            return None, False

    mods = {n: types.ModuleType(n) for n in ("monai", "monai.networks", "monai.networks.blocks", "monai.networks.layers", "monai.utils")}
    b, l, u = mods["monai.networks.blocks"], mods["monai.networks.layers"], mods["monai.utils"]
    b.MLPBlock, b.PatchEmbed, b.UnetOutBlock, b.UnetrBasicBlock, b.UnetrUpBlock = _Mlp, _Refuse, _Refuse, _Refuse, _Refuse
    l.DropPath, l.trunc_normal_ = _Refuse, torch.nn.init.trunc_normal_
    u.ensure_tuple_rep = lambda v, n: tuple(v) if isinstance(v, (list, tuple)) else (v,) * n
    u.look_up_option = lambda k, opts, default=None: opts[k] if isinstance(opts, dict) else k
    u.optional_import = optional_import
    sys.modules.update(mods)


ATTN = {  # name: (dim, heads, window, n windows per sample, batch, masked, seed)
    "swin_attn_a": (48, 3, (7, 7, 7), 2, 2, True, 41),      # the SwinUNETR stage-1 shape: 343 tokens, d_h 16
    "swin_attn_b": (24, 3, (2, 3, 4), 3, 1, False, 42),
    "swin_attn_c": (96, 6, (4, 4, 4), 4, 1, True, 43),
}
BLOCK = {  # name: (dim, heads, window, shift, input dhw, batch, seed)
    "swin_block_a": (24, 3, (4, 4, 4), (2, 2, 2), (8, 9, 10), 1, 51),      # padding + shift + mask
    "swin_block_b": (24, 3, (4, 4, 4), (0, 0, 0), (8, 8, 8), 2, 52),       # aligned, unshifted
    "swin_block_c": (48, 6, (7, 7, 7), (3, 3, 3), (4, 9, 15), 1, 53),      # window clamped on a short axis
}


def main():
    import_reference()
    _stub_monai()
    from model.dim3 import swin_unetr as ref
    out = os.path.join(ROOT, "tests", "golden")

    for name, (dim, heads, ws, nw, B, masked, seed) in ATTN.items():
        torch.manual_seed(seed)
        mod = ref.WindowAttention(dim, heads, ws, qkv_bias=True)
        with torch.no_grad():
            mod.relative_position_bias_table.normal_(0, 0.5)
        n = ws[0] * ws[1] * ws[2]
        x = torch.randn(B * nw, n, dim, requires_grad=True)
        mask = None
        if masked:      # a real shifted-window mask: regions of a padded (2w)^3 volume, first nw windows
            full = ref.compute_mask([2 * ws[0], 2 * ws[1], 2 * ws[2]], ws, tuple(i // 2 for i in ws), "cpu")
            mask = full[-nw:].clone()
            assert torch.equal(full, so.compute_mask([2 * ws[0], 2 * ws[1], 2 * ws[2]], ws, tuple(i // 2 for i in ws)))
        y = mod(x, mask)
        gy = torch.randn_like(y)
        y.backward(gy)
        assert torch.equal(mod.relative_position_index, so.relative_position_index(ws))
        p = {"qkv_w": mod.qkv.weight, "qkv_b": mod.qkv.bias, "proj_w": mod.proj.weight, "proj_b": mod.proj.bias,
             "bias_table": mod.relative_position_bias_table}
        po = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        xo = x.detach().clone().requires_grad_(True)
        yo = so.window_attention(xo, po, heads, so.relative_position_index(ws), mask)
        yo.backward(gy)
        assert torch.allclose(yo, y, atol=1e-6, rtol=1e-5) and torch.allclose(xo.grad, x.grad, atol=1e-6, rtol=1e-4)
        for k in p:
            assert torch.allclose(po[k].grad, p[k].grad, atol=1e-5, rtol=1e-4), k
        # the mask only takes the values 0 / -100: store it as a boolean "crosses a region border" map
        torch.save({"cfg": dict(dim=dim, heads=heads, window=ws), "x": x.detach(), "gy": gy,
                    "mask_cross": None if mask is None else (mask != 0),
                    "params": {k: v.detach().clone() for k, v in p.items()}, "y": y.detach(), "dx": x.grad.clone(),
                    "dparams": {k: v.grad.clone() for k, v in p.items()}}, os.path.join(out, name + ".pt"))
        print(name, "ok", tuple(x.shape))

    for name, (dim, heads, ws, ss, dhw, B, seed) in BLOCK.items():
        torch.manual_seed(seed)
        blk = ref.SwinTransformerBlock(dim, heads, ws, ss, qkv_bias=True)
        with torch.no_grad():
            blk.attn.relative_position_bias_table.normal_(0, 0.5)
            blk.norm1.weight.normal_(1, 0.1)
            blk.norm1.bias.normal_(0, 0.1)
        x = torch.randn(B, *dhw, dim, requires_grad=True)
        uws, uss = ref.get_window_size(dhw, ws, ss)
        assert (uws, uss) == so.get_window_size(dhw, ws, ss)
        pdims = [-(-dhw[i] // uws[i]) * uws[i] for i in range(3)]
        mask = ref.compute_mask(pdims, uws, uss, "cpu") if any(uss) else None
        y = blk.forward_part1(x, mask)
        gy = torch.randn_like(y)
        y.backward(gy)
        p = {"norm1_w": blk.norm1.weight, "norm1_b": blk.norm1.bias, "qkv_w": blk.attn.qkv.weight, "qkv_b": blk.attn.qkv.bias,
             "proj_w": blk.attn.proj.weight, "proj_b": blk.attn.proj.bias, "bias_table": blk.attn.relative_position_bias_table}
        po = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        xo = x.detach().clone().requires_grad_(True)
        mo = so.compute_mask(pdims, uws, uss) if any(uss) else None
        yo = so.swin_block_part1(xo, po, heads, ws, ss, mo)
        yo.backward(gy)
        assert torch.allclose(yo, y, atol=1e-6, rtol=1e-5) and torch.allclose(xo.grad, x.grad, atol=1e-5, rtol=1e-4)
        for k in p:
            assert torch.allclose(po[k].grad, p[k].grad, atol=1e-5, rtol=1e-4), k
        torch.save({"cfg": dict(dim=dim, heads=heads, window=ws, shift=ss, dhw=dhw), "x": x.detach(), "gy": gy,
                    "params": {k: v.detach().clone() for k, v in p.items()}, "y": y.detach(), "dx": x.grad.clone(),
                    "dparams": {k: v.grad.clone() for k, v in p.items()}}, os.path.join(out, name + ".pt"))
        print(name, "ok", tuple(x.shape), "window", uws, "shift", uss)

    torch.manual_seed(61)
    for name, cls, fn in (("swin_merge_a", ref.PatchMerging, so.patch_merging), ("swin_merge_b", ref.PatchMergingV2, so.patch_merging_v2)):
        mod = cls(16)
        with torch.no_grad():
            mod.norm.weight.normal_(1, 0.1)
            mod.norm.bias.normal_(0, 0.1)
        x = torch.randn(1, 5, 6, 7, 16, requires_grad=True)
        y = mod(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        xo = x.detach().clone().requires_grad_(True)
        yo = fn(xo, mod.norm.weight.detach(), mod.norm.bias.detach(), mod.reduction.weight.detach())
        yo.backward(gy)
        assert torch.allclose(yo, y, atol=1e-6) and torch.allclose(xo.grad, x.grad, atol=1e-6)
        torch.save({"x": x.detach(), "gy": gy, "norm_w": mod.norm.weight.detach().clone(), "norm_b": mod.norm.bias.detach().clone(),
                    "red_w": mod.reduction.weight.detach().clone(), "y": y.detach(), "dx": x.grad.clone()},
                   os.path.join(out, name + ".pt"))
        print(name, "ok")


if __name__ == "__main__":
    main()
