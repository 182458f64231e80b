"""Pin oracle/medformer_ops.py against the REAL reference modules and write tests/golden/{biattn,dwconv}_*.pt.

Runs only where /root/reference exists.  BidirectionAttention (medformer_utils.py:11-97) is instantiated
unmodified with proj_type='linear'; forward hooks capture what its projections produce / consume, i.e. exactly the
inputs and outputs of the fused core the CUDA kernel replaces; autograd through the reference module gives the
gradients at the same boundary.  Usage:  python oracle/make_golden_medformer.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference          # noqa: E402
from oracle import medformer_ops as mops                 # noqa: E402

BIATTN = {  # name: (B, heads, feat spatial, map spatial, feat_dim, map_dim, seed)
    "biattn_d": (1, 2, (3, 5, 9), (4, 4, 4), 16, 24, 24),
    "biattn_a": (2, 2, (4, 6, 5), (3, 3, 3), 24, 16, 21),
    "biattn_b": (1, 4, (5, 6, 7), (3, 3, 3), 32, 32, 22),
    "biattn_c": (1, 1, (2, 3, 67), (2, 2, 2), 8, 8, 23),
}
MODELS = {   # name: (ctor kwargs, classes, input BxDxHxW)
    # the BCV architecture (config/bcv/medformer_3d.yaml:9-28) on a small crop
    "medformer_bcv": (dict(map_size=[3, 3, 3], conv_num=[2, 0, 0, 0, 0, 0, 2, 2], trans_num=[0, 2, 4, 6, 4, 2, 0, 0],
                           num_heads=[1, 4, 8, 10, 8, 4, 1, 1], fusion_depth=2, fusion_dim=320, fusion_heads=10,
                           kernel_size=[[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
                           scale=[[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]], aux_loss=True), 14, (1, 16, 32, 32)),
    # a variant: conv blocks on attention levels, 64 map tokens (the 4x4x4 maps of the AMOS / KiTS YAMLs -> 192 fused
    # tokens), isotropic scales, no aux head, batch 2
    "medformer_var": (dict(map_size=[4, 4, 4], conv_num=[1, 1, 0, 0, 0, 1, 1, 1], trans_num=[0, 1, 2, 1, 1, 1, 0, 0],
                           num_heads=[1, 4, 8, 10, 8, 4, 1, 1], fusion_depth=1, fusion_dim=128, fusion_heads=4,
                           kernel_size=[[3, 3, 3]] * 5, scale=[[2, 2, 2]] * 4, aux_loss=False), 4, (2, 16, 32, 32)),
    # the ACDC YAML (config/acdc/medformer_3d.yaml:9-28): 2x6x6 = 72 map tokens, 4 heads per level -> dim_head 32/64/80,
    # fusion_dim 256.  Oracle-only for now: the B200 kernels do not cover dim_head != 32 / 72 tokens yet.
    "medformer_acdc": (dict(map_size=[2, 6, 6], conv_num=[2, 0, 0, 0, 0, 0, 2, 2], trans_num=[0, 2, 2, 2, 2, 2, 0, 0],
                            num_heads=[1, 4, 4, 4, 4, 4, 1, 1], fusion_depth=2, fusion_dim=256, fusion_heads=4,
                            kernel_size=[[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
                            scale=[[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]], aux_loss=True), 4, (1, 8, 32, 32)),
}
DWCONV = {"dwconv_a": (2, 16, (5, 6, 7), [3, 3, 3], 31), "dwconv_b": (1, 24, (4, 9, 8), [1, 3, 3], 32)}


def main():
    import_reference()
    from model.dim3.medformer_utils import BidirectionAttention
    from model.dim3.conv_layers import DepthwiseSeparableConv
    out = os.path.join(ROOT, "tests", "golden")
    for name, (B, heads, fs, ms, fd, md, seed) in BIATTN.items():
        torch.manual_seed(seed)
        mod = BidirectionAttention(fd, md, fd, heads=heads, dim_head=32, map_size=list(ms), proj_type="linear")
        cap = {}
        def grab_out(key):
            def hook(m, i, o):
                o.retain_grad()
                cap[key] = o
            return hook

        def grab_in(key):
            def hook(m, i):
                i[0].retain_grad()
                cap[key] = i[0]
            return hook

        mod.feat_qv.register_forward_hook(grab_out("fqv"))
        mod.map_qv.register_forward_hook(grab_out("mqv"))
        mod.feat_out.register_forward_pre_hook(grab_in("fo"))
        mod.map_out.register_forward_pre_hook(grab_in("mo"))
        feat = torch.randn(B, fd, *fs) * 2.0
        smap = torch.randn(B, md, *ms) * 2.0
        y_f, y_m = mod(feat, smap)
        (y_f.square().sum() + y_m.square().sum() * 5).backward()
        fq, fv = cap["fqv"].detach().chunk(2, dim=1)
        mq, mv = cap["mqv"].detach().chunk(2, dim=1)
        # the oracle must reproduce the reference at this boundary
        qf, vf, qm, vm = (t.clone().requires_grad_(True) for t in (fq, fv, mq, mv))
        fo, mo = mops.bidirection_attention_core(qf, vf, qm, vm, heads)
        assert torch.allclose(fo, cap["fo"], atol=1e-6, rtol=1e-5) and torch.allclose(mo, cap["mo"], atol=1e-6, rtol=1e-5)
        torch.autograd.backward([fo, mo], [cap["fo"].grad, cap["mo"].grad])
        g_ref_f, g_ref_m = cap["fqv"].grad, cap["mqv"].grad
        assert torch.allclose(torch.cat([qf.grad, vf.grad], 1), g_ref_f, atol=1e-5, rtol=1e-4)
        assert torch.allclose(torch.cat([qm.grad, vm.grad], 1), g_ref_m, atol=1e-5, rtol=1e-4)
        torch.save({"heads": heads, "fqv": cap["fqv"].detach(), "mqv": cap["mqv"].detach(),
                    "fo": cap["fo"].detach(), "mo": cap["mo"].detach(), "dfo": cap["fo"].grad, "dmo": cap["mo"].grad,
                    "dfqv": g_ref_f, "dmqv": g_ref_m}, os.path.join(out, name + ".pt"))
        print(name, "ok", tuple(cap["fqv"].shape))
    for name, (B, C, sp, k, seed) in DWCONV.items():
        torch.manual_seed(seed)
        mod = DepthwiseSeparableConv(C, C, kernel_size=k, bias=False).depthwise
        x = torch.randn(B, C, *sp, requires_grad=True)
        y = mod(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        xo = x.detach().clone().requires_grad_(True)
        wo = mod.weight.detach().clone().requires_grad_(True)
        yo = mops.depthwise_conv3d(xo, wo)
        yo.backward(gy)
        assert torch.equal(yo, y) and torch.allclose(xo.grad, x.grad, atol=1e-6) and torch.allclose(wo.grad, mod.weight.grad, atol=1e-5)
        torch.save({"x": x.detach(), "w": mod.weight.detach(), "y": y.detach(), "gy": gy, "dx": x.grad, "dw": mod.weight.grad},
                   os.path.join(out, name + ".pt"))
        print(name, "ok")


def models():
    """Whole-model fixtures: the unmodified reference MedFormer, a seeded state_dict, forward + CE + Dice + backward
    on CPU fp32; oracle/medformer.py must reproduce it."""
    from model.dim3.medformer import MedFormer
    from training.losses import DiceLoss
    from oracle import losses as olosses
    from oracle import medformer as omed
    from oracle.make_golden import digest
    from oracle.synth import make_volume
    from oracle.unet3d import make_state_dict
    torch.set_num_threads(8)
    for name, (kw, classes, (B, D, H, W)) in MODELS.items():
        ref = MedFormer(1, classes, 32, conv_block="BasicBlock", expansion=4, attn_drop=0, proj_drop=0,
                        proj_type="depthwise", norm="in", act="relu", **kw)
        shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        sd = make_state_dict(shapes, seed=11)
        for k in sd:                       # LayerNorm affine: ones/zeros + noise instead of the fan-in draw
            if k.endswith("norm.weight"):
                sd[k] = 1.0 + 0.1 * sd[k] / sd[k].abs().max()
        ref.load_state_dict(sd)
        ref.train()
        img, lab = make_volume(B, D, H, W, classes, seed=2023)
        w = torch.tensor([0.5] + [1.0] * (classes - 1))
        aux_w = [0.5, 0.5]
        res = ref(img)
        outs = res if isinstance(res, list) else [res]
        loss = sum((aux_w[j] if len(outs) > 1 else 1.0) *
                   (torch.nn.CrossEntropyLoss(weight=w)(r, lab.squeeze(1)) + DiceLoss()(r, lab)) for j, r in enumerate(outs))
        loss.backward()
        grads = {k: p.grad.clone() for k, p in ref.named_parameters()}
        sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        ro = omed.medformer_forward(sdo, img, kw)
        lo = olosses.total_loss(ro, lab, w, aux_w) if isinstance(ro, list) else olosses.total_loss(ro, lab, w)
        lo.backward()
        ro = ro if isinstance(ro, list) else [ro]
        e_logit = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(ro, outs))
        # a few parameters have an analytically zero gradient (e.g. the last fusion bias: a per-channel constant
        # over the map tokens is removed again by norm2); compare those on the scale of the typical gradient
        gfloor = 1e-4 * max(g.abs().max().item() for g in grads.values())
        e_grad = max(((sdo[k].grad - grads[k]).abs().max() / (grads[k].abs().max() + gfloor)).item() for k in grads)
        print("%-14s oracle-vs-reference: logits %.2e loss %.2e grads %.2e" % (name, e_logit, abs(lo.item() - loss.item()), e_grad))
        assert e_logit < 1e-4 and abs(lo.item() - loss.item()) < 1e-5 and e_grad < 1e-3
        small = [k for k in grads if grads[k].numel() <= 4096]
        torch.save({"cfg": dict(kw, classes=classes, shape=(B, D, H, W), ce_weight=w.tolist(), aux_weight=aux_w,
                                state_seed=11, data_seed=2023),
                    "shapes": shapes, "logits": [r.detach().half() for r in outs],
                    "argmax": [r.argmax(1).to(torch.uint8) for r in outs], "loss": loss.item(),
                    "grad_digest": {k: digest(g) for k, g in grads.items()},
                    "grad_small": {k: grads[k].clone() for k in small[:40]}},
                   os.path.join(ROOT, "tests", "golden", name + ".pt"))


if __name__ == "__main__":
    main()
    models()
