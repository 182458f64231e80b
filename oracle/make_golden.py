"""Pin the oracle against the REAL reference and write tests/golden/*.pt.

Runs only where /root/reference exists (the build container).  For every case it
  1. imports the unmodified reference modules (package-shell bypass of SURVEY.md §8c: model/dim3/__init__.py
     eagerly imports monai/timm-dependent files, so an empty package object is registered first),
  2. loads a deterministic state_dict, runs forward + CE + Dice + backward on CPU fp32,
  3. asserts the oracle restatement (oracle/unet3d.py, oracle/losses.py) reproduces the reference to
     ~fp32 round-off,
  4. stores inputs-by-seed, outputs and gradient digests as the committed fixture.
Usage:  python oracle/make_golden.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B200SEG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import losses as olosses          # noqa: E402
from oracle import unet3d as ounet            # noqa: E402
from oracle.synth import make_volume          # noqa: E402

CASES = {
    # name: (block, base, classes, scale, kernel, input BxDxHxW, ce_weight)
    "resunet_iso": ("BasicBlock", 8, 4, [[2, 2, 2]] * 4, [[3, 3, 3]] * 5, (1, 32, 32, 32), [0.5, 1, 1, 1]),
    "resunet_acdc": ("BasicBlock", 8, 4, [[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]],
                     [[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]], (2, 8, 64, 64), [0.5, 1, 1, 1]),
    "unet_single": ("SingleConv", 8, 3, [[2, 2, 2]] * 4, [[3, 3, 3]] * 5, (1, 32, 32, 32), [0.5, 1, 2]),
}
LOSS_CASES = {"loss_a": (2, 5, (6, 7, 8), 11), "loss_b": (1, 14, (8, 8, 8), 12), "loss_c": (3, 3, (4, 5, 6), 13)}


def import_reference():
    sys.path.insert(0, REF)
    for pkg, sub in (("model", "model"), ("model.dim3", "model/dim3")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[pkg] = m
    from model.dim3.unet import UNet          # noqa
    from training.losses import DiceLoss      # noqa
    return UNet, DiceLoss


def digest(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, min(t.numel(), 64)).long()
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "sq": (t * t).sum().item(), "sample": t[idx].float()}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    UNet, DiceLoss = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (block, base, classes, scale, kernel, shp, w) in CASES.items():
        B, D, H, W = shp
        ref = UNet(1, base, scale=scale, kernel_size=kernel, num_classes=classes, block=block, norm="in")
        shapes = ounet.unet_param_shapes(1, base, classes, kernel, block)
        assert list(shapes) == list(ref.state_dict()), "oracle key order != reference registration order"
        assert all(tuple(v.shape) == tuple(shapes[k]) for k, v in ref.state_dict().items())
        sd = ounet.make_state_dict(shapes, seed=7)
        ref.load_state_dict(sd)
        weight = torch.tensor(w, dtype=torch.float32)
        ref.train()
        # ReLU'(0) makes the gradient discontinuous: a pre-activation within fp32 rounding of 0 flips its mask
        # between two equally valid fp32 evaluations, and through InstanceNorm's whole-channel reductions a flip
        # moves gradients far more than 1e-3.  The reference's own fp32 gradients therefore differ from an fp64
        # evaluation of the same network by ~1e-3..1e-1 (measured here and stored as `ref_fp32_vs_fp64_grad_err`);
        # the parity tests use that number as the noise floor of "matches the reference's backward".
        for data_seed in range(2023, 2024):
            img, lab = make_volume(B, D, H, W, classes, seed=data_seed)
            ref.zero_grad(set_to_none=True)
            logits = ref(img)
            ce = torch.nn.CrossEntropyLoss(weight=weight)(logits, lab.squeeze(1))
            dl = DiceLoss()(logits, lab)
            loss = ce + dl
            loss.backward()
            grads = {k: p.grad.clone() for k, p in ref.named_parameters()}
            sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
            l64 = ounet.unet_forward(sd64, img.double(), scale, kernel, block)
            olosses.total_loss(l64, lab, weight.double()).backward()
            e64 = max(((sd64[k].grad - grads[k].double()).abs().max() / (sd64[k].grad.abs().max() + 1e-300)).item() for k in grads)
            print("   seed %d: reference fp32 vs fp64 evaluation, worst grad rel err %.2e" % (data_seed, e64))
            break   # first seed; the discrepancy is recorded in the fixture and used as the noise floor by the tests
        # --- oracle restatement (fp32, same ops) must reproduce the reference
        sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lo = ounet.unet_forward(sdo, img, scale, kernel, block)
        losso = olosses.total_loss(lo, lab, weight)
        losso.backward()
        err_logit = (lo - logits).abs().max().item() / logits.abs().max().item()
        err_loss = abs(losso.item() - loss.item())
        err_grad = max(((sdo[k].grad - grads[k]).abs().max() / (grads[k].abs().max() + 1e-30)).item() for k in grads)
        print("%-14s oracle-vs-reference: logits %.2e  loss %.2e  grads %.2e" % (name, err_logit, err_loss, err_grad))
        assert err_logit < 1e-5 and err_loss < 1e-5 and err_grad < 1e-4, "oracle restatement diverges from the reference"
        small = ["inc.conv1.weight", "outc.weight", "outc.bias"]
        torch.save({
            "cfg": {"block": block, "base": base, "classes": classes, "scale": scale, "kernel": kernel,
                    "shape": shp, "ce_weight": w, "state_seed": 7, "data_seed": data_seed,
                    "ref_fp32_vs_fp64_grad_err": e64},
            "keys": list(shapes), "logits": logits.detach().half(), "argmax": logits.argmax(1).to(torch.uint8),
            "loss": loss.item(), "ce": ce.item(), "dice": dl.item(),
            "grad_digest": {k: digest(g) for k, g in grads.items()},
            "grad_small": {k: grads[k].clone() for k in small},
            "logits_digest": digest(logits),
        }, os.path.join(ROOT, "tests", "golden", name + ".pt"))
    for name, (B, C, sp, seed) in LOSS_CASES.items():
        g = torch.Generator().manual_seed(seed)
        x = (torch.randn(B, C, *sp, generator=g) * 2).requires_grad_(True)
        y = torch.randint(0, C, (B, 1, *sp), generator=g)
        w = torch.rand(C, generator=g) + 0.5
        dl = DiceLoss()
        ld = dl(x, y)
        lc = torch.nn.CrossEntropyLoss(weight=w)(x, y.squeeze(1))
        (ld + lc).backward()
        xo = x.detach().clone().requires_grad_(True)
        lo = olosses.dice_loss(xo, y) + olosses.cross_entropy(xo, y, w)
        lo.backward()
        e1 = abs(lo.item() - (ld + lc).item())
        e2 = ((xo.grad - x.grad).abs().max() / x.grad.abs().max()).item()
        print("%-14s oracle-vs-reference: loss %.2e grad %.2e" % (name, e1, e2))
        assert e1 < 1e-5 and e2 < 1e-4
        torch.save({"x": x.detach(), "y": y, "w": w, "dice": ld.item(), "ce": lc.item(), "grad": x.grad.clone(),
                    "alpha": dl.alpha.detach().clone()},
                   os.path.join(ROOT, "tests", "golden", name + ".pt"))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
