import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_err(a, b):
    """max-norm relative error: ||a-b||_inf / ||b||_inf (the parity metric of SURVEY.md §8d)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def dice_per_class(a, b, C):
    """metric/utils.py:62-82 style one-hot Dice between two label maps."""
    out = []
    for c in range(C):
        x, y = (a == c), (b == c)
        den = x.sum().item() + y.sum().item()
        out.append(1.0 if den == 0 else 2.0 * (x & y).sum().item() / den)
    return out


def grad_noise_floor(sd, img, lab, weight, cfg):
    """Gradients of the reference-pinned oracle in fp64 (the exact answer) and the distance of the fp32
    evaluation (= the reference's own numbers, tests/golden) from it.  ReLU'(0) is discontinuous, so two valid
    fp32 evaluations differ wherever a pre-activation is within rounding of 0, and InstanceNorm spreads each
    such flip over a whole channel: this distance is the noise floor of "matches the reference's backward"."""
    from oracle import losses as olosses
    from oracle import unet3d as ounet
    out = {}
    for dt in (torch.float32, torch.float64):
        s = {k: v.to(dt).clone().requires_grad_(True) for k, v in sd.items()}
        lo = ounet.unet_forward(s, img.to(dt), cfg["scale"], cfg["kernel"], cfg["block"])
        olosses.total_loss(lo, lab, weight.to(dt)).backward()
        out[dt] = ({k: v.grad.double() for k, v in s.items()}, lo.detach().double())
    g32, g64 = out[torch.float32][0], out[torch.float64][0]
    floor_max = max(rel_err(g32[k], g64[k]) for k in g64)
    return g64, out[torch.float64][1], floor_max, global_l2(g32, g64)


def global_l2(ga, gb):
    num = sum(((ga[k].double().cpu() - gb[k].double().cpu()) ** 2).sum().item() for k in gb)
    den = sum((gb[k].double().cpu() ** 2).sum().item() for k in gb)
    return (num / den) ** 0.5
