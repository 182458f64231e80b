"""Seeded synthetic volumes shared by tests, smoke() and bench.py (SURVEY.md §8d "Synthetic inputs"):
image = smooth field + noise, roughly N(0,1) after the dataset's normalisation; label = blob map with a
background-dominated histogram and every class present.  TEST/BENCH INFRASTRUCTURE."""
import torch
import torch.nn.functional as F


def make_volume(B, D, H, W, classes, seed=2023, in_ch=1):
    g = torch.Generator().manual_seed(seed)
    lo = [max(2, s // 8) for s in (D, H, W)]
    field = torch.randn(B, in_ch, *lo, generator=g)
    img = F.interpolate(field, size=(D, H, W), mode="trilinear", align_corners=True)
    img = img * 1.2 + 0.35 * torch.randn(B, in_ch, D, H, W, generator=g)
    img = img.clamp(-2.5, 2.7).contiguous()
    # labels: argmax of low-res random fields, class 0 biased to dominate
    lf = torch.randn(B, classes, *lo, generator=g)
    lf[:, 0] += 1.3
    lab = F.interpolate(lf, size=(D, H, W), mode="trilinear", align_corners=True).argmax(1, keepdim=True)
    flat = lab.view(B, -1)
    for c in range(classes):          # guarantee presence of every class
        flat[:, c] = c
    return img.float(), lab.long().contiguous()
