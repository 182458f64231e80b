"""Pin b200seg.inference against the UNMODIFIED reference functions (inference/inference3d.py, inference/utils.py,
metric/utils.py) and write tests/golden/inference_ref.pt: sliding-window / whole-image probabilities of a tiny
seeded network on volumes that exercise the border-snapped last window and the pad-to-window path, and the Dice
metric's (dice, intersection, summ) on seeded label maps.  Runs only where /root/reference exists.
Usage:  python oracle/make_golden_inference.py"""
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import REF          # noqa: E402

CASES = {"exact": dict(shape=(16, 24, 24), window=[8, 8, 12]),
         "snapped": dict(shape=(18, 21, 26), window=[8, 12, 10]),          # last windows snap to the border
         "padded": dict(shape=(6, 24, 20), window=[8, 16, 16])}            # volume smaller than the window on one axis
CLASSES = 3


def tiny_net(seed=3):
    torch.manual_seed(seed)
    net = nn.Sequential(nn.Conv3d(1, 6, 3, padding=1), nn.Tanh(), nn.Conv3d(6, CLASSES, 1))
    return net


def main():
    sys.path.insert(0, REF)
    import inference.inference3d as ref_inf
    from inference.utils import get_inference as ref_get
    import metric.utils as ref_metric
    net = tiny_net()
    out = {"net_seed": 3, "classes": CLASSES, "cases": {}}
    g = torch.Generator().manual_seed(41)
    for name, c in CASES.items():
        img = torch.randn(1, 1, *c["shape"], generator=g)
        args = types.SimpleNamespace(window_size=c["window"], classes=CLASSES, dimension="3d", sliding_window=True)
        assert ref_get(args) is ref_inf.inference_sliding_window
        sw = ref_inf.inference_sliding_window(net, img, args)
        whole = ref_inf.inference_whole_image(net, img, args)
        out["cases"][name] = {"shape": c["shape"], "window": c["window"], "img": img, "sliding": sw, "whole": whole}
        print(name, tuple(sw.shape), float(sw.sum()))
    pred = torch.randint(0, 5, (70001, 1), generator=g)
    target = torch.randint(0, 5, (70001, 1), generator=g)
    target[:5000] = pred[:5000]
    d, i, s = ref_metric.calculate_dice(pred, target, 5)
    ds, is_, ss = ref_metric.calculate_dice_split(pred, target, 5, block_size=20000)
    out["dice"] = {"pred": pred.to(torch.uint8), "target": target.to(torch.uint8), "dice": d, "intersection": i, "summ": s,
                   "split": (ds, is_, ss), "block_size": 20000}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "inference_ref.pt"))


if __name__ == "__main__":
    main()
