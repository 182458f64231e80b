"""Pin b200seg.train's optimiser / EMA glue against the UNMODIFIED reference functions (training/utils.py:8-14 get_optimizer,
:98-105 update_ema_variables): three AdamW steps with seeded gradients and EMA updates on a tiny parameter set ->
tests/golden/trainutils_ref.pt.  Runs only where /root/reference exists.
Usage:  python oracle/make_golden_trainutils.py"""
import copy
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import REF          # noqa: E402


def tiny(seed=5):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv3d(1, 4, 3, bias=False), nn.Conv3d(4, 2, 1))


def grads(step, net):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(p.shape, generator=g) * 0.1 for p in net.parameters()]


def main():
    sys.path.insert(0, REF)
    import training.utils as ref
    args = types.SimpleNamespace(optimizer="adamw", base_lr=1e-3, betas=[0.9, 0.999], weight_decay=0.05, momentum=0.9)
    net = tiny()
    ema = copy.deepcopy(net)
    for p in ema.parameters():
        p.requires_grad_(False)
    opt = ref.get_optimizer(args, net)
    traj = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for step in range(3):
            for p, g in zip(net.parameters(), grads(step, net)):
                p.grad = g
            opt.step()
            ref.update_ema_variables(net, ema, 0.99, step)
            traj.append({"params": [p.detach().clone() for p in net.parameters()], "ema": [p.detach().clone() for p in ema.parameters()]})
    torch.save({"args": vars(args), "ema_alpha": 0.99, "traj": traj, "defaults": {k: v for k, v in opt.defaults.items() if k in ("lr", "betas", "eps", "weight_decay")}},
               os.path.join(ROOT, "tests", "golden", "trainutils_ref.pt"))
    print("trainutils_ref:", opt.defaults["eps"], float(traj[-1]["params"][0].sum()), float(traj[-1]["ema"][0].sum()))


if __name__ == "__main__":
    main()
