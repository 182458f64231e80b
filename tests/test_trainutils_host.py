"""CPU: the stock-path optimiser / EMA glue of b200seg.train (get_optimizer, update_ema_variables: the caller-side pieces
the reference keeps in training/utils.py:8-14,98-105) against a trajectory the UNMODIFIED reference functions produced."""
import copy

import torch

from b200seg.train import get_optimizer, update_ema_variables
from oracle.make_golden_trainutils import grads, tiny
from util import load_golden


def test_adamw_and_ema_follow_the_reference():
    g = load_golden("trainutils_ref")
    a = g["args"]
    net = tiny()
    ema = copy.deepcopy(net)
    for p in ema.parameters():
        p.requires_grad_(False)
    opt = get_optimizer(net, base_lr=a["base_lr"], betas=tuple(a["betas"]), weight_decay=a["weight_decay"])
    for k, v in g["defaults"].items():
        got = opt.defaults[k]
        assert (tuple(got) == tuple(v)) if isinstance(v, (list, tuple)) else (got == v), k          # eps 1e-5: the AMP-stability choice
    for step, ref in enumerate(g["traj"]):
        for p, gr in zip(net.parameters(), grads(step, net)):
            p.grad = gr
        opt.step()
        update_ema_variables(net, ema, g["ema_alpha"], step)
        for p, r in zip(net.parameters(), ref["params"]):
            assert torch.allclose(p, r, rtol=1e-6, atol=1e-8)
        for p, r in zip(ema.parameters(), ref["ema"]):
            assert torch.allclose(p, r, rtol=1e-6, atol=1e-8)
