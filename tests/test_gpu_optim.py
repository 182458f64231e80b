"""GPU: the fused optimiser tail against the reference loop's own pieces step by step —
torch.optim.AdamW(eps=1e-5, weight_decay=0.05) (training/utils.py:8-14), torch.amp.GradScaler
(train_ddp.py:193-195) and update_ema_variables (training/utils.py:98-105) — including a step whose gradients
overflow (skipped by both, loss scale halved, EMA still updated)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ema_reference(model_params, ema_params, alpha, global_step):
    alpha = min(1 - 1 / (global_step + 1), alpha)
    for e, p in zip(ema_params, model_params):
        e.data.mul_(alpha).add_(p.data, alpha=1 - alpha)      # training/utils.py:101-102


@pytest.mark.parametrize("amp", [True, False])
def test_fused_adamw_ema_matches_torch(amp):
    from b200seg.train import FusedAdamWEMA
    torch.manual_seed(3)
    shapes = [(33, 7, 3, 3, 3), (4096,), (5,), (128, 64, 1, 1, 1), (3, 1)]
    net = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes])
    ema = torch.nn.ParameterList([torch.nn.Parameter(p.detach().clone(), requires_grad=False) for p in net])
    ref = [p.detach().clone().requires_grad_(True) for p in net]
    ref_ema = [p.detach().clone() for p in net]
    opt = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5, fused=True)
    scaler = torch.amp.GradScaler("cuda", enabled=amp, init_scale=1024.0, growth_interval=3)
    fused = FusedAdamWEMA(net, ema, lr=1e-3, amp=amp, init_scale=1024.0, growth_interval=3)
    for it in range(7):
        grads = [torch.randn_like(p) * 0.1 for p in net]
        if amp and it == 2:
            grads[1][17] = float("inf")                  # an overflow step
        if amp:
            scaler.scale(torch.zeros(1, device="cuda"))      # lazily creates the scaler's device-side scale tensor
        scale = scaler.get_scale() if amp else 1.0
        for p, r, g in zip(net, ref, grads):
            p.grad = (g * scale).clone()
            r.grad = (g * scale).clone()
        fused.step()
        if amp:
            scaler.step(opt)
            scaler.update()
        else:
            opt.step()
        _ema_reference(ref, ref_ema, 0.99, it)
        if amp:
            assert abs(fused.scale.item() - scaler.get_scale()) < 1e-6, (it, fused.scale.item(), scaler.get_scale())
        for p, r in zip(net, ref):
            assert torch.allclose(p, r, rtol=2e-6, atol=1e-7), (it, (p - r).abs().max().item())
        for e, r in zip(ema, ref_ema):
            assert torch.allclose(e, r, rtol=2e-6, atol=1e-7), (it, (e - r).abs().max().item())
    assert fused.step_dev.item() == (6 if amp else 7)
