"""The body of the reference's ``train_epoch`` loop (train_ddp.py:171-215) as a reusable object:
zero_grad -> (autocast) forward -> CE + Dice (fused) -> GradScaler backward -> optimizer step -> EMA.
The optimiser/EMA/scaler are the reference's own caller-side glue (training/utils.py:8-14,98-105,
train_ddp.py:96,193-195) and stay stock PyTorch; the forward/backward/loss are b200seg kernels."""
import torch

from .losses import DiceCELoss


def get_optimizer(net, base_lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05):
    """training/utils.py:8-14 (AdamW, eps=1e-5); `fused=True` is the same maths in one multi-tensor kernel."""
    return torch.optim.AdamW(net.parameters(), lr=base_lr, betas=betas, weight_decay=weight_decay, eps=1e-5,
                             fused=True)


@torch.no_grad()
def update_ema_variables(model, ema_model, alpha, global_step):
    """training/utils.py:98-105, as two foreach kernels instead of 2 per parameter."""
    alpha = min(1 - 1 / (global_step + 1), alpha)
    ep = [p for p in ema_model.parameters()]
    mp = [p.detach() for p in model.parameters()]
    torch._foreach_mul_(ep, alpha)
    torch._foreach_add_(ep, mp, alpha=1 - alpha)
    for eb, mb in zip(ema_model.buffers(), model.buffers()):
        eb.copy_(mb)


class TrainStep:
    def __init__(self, net, ema_net=None, ce_weight=None, amp=True, aux_weight=None, ema_alpha=0.99, lr=1e-3):
        self.net, self.ema_net = net, ema_net
        self.amp = amp
        self.criterion = DiceCELoss(weight=ce_weight)
        self.aux_weight = aux_weight
        self.ema_alpha = ema_alpha
        self.optimizer = get_optimizer(net, base_lr=lr)
        self.scaler = torch.amp.GradScaler("cuda", enabled=amp)
        self.step_idx = 0

    def _loss(self, result, label):
        if isinstance(result, (tuple, list)):
            if self.aux_weight is None or len(self.aux_weight) < len(result):
                raise ValueError("the model returned %d outputs (deep supervision, train_ddp.py:186-189): pass "
                                 "aux_weight with one weight per output" % len(result))
            return sum(self.aux_weight[j] * self.criterion(r, label) for j, r in enumerate(result))
        return self.criterion(result, label)

    def __call__(self, img, label):
        """img [B,C,D,H,W] float (device), label [B,1,D,H,W] int64/uint8 (device) -> loss tensor (device)."""
        self.optimizer.zero_grad(set_to_none=True)
        if self.amp:
            with torch.autocast(device_type="cuda", dtype=torch.float16):
                loss = self._loss(self.net(img), label)
            self.scaler.scale(loss).backward()
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            loss = self._loss(self.net(img), label)
            loss.backward()
            self.optimizer.step()
        if self.ema_net is not None:
            m = self.net.module if hasattr(self.net, "module") else self.net
            e = self.ema_net.module if hasattr(self.ema_net, "module") else self.ema_net
            update_ema_variables(m, e, self.ema_alpha, self.step_idx)
        self.step_idx += 1
        return loss.detach()
