"""Loss modules behind the reference's loss contract:
  * ``DiceLoss()`` — no-arg constructor, ``__call__(preds[B,C,...], targets[B,1,...] int64)`` ->
    0-dim tensor (training/losses.py:8-58), incl. the ``alpha``/``beta`` side effect (losses.py:38-41);
  * ``DiceCELoss(weight)`` — the sum ``CrossEntropyLoss(weight)(x, y.squeeze(1)) + DiceLoss()(x, y)``
    the trainer builds at train_ddp.py:93-94,189-191, in ONE kernel pass each way.
Both run the fused sm_100a kernel; there is no PyTorch fallback."""
import torch
import torch.nn as nn

from .ops import DiceCEFn


class DiceCELoss(nn.Module):
    def __init__(self, weight=None, ce_scale=1.0, dice_scale=1.0):
        super().__init__()
        if weight is not None:
            self.register_buffer("weight", torch.as_tensor(weight, dtype=torch.float32), persistent=False)
        else:
            self.weight = None
        self.ce_scale = float(ce_scale)
        self.dice_scale = float(dice_scale)

    def forward(self, preds, targets):
        return DiceCEFn.apply(preds, targets, self.weight, self.ce_scale, self.dice_scale)


class DiceLoss(nn.Module):
    """Adaptive-Tversky Dice of the reference (alpha from FP/(FP+FN), clamped, differentiable)."""

    def __init__(self, alpha=0.5, beta=0.5, size_average=True, reduce=True):
        super().__init__()
        if not (size_average and reduce):
            raise ValueError("the fused kernel implements size_average=True, reduce=True (the trainer's use)")
        self.alpha = alpha
        self.beta = beta

    def forward(self, preds, targets):
        loss = DiceCEFn.apply(preds, targets, None, 0.0, 1.0)
        # the reference overwrites self.alpha / self.beta with the per-class tensors it just computed
        # (training/losses.py:38-41); the kernel leaves alpha_c in its stats buffer (dice_ce.cu: out[4+2C+c])
        st = getattr(DiceCEFn, "last_stats", None)
        if st is not None:
            C = preds.shape[1]
            self.alpha = st[4 + 2 * C:4 + 3 * C].detach().clone()
            self.beta = 1.0 - self.alpha
        return loss


class CrossEntropyLoss(nn.Module):
    """Weighted CE with the trainer's calling convention (target already squeezed or [B,1,...])."""

    def __init__(self, weight=None):
        super().__init__()
        self.weight = None if weight is None else torch.as_tensor(weight, dtype=torch.float32)

    def forward(self, preds, targets):
        return DiceCEFn.apply(preds, targets, self.weight, 1.0, 0.0)
