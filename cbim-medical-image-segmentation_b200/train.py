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


class FusedAdamWEMA:
    """The optimiser tail of the reference loop — GradScaler unscale + non-finite check, AdamW(eps=1e-5, wd=0.05)
    (training/utils.py:8-14), `scaler.update()` and `update_ema_variables` (training/utils.py:98-105,
    train_ddp.py:193-195,211) — as two multi-tensor kernels over a device-side tensor table:
    40 B/parameter of HBM traffic instead of 52, 3 launches instead of ~10 + 2 per EMA tensor, and no host sync
    (the loss scale, found_inf and the applied-step counter live on the device).  Numerics follow torch.optim.AdamW
    and torch.amp.GradScaler exactly (tests/test_gpu_optim.py compares them step by step, an overflow step included)."""

    def __init__(self, net, ema_net=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-5, weight_decay=0.05, ema_alpha=0.99, amp=True,
                 init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        from . import _lib
        self._lib = _lib
        m = net.module if hasattr(net, "module") else net
        e = None if ema_net is None else (ema_net.module if hasattr(ema_net, "module") else ema_net)
        self.params = [p for p in m.parameters()]
        self.ema_params = None if e is None else [p for p in e.parameters()]
        self.model, self.ema_model = m, e
        if self.ema_params is not None and len(self.ema_params) != len(self.params):
            raise ValueError("EMA model and model must have the same parameters (training/utils.py:101 zips them)")
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_alpha, self.amp = lr, betas, eps, weight_decay, ema_alpha, amp
        self.growth = (growth_factor, backoff_factor, growth_interval)
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.scale = torch.full((1,), init_scale if amp else 1.0, dtype=torch.float32, device=dev)
        self.growth_tracker = torch.zeros(1, dtype=torch.int32, device=dev)
        self.found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.iteration = 0
        self._tables = None
        self._live = None
        self._sig = None

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def scale_loss(self, loss):
        return loss * self.scale if self.amp else loss

    def _build(self):
        """(Re)build the device pointer table.  The chunk list depends only on the tensor sizes and is built once; the
        six-pointer rows are refreshed (pinned staging buffer, asynchronous copy) whenever a .grad was reallocated."""
        live = [i for i, p in enumerate(self.params) if p.grad is not None]
        dev = self.params[0].device
        if self._tables is None or self._live != live:
            chunk = self._lib.load().b200seg_optim_chunk_elems()
            chunks = []
            for j, i in enumerate(live):
                chunks.extend([j, s] for s in range(0, self.params[i].numel(), chunk))
            self._live = live
            self._rows_host = torch.empty(len(live), 6, dtype=torch.int64).pin_memory()
            self._tables = [torch.empty(len(live), 6, dtype=torch.int64, device=dev),
                            torch.tensor(chunks, dtype=torch.int64).to(dev), len(chunks)]
        for j, i in enumerate(live):
            p, g = self.params[i], self.params[i].grad
            if g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                raise self._lib.B200SegError("the fused optimiser tail needs contiguous fp32 parameters and gradients")
            e = 0 if self.ema_params is None else self.ema_params[i].data_ptr()
            self._rows_host[j] = torch.tensor([g.data_ptr(), p.data_ptr(), self.exp_avg[i].data_ptr(),
                                               self.exp_avg_sq[i].data_ptr(), e, p.numel()], dtype=torch.int64)
        self._tables[0].copy_(self._rows_host, non_blocking=True)

    def step(self):
        """Call after backward().  Everything stays on the device; nothing here synchronises."""
        from .ops import _stream
        sig = tuple((p.grad.data_ptr() if p.grad is not None else 0, p.data_ptr()) for p in self.params)
        if sig != self._sig:           # DDP bucket views / a reallocated .grad: rebuild the pointer table
            self._build()
            self._sig = sig
        tab, chunks, n = self._tables
        alpha = min(1 - 1 / (self.iteration + 1), self.ema_alpha)          # training/utils.py:100
        with torch.cuda.device(tab.device):
            st = _stream()
            if self.amp:
                self.found_inf.zero_()
                self._lib.call("b200seg_grads_nonfinite", tab.data_ptr(), chunks.data_ptr(), n, self.found_inf.data_ptr(), st)
            self._lib.call("b200seg_adamw_ema_step", tab.data_ptr(), chunks.data_ptr(), n, self.lr, self.betas[0], self.betas[1],
                           self.eps, self.weight_decay, alpha, self.step_dev.data_ptr(),
                           self.scale.data_ptr() if self.amp else None, self.found_inf.data_ptr() if self.amp else None, st)
            if self.amp:                   # scaler.update(): grow / back off the loss scale on the device
                torch._amp_update_scale_(self.scale, self.growth_tracker, self.found_inf, *self.growth)
        if self.ema_model is not None:
            for eb, mb in zip(self.ema_model.buffers(), self.model.buffers()):
                eb.copy_(mb)               # training/utils.py:104-105
        self.iteration += 1


class TrainStep:
    def __init__(self, net, ema_net=None, ce_weight=None, amp=True, aux_weight=None, ema_alpha=0.99, lr=1e-3, fused_tail=True):
        self.net, self.ema_net = net, ema_net
        self.amp = amp
        self.criterion = DiceCELoss(weight=ce_weight)
        self.aux_weight = aux_weight
        self.ema_alpha = ema_alpha
        self.fused = FusedAdamWEMA(net, ema_net, lr=lr, ema_alpha=ema_alpha, amp=amp) if fused_tail else None
        if self.fused is None:
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
        if self.fused is not None:
            self.fused.zero_grad()
            with torch.autocast(device_type="cuda", dtype=torch.float16, enabled=self.amp):
                loss = self._loss(self.net(img), label)
            self.fused.scale_loss(loss).backward()
            self.fused.step()
            self.step_idx += 1
            return loss.detach()
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
