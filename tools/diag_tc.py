"""Tiny smoke of the tcgen05 kernels for bring-up: each call synchronised and checked against F.conv3d, small volumes."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200seg import ops

def run(ci, co, k, D, H, W, what):
    torch.manual_seed(0)
    x = torch.randn(1, D, H, W, ci, device="cuda").half()
    dy = torch.randn(1, D, H, W, co, device="cuda").half()
    w = (torch.randn(co, ci, *k, device="cuda") * 0.05)
    st = ops.instnorm_stats(x, 0, ci)
    xn = F.relu(F.instance_norm(x.float().permute(0, 4, 1, 2, 3), eps=1e-4)).half()
    pad = [i // 2 for i in k]
    if what == "fwd":
        algo = ops.conv_algo(ci, co, k, torch.float16, 1)
        wp = (ops.pack_weight(w, torch.float16, layout=algo), algo)
        y, yst = ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k)
        torch.cuda.synchronize()
        ref = F.conv3d(xn.float(), w.half().float(), padding=pad).permute(0, 2, 3, 4, 1)
        err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
        s_err = ((yst[0, :, 0].float() - y.float().sum((0, 1, 2, 3))).abs().max() / y.float().sum((0, 1, 2, 3)).abs().max()).item()
        print("fwd  %d->%d k%s %s: rel err %.2e, stats err %.2e" % (ci, co, k, (D, H, W), err, s_err), flush=True)
    else:
        dw, _ = ops.conv3d_wgrad(x, 0, ci, st, ops.ACT_RELU, dy, 0, co, k)
        torch.cuda.synchronize()
        xr = xn.float().requires_grad_(False)
        wr = w.clone().requires_grad_(True)
        F.conv3d(xr, wr, padding=pad).backward(dy.float().permute(0, 4, 1, 2, 3))
        err = ((dw - wr.grad).abs().max() / wr.grad.abs().max()).item()
        print("wgrad %d->%d k%s %s: rel err %.2e" % (ci, co, k, (D, H, W), err), flush=True)

for what in ("fwd", "wgrad"):
    for (ci, co, k, dims) in [(32, 32, (1, 3, 3), (8, 32, 32)), (64, 64, (3, 3, 3), (8, 32, 32)), (128, 128, (3, 3, 3), (8, 16, 16)),
                              (96, 64, (1, 3, 3), (8, 32, 32)), (48, 48, (3, 3, 3), (8, 16, 16)), (384, 256, (3, 3, 3), (8, 16, 16))]:
        run(ci, co, k, *dims, what)
print("diag done")
