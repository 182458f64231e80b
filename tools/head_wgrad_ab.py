"""A/B: the 1x1x1 few-class head weight gradient on the HBM-bound special kernel (conv3d_wgrad_small, what ALGO_AUTO picks)
versus the bias-gradient pass + the tcgen05 weight-gradient kernel, for the head shapes of the benchmarked models."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200seg                                   # noqa: E402,F401
from b200seg import _lib, ops                     # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for Cin, Cout, shape in [(48, 16, (1, 128, 128, 128)), (32, 16, (1, 96, 96, 96)), (64, 16, (1, 96, 48, 48)), (32, 8, (1, 128, 128, 128))]:
    B, D, H, W = shape
    x = torch.randn(B, D, H, W, Cin, device="cuda").half()
    dy = torch.randn(B, D, H, W, Cout, device="cuda").half()
    t_auto = timed(lambda: ops.conv3d_wgrad(x, 0, Cin, None, ops.ACT_NONE, dy, 0, Cout, (1, 1, 1), want_bias=True, algo=_lib.ALGO_AUTO))
    try:
        t_tc = timed(lambda: ops.conv3d_wgrad(x, 0, Cin, None, ops.ACT_NONE, dy, 0, Cout, (1, 1, 1), want_bias=False, algo=_lib.ALGO_TC))
    except Exception as e:
        t_tc = float("nan")
        print("TC unsupported:", e)
    mb = (x.numel() + dy.numel()) * 2 / 1e6
    print("head wgrad %d->%d @%s: special kernel (with bias) %.1f us, tcgen05 (no bias; + ~10 us bias pass) %.1f us, HBM floor %.1f us"
          % (Cin, Cout, shape[1:], t_auto, t_tc, mb / 6.57e3 * 1e3 / 1e3 * 1e3 / 1e3 if False else mb / 6570.0 * 1e3))
