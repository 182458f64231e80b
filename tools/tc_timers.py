"""Cycle accounting of the tcgen05 conv kernel's warp roles (block 0) on the headline layer shapes.
   B200SEG_TC_DEBUG=1 python tools/tc_timers.py"""
import ctypes
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200SEG_TC_DEBUG", "1")
from b200seg import ops, _lib  # noqa

LAYERS = [(32, 32, (1, 3, 3), (128, 128, 128)), (64, 64, (1, 3, 3), (128, 64, 64)), (128, 128, (3, 3, 3), (128, 32, 32)),
          (256, 256, (3, 3, 3), (64, 16, 16)), (96, 64, (1, 3, 3), (128, 128, 128))]
lib = _lib.load()
buf = (ctypes.c_longlong * 32)()
for ci, co, k, (D, H, W) in LAYERS:
    x = torch.randn(1, D, H, W, ci, device="cuda").half()
    r = torch.randn(1, D, H, W, co, device="cuda").half()
    st = ops.instnorm_stats(x, 0, ci)
    algo = ops.conv_algo(ci, co, k, torch.float16, 1)
    wp = (ops.pack_weight(torch.randn(co, ci, *k, device="cuda") * 0.05, torch.float16, layout=algo), algo)
    ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k, residual=r)
    lib.b200seg_debug_tc_timers(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k, residual=r)
    e1.record()
    lib.b200seg_debug_tc_timers(buf)
    t = list(buf)
    ms = e0.elapsed_time(e1)
    ns = max(t[4], 1)
    print("layer %d->%d k%s @%s: %.1f us" % (ci, co, k, (D, H, W), ms * 1e3))
    print("  loader g0 (per stage, %d stages): wait_empty %.0f  issue_loads %.0f  transform+store(incl. load latency) %.0f  publish %.0f cycles"
          % (t[4], t[0] / ns, t[1] / ns, t[2] / ns, t[3] / ns))
    nst = max(t[12], 1)
    print("  mma: wait_T_EMPTY total %d | per stage (%d): wait_A_FULL %.0f  issue(incl. weight waits) %.0f  of which wait_B_FULL %.0f"
          % (t[8], t[12], t[9] / nst, t[11] / nst, t[10] / nst))
    nt = max(t[18], 1)
    print("  epilogue (per tile, %d tiles): wait_T_FULL %.0f  work %.0f" % (t[18], t[16] / nt, t[17] / nt))
