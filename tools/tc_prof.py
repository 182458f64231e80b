"""Role-level cycle accounting of the tcgen05 kernels (instrumented twin library: python b200seg/build.py --profile).
For one layer shape prints, per wait site / scope, the average cycles per entry and the share of the kernel's lifetime
the role's warps spend there.  usage: python tools/tc_prof.py [layer index ...]   (indices into tools/layer_times.LAYERS)"""
import ctypes
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200SEG_LIB", os.path.join(ROOT, "cbim-medical-image-segmentation_b200", "libb200seg_prof.so"))
import torch  # noqa
from b200seg import ops, _lib  # noqa

LAYERS = [(32, 32, (1, 3, 3), (128, 128, 128)), (32, 32, (3, 3, 3), (128, 128, 128)), (96, 64, (1, 3, 3), (128, 128, 128)),
          (64, 64, (1, 3, 3), (128, 64, 64)), (192, 128, (1, 3, 3), (128, 64, 64)), (128, 128, (3, 3, 3), (128, 32, 32)),
          (384, 256, (3, 3, 3), (128, 32, 32))]
CONV = {1: ("loader: wait A_EMPTY", 8), 2: ("weights warp: wait B_EMPTY", 1), 6: ("epilogue: wait T_FULL", 8), 7: ("loader: wait TMA landed", 8),
        8: ("mma: wait T_EMPTY", 1), 9: ("mma: wait A stage", 1), 10: ("mma: wait B stage", 1), 11: ("loader: cp.async wait", 8),
        12: ("loader: transform", 8), 13: ("epilogue: chunk body", 8), 15: ("loader: issue next stage (incl. A_EMPTY wait)", 8),
        17: ("loader: fence + arrive", 8), 19: ("epilogue: tcgen05.ld + wait", 8), 21: ("epilogue: whole tile after cursor step", 8), 18: ("loader: cursor advance", 8), 31: ("lifetime (all warps)", 20)}
WG = {1: ("loader: wait EMPTY", 8), 3: ("epilogue: wait DONE", 4), 4: ("transposer: wait operand", 4), 5: ("transposer: wait A_FREE", 4),
      7: ("loader: wait TMA landed", 8), 9: ("mma: wait operands", 1), 10: ("mma: wait A_READY", 1), 11: ("loader: cp.async wait", 8),
      12: ("loader: transform", 8), 15: ("loader: issue next stage", 8), 16: ("transposer: smem->TMEM", 4), 31: ("lifetime (all warps)", 13)}
lib = ctypes.CDLL(os.environ["B200SEG_LIB"])


def report(fn_name, names, title):
    out = (ctypes.c_ulonglong * 64)()
    torch.cuda.synchronize()
    assert getattr(lib, fn_name)(out, 1) == 0
    life_n = out[63]
    if not life_n:
        return
    life = out[31] / life_n
    print("  %s: lifetime %.0f cycles/warp" % (title, life))
    for code, (name, nw) in sorted(names.items()):
        if code == 31 or not out[32 + code]:
            continue
        ncta = life_n / names[31][1]
        print("    %-32s %9.0f cyc/entry  %8.1f entries/warp  %5.1f%% of the role's time"
              % (name, out[code] / out[32 + code], out[32 + code] / (ncta * nw), 100.0 * out[code] / (ncta * nw) / life))


def reset():
    out = (ctypes.c_ulonglong * 64)()
    torch.cuda.synchronize()
    lib.b200seg_conv_tc_prof(out, 1); lib.b200seg_wgrad_tc_prof(out, 1)


sel = [int(a) for a in sys.argv[1:]] or range(len(LAYERS))
for i in sel:
    ci, co, k, (D, H, W) = LAYERS[i]
    print("%d->%d k%s @%s" % (ci, co, "".join(map(str, k)), "x".join(map(str, (D, H, W)))))
    x = torch.randn(1, D, H, W, ci, device="cuda").half()
    r = torch.randn(1, D, H, W, co, device="cuda").half()
    dy = torch.randn(1, D, H, W, co, device="cuda").half()
    st = ops.instnorm_stats(x, 0, ci)
    w = torch.randn(co, ci, *k, device="cuda") * 0.05
    algo, algo_b = ops.conv_algo(ci, co, k, torch.float16, 1), ops.conv_algo(co, ci, k, torch.float16, 1)
    wp = (ops.pack_weight(w, torch.float16, layout=algo), algo)
    wpb = (ops.pack_weight(w, torch.float16, True, layout=algo_b), algo_b)
    reset()
    ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k, residual=r)
    report("b200seg_conv_tc_prof", CONV, "forward")
    ops.conv3d_fwd(dy, 0, co, None, ops.ACT_NONE, wpb, ci, k, dgrad_of=(x, 0, st, ops.ACT_RELU))
    report("b200seg_conv_tc_prof", CONV, "dgrad")
    ops.conv3d_wgrad(x, 0, ci, st, ops.ACT_RELU, dy, 0, co, k)
    report("b200seg_wgrad_tc_prof", WG, "wgrad")
    del x, r, dy
