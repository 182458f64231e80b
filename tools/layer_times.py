"""Per-layer-shape timing of the tcgen05 conv kernels (forward with IN+ReLU loader + residual epilogue, data-gradient
with mask/statistics epilogue, weight gradient) on the layer shapes of the benchmarked ResUNet, CUDA events, C entry
points called back to back.  usage: python tools/layer_times.py [reps] > profiles/<round>_layer_times.txt"""
import json
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200seg import ops, _lib  # noqa

LAYERS = [  # Cin, Cout, k, (D,H,W)
    (32, 32, (1, 3, 3), (128, 128, 128)),
    (32, 32, (3, 3, 3), (128, 128, 128)),
    (96, 64, (1, 3, 3), (128, 128, 128)),     # fused conv1|shortcut of up4 (ACDC lists)
    (96, 64, (3, 3, 3), (128, 128, 128)),     # same, isotropic lists
    (64, 64, (1, 3, 3), (128, 64, 64)),
    (32, 128, (1, 3, 3), (128, 64, 64)),      # fused conv1|shortcut of down1
    (192, 128, (1, 3, 3), (128, 64, 64)),     # fused conv1|shortcut of up3
    (128, 128, (3, 3, 3), (128, 32, 32)),
    (384, 256, (3, 3, 3), (128, 32, 32)),     # fused conv1|shortcut of up2
    (256, 256, (3, 3, 3), (64, 16, 16)),
    (320, 320, (3, 3, 3), (32, 8, 8)),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
peak = peaks.get("bf16_tflops", 1590.0)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("peak %.1f TFLOP/s (%s)" % (peak, "measured burst" if peaks else "fallback"))
print("%-34s %10s %7s | %10s %7s | %10s %7s" % ("layer", "fwd us", "frac", "dgrad us", "frac", "wgrad us", "frac"))
for ci, co, k, (D, H, W) in LAYERS:
    x = torch.randn(1, D, H, W, ci, device="cuda").half()
    r = torch.randn(1, D, H, W, co, device="cuda").half()
    dy = torch.randn(1, D, H, W, co, device="cuda").half()
    st = ops.instnorm_stats(x, 0, ci)
    w = torch.randn(co, ci, *k, device="cuda") * 0.05
    algo = ops.conv_algo(ci, co, k, torch.float16, 1)
    algo_b = ops.conv_algo(co, ci, k, torch.float16, 1)
    wp = (ops.pack_weight(w, torch.float16, layout=algo), algo)
    wpb = (ops.pack_weight(w, torch.float16, True, layout=algo_b), algo_b)
    fl = 2.0 * D * H * W * ci * co * k[0] * k[1] * k[2]
    t_f = timed(lambda: ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k, residual=r))
    t_d = timed(lambda: ops.conv3d_fwd(dy, 0, co, None, ops.ACT_NONE, wpb, ci, k, dgrad_of=(x, 0, st, ops.ACT_RELU)))
    t_w = timed(lambda: ops.conv3d_wgrad(x, 0, ci, st, ops.ACT_RELU, dy, 0, co, k))
    f = lambda t: fl / (t * 1e-6) / 1e12 / peak
    print("%-34s %10.1f %7.3f | %10.1f %7.3f | %10.1f %7.3f" % ("%d->%d k%s @%s" % (ci, co, "".join(map(str, k)), "x".join(map(str, (D, H, W)))),
                                                               t_f, f(t_f), t_d, f(t_d), t_w, f(t_w)))
    del x, r, dy
