"""Launch the conv kernels on chosen layer shapes (for ncu).  usage: python tools/prof_layers.py [reps]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200seg import ops, _lib  # noqa

LAYERS = [  # Cin, Cout, k, (D,H,W), has residual
    (32, 32, (1, 3, 3), (128, 128, 128), True),      # level-0 layer of the ACDC-list ResUNet (HBM / smem-read bound)
    (128, 128, (3, 3, 3), (128, 32, 32), True),      # the FLOP-dominant layer shape (22 % of the conv FLOPs)
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for ci, co, k, (D, H, W), res in LAYERS:
    x = torch.randn(1, D, H, W, ci, device="cuda").half()
    r = torch.randn(1, D, H, W, co, device="cuda").half() if res else None
    dy = torch.randn(1, D, H, W, co, device="cuda").half()
    st = ops.instnorm_stats(x, 0, ci)
    algo = ops.conv_algo(ci, co, k, torch.float16, 1)
    wp = (ops.pack_weight(torch.randn(co, ci, *k, device="cuda") * 0.05, torch.float16, layout=algo), algo)
    for _ in range(reps):
        ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k, residual=r)
        ops.conv3d_wgrad(x, 0, ci, st, ops.ACT_RELU, dy, 0, co, k)
    torch.cuda.synchronize()
print("done")
