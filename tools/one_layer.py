"""Launch ONE layer shape a few times (for `ncu -k regex:...`).  usage: python tools/one_layer.py cin cout kd kh kw D H W [what=fwd|dgrad|wgrad] [reps]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200seg import ops  # noqa

ci, co, kd, kh, kw, D, H, W = map(int, sys.argv[1:9])
what = sys.argv[9] if len(sys.argv) > 9 else "fwd"
reps = int(sys.argv[10]) if len(sys.argv) > 10 else 3
k = (kd, kh, kw)
x = torch.randn(1, D, H, W, ci, device="cuda").half()
r = torch.randn(1, D, H, W, co, device="cuda").half()
dy = torch.randn(1, D, H, W, co, device="cuda").half()
st = ops.instnorm_stats(x, 0, ci)
w = torch.randn(co, ci, *k, device="cuda") * 0.05
algo = ops.conv_algo(ci, co, k, torch.float16, 1)
algo_b = ops.conv_algo(co, ci, k, torch.float16, 1)
wp = (ops.pack_weight(w, torch.float16, layout=algo), algo)
wpb = (ops.pack_weight(w, torch.float16, True, layout=algo_b), algo_b)
for _ in range(reps):
    if what == "fwd":
        ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k, residual=r)
    elif what == "dgrad":
        ops.conv3d_fwd(dy, 0, co, None, ops.ACT_NONE, wpb, ci, k, dgrad_of=(x, 0, st, ops.ACT_RELU))
    else:
        ops.conv3d_wgrad(x, 0, ci, st, ops.ACT_RELU, dy, 0, co, k)
torch.cuda.synchronize()
print("done")
