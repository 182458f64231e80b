"""Launch the MedFormer hot kernels at their largest BCV-96^3 shapes (for `ncu --set full -k regex:...`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg  # noqa
from b200seg import ops
from b200seg._lib import ACT_RELU

torch.manual_seed(0)
dev = "cuda"
# B-MHA at down2/up2: N = 96*24*24, 4 heads
f = torch.randn(1, 96, 24, 24, 256, device=dev).half()
m = torch.randn(1, 3, 3, 3, 256, device=dev).half()
fo, mo, cs = ops.biattn_fwd(f, m, 4)
dfo, dmo = torch.randn_like(fo), torch.randn_like(mo)
# MBConv depthwise at the same level: 512 channels, IN+ReLU prologue, IN sums epilogue
x = torch.randn(1, 96, 24, 24, 512, device=dev).half()
st = ops.instnorm_stats(x, 0, 512)
wt = torch.randn(27, 512, device=dev)
for _ in range(3):
    ops.biattn_fwd(f, m, 4)
    ops.biattn_bwd(f, m, mo, cs, dfo, dmo, 4)
    ops.dwconv3d(x, wt, (3, 3, 3), x_stats=st, act=ACT_RELU, want_stats=True)
    ops.dwconv3d(x, wt, (3, 3, 3), flip=True)
    ops.dwconv3d_wgrad(x, x, (3, 3, 3), x_stats=st, act=ACT_RELU)
torch.cuda.synchronize()
print("done")
