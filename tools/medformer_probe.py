"""GPU timing of the MedFormer kernels against the same maths in stock PyTorch on the same GPU.
Shapes: BCV 96^3 crop, base 32 -> levels 48^3x64 (heads 2), 24^3x128 (4), 12^3x256 (8); map 3x3x3."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg  # noqa
from b200seg import ops
from oracle import medformer_ops as mops   # timing comparator only (tools/, not the product path)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = []
for (S, heads) in [(48, 2), (24, 4), (12, 8)]:
    inner = 32 * heads
    f = torch.randn(2, S, S, S, 2 * inner, device="cuda", dtype=torch.float16)
    m = torch.randn(2, 3, 3, 3, 2 * inner, device="cuda", dtype=torch.float16)
    fo, mo, cs = ops.biattn_fwd(f, m, heads)
    dfo, dmo = torch.randn_like(fo), torch.randn_like(mo)
    t_f = timeit(lambda: ops.biattn_fwd(f, m, heads))
    t_b = timeit(lambda: ops.biattn_bwd(f, m, mo, cs, dfo, dmo, heads))
    fn = f.permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    mn = m.permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)

    def torch_fb():
        a, b = mops.bidirection_attention_core(*fn.chunk(2, 1), *mn.chunk(2, 1), heads)
        torch.autograd.backward([a, b], [torch.ones_like(a), torch.ones_like(b)])
    t_t = timeit(torch_fb, 5)
    nbytes_f = f.numel() * 2 + fo.numel() * 2
    out.append({"op": "biattn", "S": S, "heads": heads, "fwd_ms": t_f, "bwd_ms": t_b, "torch_fwd_bwd_ms": t_t,
                "fwd_GBps": nbytes_f / t_f / 1e6})
    C = inner
    x = torch.randn(2, S, S, S, C, device="cuda", dtype=torch.float16)
    wt = torch.randn(27, C, device="cuda")
    t_d = timeit(lambda: ops.dwconv3d(x, wt, (3, 3, 3)))
    t_w = timeit(lambda: ops.dwconv3d_wgrad(x, x, (3, 3, 3)))
    xn = x.permute(0, 4, 1, 2, 3).contiguous()
    wn = torch.randn(C, 1, 3, 3, 3, device="cuda", dtype=torch.float16)
    t_c = timeit(lambda: torch.nn.functional.conv3d(xn, wn, padding=1, groups=C), 5)
    out.append({"op": "dwconv", "S": S, "C": C, "fwd_ms": t_d, "wgrad_ms": t_w, "cudnn_fwd_ms": t_c,
                "fwd_GBps": 2 * x.numel() * 2 / t_d / 1e6})
for o in out:
    print(json.dumps(o))
