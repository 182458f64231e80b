"""Pin oracle/medformer_ops.py against the REAL reference modules and write tests/golden/{biattn,dwconv}_*.pt.

Runs only where /root/reference exists.  BidirectionAttention (medformer_utils.py:11-97) is instantiated
unmodified with proj_type='linear'; forward hooks capture what its projections produce / consume, i.e. exactly the
inputs and outputs of the fused core the CUDA kernel replaces; autograd through the reference module gives the
gradients at the same boundary.  Usage:  python oracle/make_golden_medformer.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference          # noqa: E402
from oracle import medformer_ops as mops                 # noqa: E402

BIATTN = {  # name: (B, heads, feat spatial, map spatial, feat_dim, map_dim, seed)
    "biattn_a": (2, 2, (4, 6, 5), (3, 3, 3), 24, 16, 21),
    "biattn_b": (1, 4, (5, 6, 7), (3, 3, 3), 32, 32, 22),
    "biattn_c": (1, 1, (2, 3, 67), (2, 2, 2), 8, 8, 23),
}
DWCONV = {"dwconv_a": (2, 16, (5, 6, 7), [3, 3, 3], 31), "dwconv_b": (1, 24, (4, 9, 8), [1, 3, 3], 32)}


def main():
    import_reference()
    from model.dim3.medformer_utils import BidirectionAttention
    from model.dim3.conv_layers import DepthwiseSeparableConv
    out = os.path.join(ROOT, "tests", "golden")
    for name, (B, heads, fs, ms, fd, md, seed) in BIATTN.items():
        torch.manual_seed(seed)
        mod = BidirectionAttention(fd, md, fd, heads=heads, dim_head=32, map_size=list(ms), proj_type="linear")
        cap = {}
        def grab_out(key):
            def hook(m, i, o):
                o.retain_grad()
                cap[key] = o
            return hook

        def grab_in(key):
            def hook(m, i):
                i[0].retain_grad()
                cap[key] = i[0]
            return hook

        mod.feat_qv.register_forward_hook(grab_out("fqv"))
        mod.map_qv.register_forward_hook(grab_out("mqv"))
        mod.feat_out.register_forward_pre_hook(grab_in("fo"))
        mod.map_out.register_forward_pre_hook(grab_in("mo"))
        feat = torch.randn(B, fd, *fs) * 2.0
        smap = torch.randn(B, md, *ms) * 2.0
        y_f, y_m = mod(feat, smap)
        (y_f.square().sum() + y_m.square().sum() * 5).backward()
        fq, fv = cap["fqv"].detach().chunk(2, dim=1)
        mq, mv = cap["mqv"].detach().chunk(2, dim=1)
        # the oracle must reproduce the reference at this boundary
        qf, vf, qm, vm = (t.clone().requires_grad_(True) for t in (fq, fv, mq, mv))
        fo, mo = mops.bidirection_attention_core(qf, vf, qm, vm, heads)
        assert torch.allclose(fo, cap["fo"], atol=1e-6, rtol=1e-5) and torch.allclose(mo, cap["mo"], atol=1e-6, rtol=1e-5)
        torch.autograd.backward([fo, mo], [cap["fo"].grad, cap["mo"].grad])
        g_ref_f, g_ref_m = cap["fqv"].grad, cap["mqv"].grad
        assert torch.allclose(torch.cat([qf.grad, vf.grad], 1), g_ref_f, atol=1e-5, rtol=1e-4)
        assert torch.allclose(torch.cat([qm.grad, vm.grad], 1), g_ref_m, atol=1e-5, rtol=1e-4)
        torch.save({"heads": heads, "fqv": cap["fqv"].detach(), "mqv": cap["mqv"].detach(),
                    "fo": cap["fo"].detach(), "mo": cap["mo"].detach(), "dfo": cap["fo"].grad, "dmo": cap["mo"].grad,
                    "dfqv": g_ref_f, "dmqv": g_ref_m}, os.path.join(out, name + ".pt"))
        print(name, "ok", tuple(cap["fqv"].shape))
    for name, (B, C, sp, k, seed) in DWCONV.items():
        torch.manual_seed(seed)
        mod = DepthwiseSeparableConv(C, C, kernel_size=k, bias=False).depthwise
        x = torch.randn(B, C, *sp, requires_grad=True)
        y = mod(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        xo = x.detach().clone().requires_grad_(True)
        wo = mod.weight.detach().clone().requires_grad_(True)
        yo = mops.depthwise_conv3d(xo, wo)
        yo.backward(gy)
        assert torch.equal(yo, y) and torch.allclose(xo.grad, x.grad, atol=1e-6) and torch.allclose(wo.grad, mod.weight.grad, atol=1e-5)
        torch.save({"x": x.detach(), "w": mod.weight.detach(), "y": y.detach(), "gy": gy, "dx": x.grad, "dw": mod.weight.grad},
                   os.path.join(out, name + ".pt"))
        print(name, "ok")


if __name__ == "__main__":
    main()
