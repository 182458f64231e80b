"""CPU emulation of the MedFormer-specific C-ABI ops (TEST INFRASTRUCTURE, companion of emu_ops.py).

The launch wrappers that have one (`dwconv3d`, `biattn_fwd`, ...) are replaced at the wrapper level, so the real
autograd Functions (their backward algebra) still run; Functions that talk to the library directly through `call`
are replaced by plain-PyTorch stand-ins with the same `apply` signature, so the module wiring of
b200seg.medformer (norm eps per call site, which outputs carry IN sums, residual routing, channel padding, concat
order, token layout) is exercised end to end without a GPU.  The product never imports this file."""
import torch
import torch.nn.functional as F

import emu_ops
from emu_ops import _ncdhw, _normalise, _stats


def _cl(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


def install(monkeypatch):
    emu_ops.install(monkeypatch)
    from b200seg import medformer as mf
    from b200seg import medformer_ops as mo
    from b200seg import ops
    from oracle import medformer_ops as oracle_mops

    # launch wrappers that medformer_ops imported by name: point them at the (already emulated) ops.* versions
    for name in ("conv3d_fwd", "conv3d_wgrad", "in_bwd_apply", "in_bwd_reduce", "instnorm_stats", "copy_channels"):
        monkeypatch.setattr(mo, name, getattr(ops, name))

    def _dw_weight(w_taps, ksize, flip, cmajor):
        if cmajor:                                   # the module's own [C,1,kd,kh,kw] parameter
            w = w_taps.reshape(-1, 1, *ksize).double()
        else:
            C = w_taps.shape[1]
            w = w_taps.t().reshape(C, 1, *ksize).double()
        return w.flip(2, 3, 4) if flip else w

    def dwconv3d(x, w_taps, ksize, x_stats=None, act=0, flip=False, want_stats=False, eps=ops.IN_EPS, cmajor=False):
        C = x.shape[-1]
        a = _ncdhw(x, 0, C)
        a = _normalise(a, x_stats, act, eps) if x_stats is not None else (F.relu(a) if act else a)
        y = F.conv3d(a.double(), _dw_weight(w_taps, ksize, flip, cmajor), padding=[k // 2 for k in ksize], groups=C)
        return _cl(y.to(x.dtype)), (_stats(y) if want_stats else None)

    def dwconv3d_wgrad(x, dy, ksize, x_stats=None, act=0, eps=ops.IN_EPS, cmajor=False):
        C = x.shape[-1]
        a = _ncdhw(x, 0, C)
        a = (_normalise(a, x_stats, act, eps) if x_stats is not None else (F.relu(a) if act else a)).double()
        w = torch.zeros(C, 1, *ksize, dtype=torch.float64, requires_grad=True)
        with torch.enable_grad():
            F.conv3d(a, w, padding=[k // 2 for k in ksize], groups=C).backward(_ncdhw(dy, 0, C).double())
        return w.grad.float() if cmajor else w.grad.reshape(C, -1).t().contiguous().float()

    def _core(fqv, mqv, heads):
        f = fqv.permute(0, 4, 1, 2, 3).double()
        m = mqv.permute(0, 4, 1, 2, 3).double()
        return oracle_mops.bidirection_attention_core(*f.chunk(2, 1), *m.chunk(2, 1), heads)

    def biattn_fwd(fqv, mqv, heads, dim_head=32):
        fo, mo_ = _core(fqv, mqv, heads)
        return _cl(fo.to(fqv.dtype)), _cl(mo_.to(fqv.dtype)), torch.zeros(1)

    def biattn_bwd(fqv, mqv, mo_, colstat, dfo, dmo, heads, dim_head=32):
        with torch.enable_grad():
            f = fqv.detach().clone().requires_grad_(True)
            m = mqv.detach().clone().requires_grad_(True)
            fo, mo2 = _core(f, m, heads)
            torch.autograd.backward([fo, mo2], [dfo.permute(0, 4, 1, 2, 3).double(), dmo.permute(0, 4, 1, 2, 3).double()])
        return f.grad, m.grad

    for name, fn in dict(dwconv3d=dwconv3d, dwconv3d_wgrad=dwconv3d_wgrad, biattn_fwd=biattn_fwd, biattn_bwd=biattn_bwd).items():
        monkeypatch.setattr(ops, name, fn)
        if hasattr(mo, name):
            monkeypatch.setattr(mo, name, fn)

    # ---- Functions that call the library directly: plain-PyTorch stand-ins with the same apply() signature
    def _st(y_cl):
        return _stats(y_cl.detach().permute(0, 4, 1, 2, 3))

    class SpaceToDepthFn:
        @staticmethod
        def apply(x, scale):
            sd, sh, sw = scale
            y = torch.cat([x[:, i::sd, j::sh, k::sw, :] for i in range(sd) for j in range(sh) for k in range(sw)], -1).contiguous()
            return y, _st(y)

    class MapGenFn:
        @staticmethod
        def apply(fw, C, K, map_size):
            B = fw.shape[0]
            flat = fw.reshape(B, -1, fw.shape[-1])
            wm = F.softmax(flat[..., C:C + K], dim=1)                       # softmax over the voxels
            return torch.einsum("bjc,bjk->bkc", flat[..., :C], wm).reshape(B, *map_size, C)

    class SEScaleFn:
        @staticmethod
        def apply(x, x_stats, w1, b1, w2, b2):
            mean = x.mean(dim=(1, 2, 3))
            h = F.relu(F.linear(mean, w1.flatten(1), b1))
            gate = torch.sigmoid(F.linear(h, w2.flatten(1), b2))
            y = x * gate[:, None, None, None, :]
            return y, _st(y)

    class UpsampleFn:
        @staticmethod
        def apply(x, size):
            return _cl(F.interpolate(x.permute(0, 4, 1, 2, 3), size=size, mode="trilinear", align_corners=True))

    class LayerNormFn:
        @staticmethod
        def apply(x, gamma, beta, eps):
            return F.layer_norm(x, x.shape[-1:], gamma, beta, eps)

    class GeluFn:
        @staticmethod
        def apply(x):
            return F.gelu(x)

    class MHSAFn:
        @staticmethod
        def apply(qkv, heads, dim_head):
            B, L, _ = qkv.shape
            q, k, v = (t.reshape(B, L, heads, dim_head).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
            att = F.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * dim_head ** -0.5, dim=-1)
            return torch.einsum("bhij,bhjd->bhid", att, v).permute(0, 2, 1, 3).reshape(B, L, heads * dim_head)

    for name, cls in dict(SpaceToDepthFn=SpaceToDepthFn, MapGenFn=MapGenFn, SEScaleFn=SEScaleFn, UpsampleFn=UpsampleFn,
                          LayerNormFn=LayerNormFn, GeluFn=GeluFn, MHSAFn=MHSAFn).items():
        monkeypatch.setattr(mf, name, cls)
    monkeypatch.setattr(mf, "UpCatFn", ops.UpCatFn)        # emu_ops' version
