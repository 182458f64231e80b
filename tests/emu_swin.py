"""CPU emulation of the SwinUNETR-specific C-ABI ops (TEST INFRASTRUCTURE, companion of emu_ops.py / emu_medformer.py).

The autograd Functions of b200seg.swin_unetr that talk to the library directly are replaced by plain-PyTorch stand-ins
with the same `apply` signature — written from the reference's semantics (model/dim3/swin_unetr.py), not from the
kernels — so the module wiring of b200seg.SwinUNETR (which tensor is the shortcut, what is padded when, the qkv channel
order, merging order, the depth<->space shuffles around the GEMMs, decoder concat order) runs end to end on the CPU."""
import torch
import torch.nn.functional as F

import emu_medformer
from oracle import swin_ops as so


def install(monkeypatch):
    emu_medformer.install(monkeypatch)
    from b200seg import swin_unetr as sw

    class DepthSpaceFn:
        """space-to-depth with channel q*C + c, q = (i*2 + j)*2 + k, or its inverse"""
        @staticmethod
        def apply(x, to_depth):
            B, D, H, W, C = x.shape
            if to_depth:
                y = x.reshape(B, D // 2, 2, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 2, 4, 6, 7)
                return y.reshape(B, D // 2, H // 2, W // 2, 8 * C)
            y = x.reshape(B, D, H, W, 2, 2, 2, C // 8).permute(0, 1, 4, 2, 5, 3, 6, 7)
            return y.reshape(B, 2 * D, 2 * H, 2 * W, C // 8)

    class SwinMergeFn:
        """PatchMerging's gather: v0.9 offsets with the duplicated slices (swin_unetr.py:717-727) or V2's product order"""
        @staticmethod
        def apply(x, v2):
            B, D, H, W, C = x.shape
            x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2, 0, D % 2))
            offs = ([(i, j, k) for i in range(2) for j in range(2) for k in range(2)] if v2 else
                    [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (0, 0, 1), (1, 1, 1)])
            return torch.cat([x[:, i::2, j::2, k::2, :] for i, j, k in offs], -1)

    class ResOutFn:
        @staticmethod
        def apply(r2, st2, r3, st3, act):
            n = lambda t: F.instance_norm(t.permute(0, 4, 1, 2, 3), eps=sw.IN_EPS).permute(0, 2, 3, 4, 1)     # noqa: E731
            s = n(r2) + (n(r3) if st3 is not None else r3)
            return F.leaky_relu(s, 0.01) if act == sw.ACT_LRELU else F.relu(s)

    class WindowAttnFn:
        """forward_part1 between the two Linears (swin_unetr.py:554-606, 467-490) on the qkv tensor of the REAL tokens: a
        padding token's q / k / v is the qkv bias, because the reference pads before the Linear."""
        @staticmethod
        def apply(qkv, qkv_bias, table, heads, window, shift):
            B, D, H, W, C3 = qkv.shape
            C = C3 // 3
            ws, ss = so.get_window_size((D, H, W), window, shift)
            Dp, Hp, Wp = [-(-s // k) * k for s, k in zip((D, H, W), ws)]
            fill = qkv_bias if qkv_bias is not None else torch.zeros(C3, dtype=qkv.dtype)
            x = fill.to(qkv.dtype).expand(B, Dp, Hp, Wp, C3).clone()
            x[:, :D, :H, :W] = qkv
            shifted = any(s > 0 for s in ss)
            mask = None
            if shifted:
                x = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
                mask = so.compute_mask([Dp, Hp, Wp], ws, ss).to(qkv.dtype)
            xw = so.window_partition(x, ws)                                            # [B*nW, n, 3C]
            b_, n, _ = xw.shape
            q, k, v = xw.reshape(b_, n, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
            attn = (q * (C // heads) ** -0.5) @ k.transpose(-2, -1)
            rel = so.relative_position_index(window)                                   # FULL window table, sliced [:n,:n] (:474)
            attn = attn + table[rel[:n, :n].reshape(-1)].reshape(n, n, -1).permute(2, 0, 1).unsqueeze(0).to(attn.dtype)
            if mask is not None:
                nw = mask.shape[0]
                attn = (attn.view(b_ // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, n, n)
            out = (F.softmax(attn, dim=-1) @ v).transpose(1, 2).reshape(b_, n, C)
            y = so.window_reverse(out, ws, [B, Dp, Hp, Wp])
            if shifted:
                y = torch.roll(y, shifts=(ss[0], ss[1], ss[2]), dims=(1, 2, 3))
            return y[:, :D, :H, :W, :].contiguous()

    class LayerNormFn:
        @staticmethod
        def apply(x, gamma, beta, eps):
            return F.layer_norm(x, x.shape[-1:], gamma, beta, eps)

    class GeluFn:
        @staticmethod
        def apply(x):
            return F.gelu(x)

    for name, cls in dict(DepthSpaceFn=DepthSpaceFn, SwinMergeFn=SwinMergeFn, ResOutFn=ResOutFn, WindowAttnFn=WindowAttnFn,
                          LayerNormFn=LayerNormFn, GeluFn=GeluFn).items():
        monkeypatch.setattr(sw, name, cls)
    monkeypatch.setattr(sw, "_need_cuda", lambda t: None)
    return sw
