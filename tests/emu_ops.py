"""CPU emulation of the C-ABI op semantics in plain PyTorch (TEST INFRASTRUCTURE).

Lets the `-m "not gpu"` suite drive the package's host-side orchestration (autograd Functions, fused-block
backward algebra, channel-slice bookkeeping, weight packing order) without a GPU, by monkeypatching the
launch wrappers in b200seg.ops.  The product never imports this file."""
import torch
import torch.nn.functional as F


def _ncdhw(t, coff, C):
    return t[..., coff:coff + C].permute(0, 4, 1, 2, 3)


def _stats(x_ncdhw):
    xd = x_ncdhw.double().flatten(2)
    return torch.stack([xd.sum(-1), (xd * xd).sum(-1)], -1).contiguous()


def _mean_rstd(stats, n, eps):
    m = stats[..., 0] / n
    var = (stats[..., 1] / n - m * m).clamp_min(0)
    return m, 1.0 / torch.sqrt(var + eps)


def _normalise(x_ncdhw, stats, act, eps):
    if stats is None:
        return x_ncdhw
    n = x_ncdhw[0, 0].numel()
    m, r = _mean_rstd(stats, n, eps)
    h = (x_ncdhw.double() - m[:, :, None, None, None]) * r[:, :, None, None, None]
    return _act(h, act).to(x_ncdhw.dtype)


def _act(h, act):
    """B200SEG_ACT_NONE / RELU / LRELU(0.01)"""
    return F.relu(h) if act == 1 else (F.leaky_relu(h, 0.01) if act == 2 else h)


def _act_grad(h, act):
    one = torch.ones_like(h)
    return (h > 0).to(h.dtype) if act == 1 else (torch.where(h > 0, one, 0.01 * one) if act == 2 else one)


def install(monkeypatch):
    from b200seg import ops

    def conv_algo(Cin, Cout, ksize, dtype, B=1):
        return ops.ALGO_DIRECT

    def pack_weight(w, dtype, transpose_flip=False, out=None, co_off=0, co_total=None, layout=ops.ALGO_DIRECT):
        Cout, Cin = w.shape[:2]
        taps = w[0, 0].numel()
        co_total = Cout if co_total is None else co_total
        if out is None:
            out = torch.zeros(taps * co_total * Cin, dtype=dtype)
        wt = w.detach().reshape(Cout, Cin, taps).to(dtype)
        if not transpose_flip:
            v = out.view(taps, co_total, Cin)
            v[:, co_off:co_off + Cout, :] = wt.permute(2, 0, 1)
        else:
            v = out.view(taps, Cin, co_total)
            v[:, :, co_off:co_off + Cout] = wt.flip(2).permute(2, 1, 0)
        return out

    def conv3d_fwd(x, x_coff, Cin, x_stats, act, wp, Cout, ksize, bias=None, residual=None, r_coff=0,
                   want_stats=True, dgrad_of=None, algo=None, eps=ops.IN_EPS):
        if isinstance(wp, tuple):
            wp = wp[0]
        taps = ksize[0] * ksize[1] * ksize[2]
        w = wp.view(taps, Cout, Cin).permute(1, 2, 0).reshape(Cout, Cin, *ksize)
        a = _normalise(_ncdhw(x, x_coff, Cin), x_stats, act, eps)
        y = F.conv3d(a.double(), w.double(), padding=[k // 2 for k in ksize])
        if bias is not None:
            y = y + bias.double()[None, :, None, None, None]
        st = None
        if dgrad_of is not None:
            gx, gcoff, gstats, gact = dgrad_of
            xg = _ncdhw(gx, gcoff, Cout)
            n = xg[0, 0].numel()
            m, r = _mean_rstd(gstats, n, eps)
            hx = (xg.double() - m[:, :, None, None, None]) * r[:, :, None, None, None]
            if gact:
                y = y * _act_grad(hx, gact)
            if want_stats:
                st = torch.stack([y.flatten(2).sum(-1), (y * hx).flatten(2).sum(-1)], -1).contiguous()
        else:
            if residual is not None:
                y = y + _ncdhw(residual, r_coff, Cout).double()
            if want_stats:
                st = _stats(y)
        return y.to(x.dtype).permute(0, 2, 3, 4, 1).contiguous(), st

    def conv3d_wgrad(x, x_coff, Cin, x_stats, act, dy, dy_coff, Cout, ksize, want_bias=False, algo=0, eps=ops.IN_EPS):
        a = _normalise(_ncdhw(x, x_coff, Cin), x_stats, act, eps).double()
        g = _ncdhw(dy, dy_coff, Cout).double()
        w = torch.zeros(Cout, Cin, *ksize, dtype=torch.float64, requires_grad=True)
        with torch.enable_grad():
            F.conv3d(a, w, padding=[k // 2 for k in ksize]).backward(g)
        db = g.sum((0, 2, 3, 4)).float() if want_bias else None
        return w.grad.float(), db

    def in_bwd_apply(g, x, x_coff, C, x_stats, bstats, add=None, add_coff=0, out=None, out_coff=0, eps=ops.IN_EPS):
        xs = _ncdhw(x, x_coff, C).double()
        n = xs[0, 0].numel()
        m, r = _mean_rstd(x_stats, n, eps)
        hx = (xs - m[:, :, None, None, None]) * r[:, :, None, None, None]
        gg = _ncdhw(g, 0, C).double()
        dx = r[:, :, None, None, None] * (gg - (bstats[..., 0] / n)[:, :, None, None, None]
                                          - hx * (bstats[..., 1] / n)[:, :, None, None, None])
        if add is not None:
            dx = dx + _ncdhw(add, add_coff, C).double()
        dx = dx.permute(0, 2, 3, 4, 1).to(x.dtype)
        if out is None:
            return dx.contiguous()
        out[..., out_coff:out_coff + C] = dx
        return out

    def copy_channels(x, x_coff, y, y_coff, C, accumulate=False):
        if accumulate:
            y[..., y_coff:y_coff + C] += x[..., x_coff:x_coff + C].to(y.dtype)
        else:
            y[..., y_coff:y_coff + C] = x[..., x_coff:x_coff + C].to(y.dtype)
        return y

    def in_apply(x, C, stats, act, eps=ops.IN_EPS):
        return _normalise(_ncdhw(x, 0, C), stats, act, eps).permute(0, 2, 3, 4, 1).contiguous()

    def in_bwd_reduce(dy, x, C, stats, act, eps=ops.IN_EPS):
        xs = _ncdhw(x, 0, C).double()
        m, r = _mean_rstd(stats, xs[0, 0].numel(), eps)
        hx = (xs - m[:, :, None, None, None]) * r[:, :, None, None, None]
        g = _ncdhw(dy, 0, C).double()
        if act:
            g = g * _act_grad(hx, act)
        bst = torch.stack([g.flatten(2).sum(-1), (g * hx).flatten(2).sum(-1)], -1).contiguous()
        return g.to(x.dtype).permute(0, 2, 3, 4, 1).contiguous(), bst

    def instnorm_stats(x, x_coff, C):
        return _stats(_ncdhw(x, x_coff, C))

    class MaxPoolFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, scale, want_stats):
            xn = x.permute(0, 4, 1, 2, 3)
            y, idx = F.max_pool3d(xn, scale, return_indices=True)
            ctx.save_for_backward(idx)
            ctx.meta = (xn.shape, scale)
            st = _stats(y)
            ctx.mark_non_differentiable(st)
            return y.permute(0, 2, 3, 4, 1).contiguous(), st

        @staticmethod
        def backward(ctx, dy, _):
            (idx,) = ctx.saved_tensors
            shp, scale = ctx.meta
            dx = F.max_unpool3d(dy.permute(0, 4, 1, 2, 3).contiguous(), idx, scale, output_size=shp[2:])
            return dx.permute(0, 2, 3, 4, 1).contiguous(), None, None

    class UpCatFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, low, skip, skip_stats, skip_first):
            ln = low.permute(0, 4, 1, 2, 3)
            up = F.interpolate(ln, size=skip.shape[1:4], mode="trilinear", align_corners=True).permute(0, 2, 3, 4, 1)
            if skip_stats is None or skip_stats.numel() == 0:
                skip_stats = _stats(skip.permute(0, 4, 1, 2, 3))
            ust = _stats(up.permute(0, 4, 1, 2, 3))
            ctx.meta = (low.shape, skip.shape[-1], skip.shape[1:4])
            cat = torch.cat([skip, up] if skip_first else [up, skip], -1).contiguous()
            st = torch.cat([skip_stats, ust] if skip_first else [ust, skip_stats], 1).contiguous()
            ctx.skip_first = skip_first
            ctx.mark_non_differentiable(st)
            return cat, st

        @staticmethod
        def backward(ctx, d_cat, _):
            lshape, Cs, size = ctx.meta
            Cl = lshape[-1]
            ds, du = (d_cat[..., :Cs], d_cat[..., Cs:]) if ctx.skip_first else (d_cat[..., Cl:], d_cat[..., :Cl])
            low = torch.zeros(lshape, dtype=torch.float64, requires_grad=True)
            with torch.enable_grad():
                up = F.interpolate(low.permute(0, 4, 1, 2, 3), size=size, mode="trilinear", align_corners=True)
                up.backward(du.permute(0, 4, 1, 2, 3).double())
            return low.grad.to(d_cat.dtype), ds.contiguous(), None, None

    class DiceCEFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, logits, labels, weight, ce_scale, dice_scale):
            from oracle import losses as ol
            with torch.enable_grad():
                x = logits.detach().double().requires_grad_(True)
                lab = labels.view(labels.shape[0], 1, *logits.shape[2:]).long()
                loss = ce_scale * F.cross_entropy(x, lab.squeeze(1), weight=None if weight is None else weight.double()) \
                    + dice_scale * ol.dice_loss(x, lab)
                loss.backward()
            ctx.g = x.grad
            return loss.detach().float()

        @staticmethod
        def backward(ctx, g):
            return (ctx.g * g).float(), None, None, None, None

    for name, fn in dict(conv_algo=conv_algo, pack_weight=pack_weight, conv3d_fwd=conv3d_fwd, conv3d_wgrad=conv3d_wgrad,
                         in_bwd_apply=in_bwd_apply, in_apply=in_apply, in_bwd_reduce=in_bwd_reduce, copy_channels=copy_channels, instnorm_stats=instnorm_stats,
                         MaxPoolFn=MaxPoolFn, UpCatFn=UpCatFn, DiceCEFn=DiceCEFn).items():
        monkeypatch.setattr(ops, name, fn)
    import b200seg.unet3d as u
    import b200seg.losses as lo
    monkeypatch.setattr(u, "MaxPoolFn", MaxPoolFn)
    monkeypatch.setattr(u, "UpCatFn", UpCatFn)
    monkeypatch.setattr(lo, "DiceCEFn", DiceCEFn)
    monkeypatch.setattr(ops, "_need_cuda", lambda t: None)

    class _FakeCuda:
        pass
    return ops
