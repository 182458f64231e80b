"""GPU: SwinUNETR path (rows a15-a17) against fixtures produced by the reference's VENDORED classes
(oracle/make_golden_swin.py, oracle/make_golden_swin_unetr.py):
  * SwinTransformerBlock.forward_part1 — LayerNorm, padding to a window multiple, cyclic shift, window partition,
    attention with relative-position bias and shift mask, reverse, un-shift, crop — including a window clamped on a
    short axis (the relative_position_index[:n,:n] quirk) — forward and every gradient;
  * both PatchMerging variants (the v0.9 one with its duplicated slices), odd extents;
  * the whole SwinUNETR: state_dict contract, logits / argmax / loss vs the fixture, gradients vs the oracle in fp64."""
import pytest
import torch

from oracle import losses as olosses
from oracle import swin_unetr as osw
from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import global_l2, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _lin(weight, bias):
    lin = torch.nn.Linear(weight.shape[1], weight.shape[0], bias=bias is not None)
    with torch.no_grad():
        lin.weight.copy_(weight)
        if bias is not None:
            lin.bias.copy_(bias)
    return lin.cuda()


@pytest.mark.parametrize("name", ["swin_block_a", "swin_block_b", "swin_block_c"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_swin_block_part1_matches_reference(name, dtype):
    from b200seg import ops
    from b200seg.medformer_ops import LayerNormFn
    from b200seg.swin_unetr import WindowAttnFn, _linear
    g = load_golden(name)
    cfg, p = g["cfg"], g["params"]
    heads, ws, ss = cfg["heads"], tuple(cfg["window"]), tuple(cfg["shift"])
    x = g["x"].cuda().to(dtype).requires_grad_(True)
    nw, nb = p["norm1_w"].cuda().requires_grad_(True), p["norm1_b"].cuda().requires_grad_(True)
    qkv, proj = _lin(p["qkv_w"], p["qkv_b"]), _lin(p["proj_w"], p["proj_b"])
    table = p["bias_table"].cuda().requires_grad_(True)
    pq, pp = ops.PackedWeights(), ops.PackedWeights()
    xn = LayerNormFn.apply(x, nw, nb, 1e-5)
    att = WindowAttnFn.apply(_linear(pq, xn, qkv), qkv.bias, table, heads, ws, ss)
    y = _linear(pp, att, proj)
    y.backward(g["gy"].cuda().to(dtype))
    tol = 2e-4 if dtype == torch.float32 else 2e-2
    assert rel_err(y, g["y"]) < tol
    assert rel_err(x.grad, g["dx"]) < tol * 5
    got = {"norm1_w": nw.grad, "norm1_b": nb.grad, "qkv_w": qkv.weight.grad, "qkv_b": qkv.bias.grad,
           "proj_w": proj.weight.grad, "proj_b": proj.bias.grad, "bias_table": table.grad}
    errs = {k: rel_err(got[k], g["dparams"][k]) for k in got}
    print(name, dtype, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) < tol * 5, errs


@pytest.mark.parametrize("name,v2", [("swin_merge_a", False), ("swin_merge_b", True)])
def test_patch_merging_matches_reference(name, v2):
    from b200seg import ops
    from b200seg.medformer_ops import LayerNormFn
    from b200seg.swin_unetr import SwinMergeFn, _linear
    g = load_golden(name)
    x = g["x"].cuda().requires_grad_(True)
    red = _lin(g["red_w"], None)
    y = _linear(ops.PackedWeights(), LayerNormFn.apply(SwinMergeFn.apply(x, v2), g["norm_w"].cuda(), g["norm_b"].cuda(), 1e-5), red)
    y.backward(g["gy"].cuda())
    assert rel_err(y, g["y"]) < 1e-4
    assert rel_err(x.grad, g["dx"]) < 1e-4


def _build(cfg):
    import b200seg
    net = b200seg.SwinUNETR(cfg["size"], cfg["in_ch"], cfg["classes"], feature_size=cfg["feature_size"])
    return net


def test_swin_unetr_state_dict_contract():
    g = load_golden("swin_unetr_small")
    net = _build(g["cfg"])
    keys = [k for k in net.state_dict() if not k.endswith("relative_position_index")]
    assert keys == list(g["shapes"])
    for k in keys:
        assert tuple(net.state_dict()[k].shape) == tuple(g["shapes"][k]), k
    assert sum(1 for k in net.state_dict() if k.endswith("relative_position_index")) == 6


@pytest.mark.parametrize("amp", [False, True])
def test_swin_unetr_forward_backward(amp):
    import b200seg
    g = load_golden("swin_unetr_small")
    cfg = g["cfg"]
    net = _build(cfg)
    sd = ounet.make_state_dict(g["shapes"], seed=cfg["state_seed"])
    for k in sd:
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k.endswith("norm.weight"):
            sd[k] = 1.0 + 0.1 * sd[k] / sd[k].abs().max()
        if k.endswith("relative_position_bias_table"):
            sd[k] = sd[k] * 3.0
    missing = net.load_state_dict(sd, strict=False)
    assert all(k.endswith("relative_position_index") for k in missing.missing_keys) and not missing.unexpected_keys
    net = net.cuda()
    img, lab = make_volume(1, *cfg["size"], cfg["classes"], seed=cfg["data_seed"], in_ch=cfg["in_ch"])
    w = torch.tensor(cfg["ce_weight"])
    S = 1024.0 if amp else 1.0
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        logits = net(img.cuda())
        loss = b200seg.DiceCELoss(weight=w)(logits, lab.cuda())
    (loss * S).backward()
    lg = logits.detach().float().cpu()
    e = rel_err(lg, g["logits"].float())
    agree = (lg.argmax(1).to(torch.uint8) == g["argmax"]).float().mean().item()
    # exact answer: the reference-pinned oracle in fp64 on this GPU
    so = {k: v.double().cuda().requires_grad_(True) for k, v in sd.items()}
    lo = osw.swin_unetr_forward(so, img.double().cuda())
    olosses.total_loss(lo, lab.cuda(), w.double().cuda()).backward()
    g64 = {k: v.grad.cpu() for k, v in so.items()}
    ours = {k: (p.grad / S).double().cpu() for k, p in net.named_parameters()}
    assert set(ours) == set(g64)
    l2 = global_l2(ours, g64)
    worst = sorted(((rel_err(ours[k], g64[k]), k) for k in g64), reverse=True)[:4]
    print("swin_unetr amp=%d: logits rel err %.2e, label agreement %.5f, loss %.5f (ref %.5f), grads global-L2 %.2e, worst %s"
          % (amp, e, agree, loss.item(), g["loss"], l2, [(k, "%.1e" % v) for v, k in worst]))
    if amp:
        assert e < 5e-2 and agree > 0.97 and abs(loss.item() - g["loss"]) < 3e-2 and l2 < 0.25
    else:
        assert e < 2e-3 and agree > 0.9995 and abs(loss.item() - g["loss"]) < 1e-4 and l2 < 5e-3


@pytest.mark.parametrize("dims,heads,dh,window,shift", [((9, 14, 8), 3, 16, (7, 7, 7), (3, 3, 3)),      # SwinUNETR fs48 head size, clamped + shifted
                                                        ((7, 7, 7), 2, 16, (7, 7, 7), (0, 0, 0)),      # one full window, no mask
                                                        ((5, 6, 9), 2, 32, (4, 4, 4), (2, 2, 2)),
                                                        ((8, 8, 8), 3, 8, (4, 4, 4), (2, 0, 2))])
def test_window_attention_mma_path(dims, heads, dh, window, shift):
    """The tensor-core kernels (swin_mma.cu, fp16) against the CUDA-core kernels (swin.cu) — which the block fixtures
    above pin to the reference — on the same inputs: fp16 CUDA-core (same rounding of the inputs) and fp32."""
    import os
    from b200seg.swin_unetr import WindowAttnFn
    g = torch.Generator().manual_seed(dh * 100 + heads)
    C = heads * dh
    B = 2
    qkv = (torch.randn(B, *dims, 3 * C, generator=g) * 0.8)
    bias = torch.randn(3 * C, generator=g) * 0.3
    table = torch.randn((2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1), heads, generator=g) * 0.5
    dout = torch.randn(B, *dims, C, generator=g)

    def run(dtype, mma):
        os.environ["B200SEG_WINATTN_MMA"] = "1" if mma else "0"
        try:
            x = qkv.cuda().to(dtype).requires_grad_(True)
            b = bias.cuda().requires_grad_(True)
            t = table.cuda().requires_grad_(True)
            y = WindowAttnFn.apply(x, b, t, heads, window, shift)
            y.backward(dout.cuda().to(dtype))
            torch.cuda.synchronize()
            return [v.detach().float().cpu() for v in (y, x.grad, t.grad, b.grad)]
        finally:
            os.environ.pop("B200SEG_WINATTN_MMA", None)
    ref32 = run(torch.float32, False)
    cc16 = run(torch.float16, False)
    mma16 = run(torch.float16, True)
    names = ("out", "dqkv", "dtable", "dbias")
    for nme, a, c, r in zip(names, mma16, cc16, ref32):
        e_ref, e_cc, floor = rel_err(a, r), rel_err(a, c), rel_err(c, r)
        print("win-attn mma %s dh=%d: vs fp32 %.2e, vs fp16 CUDA-core %.2e (fp16 CUDA-core vs fp32 %.2e)" % (nme, dh, e_ref, e_cc, floor))
        assert e_ref < max(2e-2, 3 * floor), nme
    assert not torch.equal(mma16[0], cc16[0]) or dh not in (8, 16, 32)       # the two paths really are different kernels
