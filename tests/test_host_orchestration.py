"""CPU: the package's autograd orchestration (fused BasicBlock backward algebra, channel-slice bookkeeping,
stats plumbing, weight packing order) driven end-to-end with the C-ABI ops emulated in PyTorch
(tests/emu_ops.py) and compared with the reference-pinned oracle."""
import pytest
import torch

import b200seg
from oracle import losses as olosses
from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import global_l2, grad_noise_floor, load_golden, rel_err
import emu_ops


@pytest.mark.parametrize("name", ["resunet_iso", "resunet_acdc", "unet_single"])
def test_orchestration_matches_oracle(monkeypatch, name):
    emu_ops.install(monkeypatch)
    g = load_golden(name)
    cfg = g["cfg"]
    net = b200seg.UNet(1, cfg["base"], scale=cfg["scale"], kernel_size=cfg["kernel"], num_classes=cfg["classes"],
                       block=cfg["block"], norm="in")
    shapes = ounet.unet_param_shapes(1, cfg["base"], cfg["classes"], cfg["kernel"], cfg["block"])
    sd = ounet.make_state_dict(shapes, seed=cfg["state_seed"])
    net.load_state_dict(sd)
    img, lab = make_volume(*cfg["shape"], cfg["classes"], seed=cfg["data_seed"])
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    logits = net(img)
    w = torch.tensor(cfg["ce_weight"])
    loss = b200seg.DiceCELoss(weight=w)(logits, lab)
    loss.backward()
    g64, l64, floor_max, floor_l2 = grad_noise_floor(sd, img, lab, w, cfg)
    assert rel_err(logits, l64) < 1e-4
    assert torch.equal(logits.argmax(1), l64.argmax(1))
    ours = {k: p.grad for k, p in net.named_parameters()}
    errs = {k: rel_err(ours[k], g64[k]) for k in g64}
    # the emulated ops compute in fp64 but hand fp32 tensors to each other, so (like any fp32 evaluation) the
    # result sits within the reference's own fp32 noise floor of the fp64 answer
    assert max(errs.values()) < max(1e-3, 3 * floor_max), sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    assert global_l2(ours, g64) < max(1e-3, 3 * floor_l2)


@pytest.mark.parametrize("name", ["medformer_bcv", "medformer_var"])
def test_medformer_orchestration_matches_oracle(monkeypatch, name):
    """b200seg.MedFormer's module wiring (which norm eps where, which tensors carry IN sums, residual routing, padded
    output channels, [up, skip] concat order, token layouts, deep-supervision head) with every C-ABI op emulated."""
    import emu_medformer
    from oracle import medformer as omed
    from oracle.unet3d import make_state_dict
    emu_medformer.install(monkeypatch)
    g = load_golden(name)
    cfg = g["cfg"]
    kw = {k: cfg[k] for k in ("map_size", "conv_num", "trans_num", "num_heads", "fusion_depth", "fusion_dim",
                              "fusion_heads", "kernel_size", "scale", "aux_loss")}
    net = b200seg.MedFormer(1, cfg["classes"], 32, conv_block="BasicBlock", expansion=4, attn_drop=0, proj_drop=0,
                            proj_type="depthwise", norm="in", act="relu", **kw)
    sd = make_state_dict(g["shapes"], seed=cfg["state_seed"])
    for k in sd:
        if k.endswith("norm.weight"):
            sd[k] = 1.0 + 0.1 * sd[k] / sd[k].abs().max()
    net.load_state_dict(sd)
    img, lab = make_volume(*cfg["shape"], cfg["classes"], seed=cfg["data_seed"])
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    res = net(img)
    outs = res if isinstance(res, list) else [res]
    w = torch.tensor(cfg["ce_weight"])
    crit = b200seg.DiceCELoss(weight=w)
    loss = sum(cfg["aux_weight"][j] * crit(r, lab) for j, r in enumerate(outs)) if len(outs) > 1 else crit(outs[0], lab)
    loss.backward()
    for o, ref in zip(outs, g["logits"]):
        assert o.shape == ref.shape and rel_err(o, ref.float()) < 2e-3          # fixture stored in fp16
    assert abs(loss.item() - g["loss"]) < 1e-4
    s64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    r64 = omed.medformer_forward(s64, img.double(), kw)
    l64 = olosses.total_loss(r64, lab, w.double(), cfg["aux_weight"]) if isinstance(r64, list) else olosses.total_loss(r64, lab, w.double())
    l64.backward()
    g64 = {k: v.grad for k, v in s64.items()}
    ours = {k: p.grad for k, p in net.named_parameters()}
    assert all(v is not None for v in ours.values()), [k for k, v in ours.items() if v is None]
    err = global_l2(ours, g64)
    print("%s emulated-orchestration grad L2 err vs fp64 oracle: %.2e" % (name, err))
    assert err < 5e-2       # fp32 hand-offs between emulated ops; the GPU test holds the tight bar
