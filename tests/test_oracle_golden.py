"""CPU: the oracle restatement reproduces the fixtures generated from the real reference
(oracle/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import pytest
import torch

from oracle import losses as olosses
from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import load_golden, rel_err


@pytest.mark.parametrize("name", ["resunet_iso", "resunet_acdc", "unet_single"])
def test_unet_oracle_matches_reference_fixture(name):
    g = load_golden(name)
    cfg = g["cfg"]
    shapes = ounet.unet_param_shapes(1, cfg["base"], cfg["classes"], cfg["kernel"], cfg["block"])
    assert list(shapes) == g["keys"]
    sd = {k: v.requires_grad_(True) for k, v in ounet.make_state_dict(shapes, seed=cfg["state_seed"]).items()}
    img, lab = make_volume(*cfg["shape"], cfg["classes"], seed=cfg["data_seed"])
    torch.set_num_threads(8)
    logits = ounet.unet_forward(sd, img, cfg["scale"], cfg["kernel"], cfg["block"])
    w = torch.tensor(cfg["ce_weight"], dtype=torch.float32)
    loss = olosses.total_loss(logits, lab, w)
    loss.backward()
    assert rel_err(logits, g["logits"].float()) < 2e-3          # fixture logits are stored in fp16
    assert torch.equal(logits.argmax(1).to(torch.uint8), g["argmax"])
    assert abs(loss.item() - g["loss"]) < 1e-5
    for k in g["grad_small"]:
        assert rel_err(sd[k].grad, g["grad_small"][k]) < 1e-4, k
    for k, d in g["grad_digest"].items():
        gs = sd[k].grad.double()
        assert abs(gs.abs().sum().item() - d["abs"]) <= 1e-4 * d["abs"] + 1e-12, k


@pytest.mark.parametrize("name", ["loss_a", "loss_b", "loss_c"])
def test_loss_oracle_matches_reference_fixture(name):
    g = load_golden(name)
    x = g["x"].clone().requires_grad_(True)
    d = olosses.dice_loss(x, g["y"])
    c = olosses.cross_entropy(x, g["y"], g["w"])
    (d + c).backward()
    assert abs(d.item() - g["dice"]) < 1e-6
    assert abs(c.item() - g["ce"]) < 1e-6
    assert rel_err(x.grad, g["grad"]) < 1e-5


def test_dice_alpha_is_differentiable():
    """Detaching alpha changes the gradient — the oracle must keep it attached."""
    g = load_golden("loss_a")
    x = g["x"].clone().requires_grad_(True)
    olosses.dice_loss(x, g["y"]).backward()
    ga = x.grad.clone()
    # finite-difference check of one logit
    with torch.no_grad():
        e = torch.zeros_like(x)
        e.view(-1)[17] = 1e-3
        xd = x.detach().double()
        fd = (olosses.dice_loss(xd + e.double(), g["y"]) - olosses.dice_loss(xd - e.double(), g["y"])) / 2e-3
    assert abs(fd.item() - ga.view(-1)[17].item()) < 5e-3 * abs(fd.item()) + 1e-9


@pytest.mark.parametrize("name", ["biattn_a", "biattn_b", "biattn_c", "biattn_d"])
def test_biattn_oracle_matches_reference_fixture(name):
    """fixtures captured at the projection boundaries of the unmodified BidirectionAttention module."""
    from oracle import medformer_ops as mops
    g = load_golden(name)
    fqv = g["fqv"].clone().requires_grad_(True)
    mqv = g["mqv"].clone().requires_grad_(True)
    fo, mo = mops.bidirection_attention_core(*fqv.chunk(2, 1), *mqv.chunk(2, 1), g["heads"])
    assert rel_err(fo, g["fo"]) < 1e-5 and rel_err(mo, g["mo"]) < 1e-5
    torch.autograd.backward([fo, mo], [g["dfo"], g["dmo"]])
    assert rel_err(fqv.grad, g["dfqv"]) < 1e-4 and rel_err(mqv.grad, g["dmqv"]) < 1e-4


@pytest.mark.parametrize("name", ["dwconv_a", "dwconv_b"])
def test_dwconv_oracle_matches_reference_fixture(name):
    from oracle import medformer_ops as mops
    g = load_golden(name)
    x = g["x"].clone().requires_grad_(True)
    w = g["w"].clone().requires_grad_(True)
    y = mops.depthwise_conv3d(x, w)
    y.backward(g["gy"])
    assert rel_err(y, g["y"]) < 1e-6 and rel_err(x.grad, g["dx"]) < 1e-5 and rel_err(w.grad, g["dw"]) < 1e-5


@pytest.mark.parametrize("name", ["medformer_bcv", "medformer_var", "medformer_acdc"])
def test_medformer_oracle_matches_reference_fixture(name):
    from oracle import medformer as omed
    from oracle.unet3d import make_state_dict
    g = load_golden(name)
    cfg = g["cfg"]
    sd = make_state_dict(g["shapes"], seed=cfg["state_seed"])
    for k in sd:
        if k.endswith("norm.weight"):
            sd[k] = 1.0 + 0.1 * sd[k] / sd[k].abs().max()
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    img, lab = make_volume(*cfg["shape"], cfg["classes"], seed=cfg["data_seed"])
    torch.set_num_threads(8)
    res = omed.medformer_forward(sd, img, cfg)
    w = torch.tensor(cfg["ce_weight"])
    loss = olosses.total_loss(res, lab, w, cfg["aux_weight"]) if isinstance(res, list) else olosses.total_loss(res, lab, w)
    loss.backward()
    for r, ref, am in zip(res if isinstance(res, list) else [res], g["logits"], g["argmax"]):
        assert rel_err(r, ref.float()) < 2e-3 and torch.equal(r.argmax(1).to(torch.uint8), am)
    assert abs(loss.item() - g["loss"]) < 1e-5
    gmax = max(d["abs"] / max(sd[k].numel(), 1) for k, d in g["grad_digest"].items())
    for k, ref in g["grad_small"].items():
        assert ((sd[k].grad - ref).abs().max() / (ref.abs().max() + 1e-2 * gmax)).item() < 1e-3, k


def test_medformer_state_dict_contract():
    """b200seg.MedFormer registers exactly the reference's parameters (names, shapes, order) — checked against
    the key list captured from the unmodified reference module."""
    import b200seg
    g = load_golden("medformer_bcv")
    cfg = g["cfg"]
    kw = {k: cfg[k] for k in ("map_size", "conv_num", "trans_num", "num_heads", "fusion_depth", "fusion_dim",
                              "fusion_heads", "kernel_size", "scale", "aux_loss")}
    net = b200seg.MedFormer(1, cfg["classes"], 32, conv_block="BasicBlock", expansion=4, attn_drop=0, proj_drop=0,
                            proj_type="depthwise", norm="in", act="relu", **kw)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == list(g["shapes"].items())
    assert len(list(net.buffers())) == 0
    assert sum(p.numel() for p in net.parameters()) == 36669180
    with pytest.raises(ValueError):
        b200seg.MedFormer(1, 14, 32, norm="bn")
    with pytest.raises(b200seg.B200SegError):
        net(torch.zeros(1, 1, 16, 32, 32))          # CPU tensor: loud failure, no fallback


def _req(d):
    return {k: v.clone().requires_grad_(True) for k, v in d.items()}


@pytest.mark.parametrize("name", ["swin_attn_a", "swin_attn_b", "swin_attn_c"])
def test_swin_window_attention_oracle_matches_reference_fixture(name):
    """WindowAttention.forward of the vendored swin_unetr.py (rel-pos bias, shift mask) — rows a15 of SURVEY §8."""
    from oracle import swin_ops as so
    g = load_golden(name)
    cfg = g["cfg"]
    mask = None if g["mask_cross"] is None else g["mask_cross"].float() * -100.0
    x, p = g["x"].clone().requires_grad_(True), _req(g["params"])
    y = so.window_attention(x, p, cfg["heads"], so.relative_position_index(cfg["window"]), mask)
    y.backward(g["gy"])
    assert rel_err(y, g["y"]) < 1e-5 and rel_err(x.grad, g["dx"]) < 1e-4
    for k in p:
        assert rel_err(p[k].grad, g["dparams"][k]) < 1e-4, k


@pytest.mark.parametrize("name", ["swin_block_a", "swin_block_b", "swin_block_c"])
def test_swin_block_part1_oracle_matches_reference_fixture(name):
    """LayerNorm + pad + cyclic shift + window partition + attention + reverse (SwinTransformerBlock.forward_part1)."""
    from oracle import swin_ops as so
    g = load_golden(name)
    cfg = g["cfg"]
    ws, ss = so.get_window_size(cfg["dhw"], cfg["window"], cfg["shift"])
    pdims = [-(-cfg["dhw"][i] // ws[i]) * ws[i] for i in range(3)]
    mask = so.compute_mask(pdims, ws, ss) if any(ss) else None
    x, p = g["x"].clone().requires_grad_(True), _req(g["params"])
    y = so.swin_block_part1(x, p, cfg["heads"], cfg["window"], cfg["shift"], mask)
    y.backward(g["gy"])
    assert rel_err(y, g["y"]) < 1e-5 and rel_err(x.grad, g["dx"]) < 1e-4
    for k in p:
        assert rel_err(p[k].grad, g["dparams"][k]) < 1e-4, k


@pytest.mark.parametrize("name,fn", [("swin_merge_a", "patch_merging"), ("swin_merge_b", "patch_merging_v2")])
def test_swin_patch_merging_oracle_matches_reference_fixture(name, fn):
    from oracle import swin_ops as so
    g = load_golden(name)
    x = g["x"].clone().requires_grad_(True)
    y = getattr(so, fn)(x, g["norm_w"], g["norm_b"], g["red_w"])
    y.backward(g["gy"])
    assert rel_err(y, g["y"]) < 1e-5 and rel_err(x.grad, g["dx"]) < 1e-5
