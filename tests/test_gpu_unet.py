"""GPU: whole-model parity of b200seg.UNet against the reference fixtures / reference-pinned oracle:
bit-exact argmax label maps, <=1e-3 rel logits and gradients in fp32 (north_star)."""
import pytest
import torch

from oracle import losses as olosses
from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import dice_per_class, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _build(cfg):
    import b200seg
    net = b200seg.UNet(1, cfg["base"], scale=cfg["scale"], kernel_size=cfg["kernel"], num_classes=cfg["classes"],
                       block=cfg["block"], norm="in")
    shapes = ounet.unet_param_shapes(1, cfg["base"], cfg["classes"], cfg["kernel"], cfg["block"])
    sd = ounet.make_state_dict(shapes, seed=cfg["state_seed"])
    net.load_state_dict(sd)
    return net.cuda(), sd


@pytest.mark.parametrize("name", ["resunet_iso", "resunet_acdc", "unet_single"])
def test_fp32_forward_backward_matches_reference(name):
    import b200seg
    g = load_golden(name)
    cfg = g["cfg"]
    net, sd = _build(cfg)
    img, lab = make_volume(*cfg["shape"], cfg["classes"], seed=cfg["data_seed"])
    logits = net(img.cuda())
    assert logits.shape == (cfg["shape"][0], cfg["classes"], *cfg["shape"][1:])
    w = torch.tensor(cfg["ce_weight"])
    loss = b200seg.DiceCELoss(weight=w)(logits, lab.cuda())
    loss.backward()
    lg = logits.detach().float().cpu()
    assert rel_err(lg, g["logits"].float()) < 2e-3                       # fixture stored in fp16
    am = lg.argmax(1).to(torch.uint8)
    assert torch.equal(am, g["argmax"]), "argmax label map differs from the reference"
    assert min(dice_per_class(am, g["argmax"], cfg["classes"])) == 1.0   # "Dice vs ref" = 1.0
    assert abs(loss.item() - g["loss"]) < 1e-4
    # full-precision check against the oracle evaluated here (fixture holds digests only)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = ounet.unet_forward(sdo, img, cfg["scale"], cfg["kernel"], cfg["block"])
    olosses.total_loss(lo, lab, w).backward()
    assert rel_err(lg, lo) < 1e-3
    errs = {k: rel_err(p.grad, sdo[k].grad) for k, p in net.named_parameters()}
    worst = max(errs.values())
    assert worst < 1e-3, "worst gradient rel errs %s" % (sorted(errs.items(), key=lambda kv: -kv[1])[:8],)
    for k in g["grad_small"]:
        assert rel_err(dict(net.named_parameters())[k].grad, g["grad_small"][k]) < 1e-3


@pytest.mark.parametrize("name", ["resunet_iso", "resunet_acdc"])
def test_amp_forward_backward_close_to_fp32_reference(name):
    """fp16 storage + fp32 accumulate (the --amp path).  fp16 eps ~1e-3 per rounding, so the bar is the
    one SURVEY.md §7 sets: judged against the fp32 oracle with a tolerance an fp16 pipeline can meet."""
    import b200seg
    g = load_golden(name)
    cfg = g["cfg"]
    net, sd = _build(cfg)
    img, lab = make_volume(*cfg["shape"], cfg["classes"], seed=cfg["data_seed"])
    w = torch.tensor(cfg["ce_weight"])
    scaler_scale = 1024.0
    with torch.autocast("cuda", dtype=torch.float16):
        logits = net(img.cuda())
        assert logits.dtype == torch.float16
        loss = b200seg.DiceCELoss(weight=w)(logits, lab.cuda())
    (loss * scaler_scale).backward()
    lg = logits.detach().float().cpu()
    assert rel_err(lg, g["logits"].float()) < 3e-2
    agree = (lg.argmax(1).to(torch.uint8) == g["argmax"]).float().mean().item()
    assert agree > 0.995, agree
    assert abs(loss.item() - g["loss"]) < 2e-2
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = ounet.unet_forward(sdo, img, cfg["scale"], cfg["kernel"], cfg["block"])
    olosses.total_loss(lo, lab, w).backward()
    errs = {k: rel_err(p.grad / scaler_scale, sdo[k].grad) for k, p in net.named_parameters()}
    assert max(errs.values()) < 8e-2, sorted(errs.items(), key=lambda kv: -kv[1])[:3]


def test_eval_no_grad_and_state_dict_roundtrip():
    import b200seg
    cfg = load_golden("resunet_iso")["cfg"]
    net, sd = _build(cfg)
    img, _ = make_volume(*cfg["shape"], cfg["classes"], seed=1)
    net.eval()
    with torch.no_grad():
        a = net(img.cuda())
    net2, _ = _build(cfg)
    net2.load_state_dict(net.state_dict())
    with torch.no_grad():
        b = net2(img.cuda())
    assert torch.equal(a, b)
