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
