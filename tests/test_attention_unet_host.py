"""CPU: b200seg.AttentionUNet's contract (state_dict keys / order / shapes, get_model dispatch) and its module wiring
(upsample -> gate -> cat order, which tensors carry IN sums, unused conv_ch) with every C-ABI op emulated in PyTorch."""
import types

import pytest
import torch
import torch.nn.functional as F

import b200seg
import emu_medformer
from emu_ops import _stats
from oracle import attention_unet as oatt
from oracle import losses as olosses
from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import global_l2, load_golden, rel_err


def test_state_dict_contract_and_factory():
    g = load_golden("attention_unet_small")
    c = g["cfg"]
    args = types.SimpleNamespace(dimension="3d", model="attention_unet", in_chan=1, base_chan=c["base"], classes=c["classes"],
                                 down_scale=c["scale"], norm="in", kernel_size=c["kernel"], block=c["block"])
    net = b200seg.get_model(args)
    assert isinstance(net, b200seg.AttentionUNet)
    sd = net.state_dict()
    assert list(sd.keys()) == list(g["shapes"].keys())
    assert all(tuple(v.shape) == tuple(g["shapes"][k]) for k, v in sd.items())
    assert [k for k, _ in net.named_parameters()] == list(g["shapes"].keys())
    with pytest.raises(ValueError):
        b200seg.AttentionUNet(1, 16, scale=c["scale"], kernel_size=c["kernel"], num_classes=4, block="BasicBlock", norm="bn")
    with pytest.raises(ValueError):      # gate widths must be multiples of 8 (base 8 -> int_ch 4)
        b200seg.AttentionUNet(1, 8, scale=c["scale"], kernel_size=c["kernel"], num_classes=4, block="BasicBlock", norm="in")


def _install(monkeypatch):
    emu_medformer.install(monkeypatch)
    from b200seg import attention_unet as au

    def _cl(t):
        return t.permute(0, 2, 3, 4, 1).contiguous()

    class UpsampleStatsFn:
        @staticmethod
        def apply(x, size):
            y = _cl(F.interpolate(x.permute(0, 4, 1, 2, 3), size=size, mode="trilinear", align_corners=True))
            return y, _stats(y.detach().permute(0, 4, 1, 2, 3))

    class ResOutFn:
        @staticmethod
        def apply(r2, st2, r3, st3, act):
            assert act == au.ACT_RELU and st3 is not None
            a = F.instance_norm(r2.permute(0, 4, 1, 2, 3), eps=au.GATE_EPS) + F.instance_norm(r3.permute(0, 4, 1, 2, 3), eps=au.GATE_EPS)
            return _cl(F.relu(a))

    class AttnGateFn:
        @staticmethod
        def apply(x, t, w):
            p = F.conv3d(t.permute(0, 4, 1, 2, 3), w)
            y = x * _cl(torch.sigmoid(F.instance_norm(p, eps=au.GATE_EPS)))
            return y, _stats(y.detach().permute(0, 4, 1, 2, 3))

    for name, cls in dict(UpsampleStatsFn=UpsampleStatsFn, ResOutFn=ResOutFn, AttnGateFn=AttnGateFn).items():
        monkeypatch.setattr(au, name, cls)
    monkeypatch.setattr(au, "_need_cuda", lambda t: None)


def test_orchestration_matches_oracle(monkeypatch):
    _install(monkeypatch)
    g = load_golden("attention_unet_small")
    c = g["cfg"]
    net = b200seg.AttentionUNet(1, c["base"], scale=c["scale"], kernel_size=c["kernel"], num_classes=c["classes"], block=c["block"], norm="in")
    sd = ounet.make_state_dict(g["shapes"], seed=c["state_seed"])
    net.load_state_dict(sd)
    img, lab = make_volume(*c["shape"], c["classes"], seed=c["data_seed"])
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    logits = net(img)
    w = torch.tensor(c["ce_weight"])
    loss = b200seg.DiceCELoss(weight=w)(logits, lab)
    loss.backward()
    assert rel_err(logits, g["logits"].float()) < 2e-3 and abs(loss.item() - g["loss"]) < 1e-4
    s64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    l64 = oatt.attention_unet_forward(s64, img.double(), c["scale"], c["kernel"], c["block"])
    olosses.total_loss(l64, lab, w.double()).backward()
    assert rel_err(logits, l64) < 1e-4
    ours = {k: p.grad for k, p in net.named_parameters()}
    assert sorted(k for k, v in ours.items() if v is None) == sorted(g["unused"])        # conv_ch: unused, as in the reference
    g64 = {k: v.grad for k, v in s64.items() if v.grad is not None}
    err = global_l2({k: ours[k] for k in g64}, g64)
    print("attention-unet emulated-orchestration grad L2 err vs fp64 oracle: %.2e" % err)
    assert err < 5e-2
