"""CPU: b200seg.UNetPlusPlus — state_dict contract, get_model dispatch and the nested-skip wiring with the C-ABI ops emulated
(tests/emu_ops.py) against the fixture the UNMODIFIED reference class produced (oracle/make_golden_unetpp.py)."""
import types

import torch

import b200seg
import emu_medformer
from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import load_golden, rel_err


def test_state_dict_contract_and_factory():
    g = load_golden("unetpp_small")
    c = g["cfg"]
    args = types.SimpleNamespace(dimension="3d", model="unet++", in_chan=1, base_chan=c["base"], classes=c["classes"],
                                 down_scale=c["scale"], norm="in", kernel_size=c["kernel"], block=c["block"])
    net = b200seg.get_model(args)
    assert isinstance(net, b200seg.UNetPlusPlus)
    assert list(net.state_dict().keys()) == list(g["shapes"].keys())
    assert all(tuple(v.shape) == tuple(g["shapes"][k]) for k, v in net.state_dict().items())


def test_orchestration_matches_reference_fixture(monkeypatch):
    emu_medformer.install(monkeypatch)
    from b200seg import ops, unetpp
    monkeypatch.setattr(unetpp, "MaxPoolFn", ops.MaxPoolFn)       # the emulated versions emu_ops installed on `ops`
    monkeypatch.setattr(unetpp, "UpCatFn", ops.UpCatFn)
    g = load_golden("unetpp_small")
    c = g["cfg"]
    net = b200seg.UNetPlusPlus(1, c["base"], scale=c["scale"], kernel_size=c["kernel"], num_classes=c["classes"], block=c["block"], norm="in")
    net.load_state_dict(ounet.make_state_dict(g["shapes"], seed=c["state_seed"]))
    img, lab = make_volume(*c["shape"], c["classes"], seed=c["data_seed"])
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    logits = net(img)
    loss = b200seg.DiceCELoss(weight=torch.tensor(c["ce_weight"]))(logits, lab)
    loss.backward()
    assert rel_err(logits, g["logits"].float()) < 2e-3                    # fixture stored in fp16
    assert (logits.argmax(1).to(torch.uint8) == g["argmax"]).float().mean().item() > 0.9995
    assert abs(loss.item() - g["loss"]) < 1e-4
    num = den = 0.0
    for k, p in net.named_parameters():
        d = g["grad_digest"][k]
        t = p.grad.detach().double().flatten()
        idx = torch.linspace(0, t.numel() - 1, min(t.numel(), 64)).long()
        num += ((t[idx] - d["sample"].double()) ** 2).sum().item()
        den += (d["sample"].double() ** 2).sum().item()
        assert abs((t * t).sum().item() - d["sq"]) <= 0.05 * d["sq"] + 1e-12, k
    assert (num / den) ** 0.5 < 2e-2
