"""GPU: b200seg.UNetPlusPlus against the fixture the UNMODIFIED reference class produced
(oracle/make_golden_unetpp.py): state_dict contract, logits, label map, loss, gradients."""
import pytest
import torch

from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("amp", [False, True])
def test_unetpp_matches_reference(amp):
    import b200seg
    g = load_golden("unetpp_small")
    c = g["cfg"]
    net = b200seg.UNetPlusPlus(1, c["base"], scale=c["scale"], kernel_size=c["kernel"], num_classes=c["classes"], block=c["block"], norm="in")
    assert list(net.state_dict().keys()) == list(g["shapes"].keys())
    assert all(tuple(v.shape) == tuple(g["shapes"][k]) for k, v in net.state_dict().items())
    net.load_state_dict(ounet.make_state_dict(g["shapes"], seed=c["state_seed"]))
    net = net.cuda()
    img, lab = make_volume(*c["shape"], c["classes"], seed=c["data_seed"])
    S = 1024.0 if amp else 1.0
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        logits = net(img.cuda())
        loss = b200seg.DiceCELoss(weight=torch.tensor(c["ce_weight"]))(logits, lab.cuda())
    (loss * S).backward()
    lg = logits.detach().float().cpu()
    e = rel_err(lg, g["logits"].float())
    agree = (lg.argmax(1).to(torch.uint8) == g["argmax"]).float().mean().item()
    # gradients: against the reference's per-parameter digests (sum of squares and 64 strided samples per tensor)
    num = den = 0.0
    for k, p in net.named_parameters():
        d = g["grad_digest"][k]
        t = (p.grad / S).detach().double().flatten().cpu()
        idx = torch.linspace(0, t.numel() - 1, min(t.numel(), 64)).long()
        num += ((t[idx] - d["sample"].double()) ** 2).sum().item()
        den += (d["sample"].double() ** 2).sum().item()
        assert abs((t * t).sum().item() - d["sq"]) <= (0.5 if amp else 0.05) * d["sq"] + 1e-12, k
    l2 = (num / den) ** 0.5
    print("unet++ amp=%d: logits rel err %.2e, label agreement %.5f, loss %.6f (ref %.6f), grads global-L2 vs reference fp32 %.2e"
          % (amp, e, agree, loss.item(), g["loss"], l2))
    if amp:
        assert e < 5e-2 and agree > 0.97 and abs(loss.item() - g["loss"]) < 3e-2 and l2 < 0.3
    else:
        assert e < 2e-3 and agree > 0.9995 and abs(loss.item() - g["loss"]) < 1e-4 and l2 < 2e-2
