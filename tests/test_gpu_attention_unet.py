"""GPU: the Attention-UNet path — csrc/attn_gate.cu through the C ABI against fixtures the UNMODIFIED
reference classes produced (oracle/make_golden_attention_unet.py) and against fp64 PyTorch on further shapes."""
import pytest
import torch
import torch.nn.functional as F

from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


@pytest.mark.parametrize("name", ["a", "b"])
def test_attention_block_matches_reference_fixture(name):
    from b200seg.attention_unet import AttentionBlock
    f = load_golden("attn_gate_" + name)
    c = f["cfg"]
    blk = AttentionBlock(c["g_ch"], c["l_ch"], c["int_ch"])
    blk.load_state_dict(f["sd"])
    blk = blk.cuda()
    g = _cl(f["g"]).cuda().requires_grad_(True)
    x = _cl(f["x"]).cuda().requires_grad_(True)
    out, ost = blk(g, x)
    out.backward(_cl(f["dout"]).cuda())
    torch.cuda.synchronize()
    assert rel_err(out.permute(0, 4, 1, 2, 3), f["out"]) < 2e-4
    ref_st = f["out"].double().flatten(2)
    assert rel_err(ost[..., 0].cpu(), ref_st.sum(-1)) < 1e-4 and rel_err(ost[..., 1].cpu(), (ref_st * ref_st).sum(-1)) < 1e-4
    assert rel_err(g.grad.permute(0, 4, 1, 2, 3), f["dg"]) < 2e-3
    assert rel_err(x.grad.permute(0, 4, 1, 2, 3), f["dx"]) < 2e-3
    for k, p in blk.named_parameters():
        assert rel_err(p.grad, f["dw"][k]) < 2e-3, k


@pytest.mark.parametrize("dtype,B,V3,Ct,Cx", [(torch.float32, 2, (5, 7, 9), 24, 40), (torch.float32, 1, (8, 16, 16), 128, 256),
                                               (torch.float16, 2, (6, 10, 11), 16, 32), (torch.float16, 1, (4, 12, 16), 64, 8)])
def test_gate_kernels_vs_fp64(dtype, B, V3, Ct, Cx):
    """psi dot product + single-channel InstanceNorm + sigmoid gate, forward and backward, on channel counts that do
    and do not divide the block (24, 40), and voxel counts that are not multiples of the 32-voxel pass."""
    from b200seg.attention_unet import GATE_EPS, AttnGateFn
    g = torch.Generator().manual_seed(Ct * 1000 + Cx)
    t = torch.randn(B, *V3, Ct, generator=g).relu().to(dtype)
    x = torch.randn(B, *V3, Cx, generator=g).to(dtype)
    w = (torch.randn(1, Ct, 1, 1, 1, generator=g) * 0.3)
    dout = torch.randn(B, *V3, Cx, generator=g).to(dtype)
    tg, xg, wg = t.cuda().requires_grad_(True), x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    out, ost = AttnGateFn.apply(xg, tg, wg)
    out.backward(dout.cuda())
    torch.cuda.synchronize()
    t64, x64, w64 = t.double().requires_grad_(True), x.double().requires_grad_(True), w.double().requires_grad_(True)
    p = F.conv3d(t64.permute(0, 4, 1, 2, 3), w64)
    ref = x64 * torch.sigmoid(F.instance_norm(p, eps=GATE_EPS)).permute(0, 2, 3, 4, 1)
    ref.backward(dout.double())
    tol = 2e-4 if dtype == torch.float32 else 2e-2
    assert rel_err(out, ref) < tol
    assert rel_err(xg.grad, x64.grad) < tol * 5
    assert rel_err(tg.grad, t64.grad) < tol * 5
    assert rel_err(wg.grad, w64.grad) < tol * 5
    assert rel_err(ost[..., 0].cpu(), ref.detach().flatten(1, 3).sum(1)) < tol


@pytest.mark.parametrize("amp", [False, True])
def test_attention_unet_matches_reference(amp):
    import b200seg
    g = load_golden("attention_unet_small")
    c = g["cfg"]
    net = b200seg.AttentionUNet(1, c["base"], scale=c["scale"], kernel_size=c["kernel"], num_classes=c["classes"], block=c["block"], norm="in")
    assert list(net.state_dict().keys()) == list(g["shapes"].keys())
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
    num = den = 0.0
    for k, p in net.named_parameters():
        if k in g["unused"]:
            assert p.grad is None, k            # conv_ch is never called (attention_unet_utils.py:55-64)
            continue
        d = g["grad_digest"][k]
        t = (p.grad / S).detach().double().flatten().cpu()
        idx = torch.linspace(0, t.numel() - 1, min(t.numel(), 64)).long()
        num += ((t[idx] - d["sample"].double()) ** 2).sum().item()
        den += (d["sample"].double() ** 2).sum().item()
        assert abs((t * t).sum().item() - d["sq"]) <= (0.5 if amp else 0.05) * d["sq"] + 1e-12, k
    l2 = (num / den) ** 0.5
    print("attention-unet amp=%d: logits rel err %.2e, label agreement %.5f, loss %.6f (ref %.6f), grads global-L2 vs reference fp32 %.2e"
          % (amp, e, agree, loss.item(), g["loss"], l2))
    if amp:
        assert e < 5e-2 and agree > 0.97 and abs(loss.item() - g["loss"]) < 3e-2 and l2 < 0.3
    else:
        assert e < 2e-3 and agree > 0.9995 and abs(loss.item() - g["loss"]) < 1e-4 and l2 < 2e-2
