"""Pin oracle/attention_unet.py and b200seg.AttentionUNet against the UNMODIFIED reference class
(model/dim3/attention_unet.py) and write tests/golden/attention_unet_small.pt (state_dict keys / shapes, logits, loss,
per-parameter gradient digests of one CPU fp32 step with seeded weights) and tests/golden/attn_gate_{a,b}.pt (one
AttentionBlock forward / backward with full tensors).  Runs only where /root/reference exists.
Usage:  python oracle/make_golden_attention_unet.py"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import digest, import_reference     # noqa: E402
from oracle import attention_unet as oatt                    # noqa: E402
from oracle import losses as olosses                         # noqa: E402
from oracle import unet3d as ounet                           # noqa: E402
from oracle.synth import make_volume                         # noqa: E402

CFG = dict(base=16, classes=4, scale=[[1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
           kernel=[[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]], block="BasicBlock", shape=(1, 16, 32, 32),
           ce_weight=[0.5, 1.0, 1.0, 2.0], state_seed=91, data_seed=92)
GATES = {"a": dict(g_ch=32, l_ch=16, int_ch=8, shape=(2, 4, 6, 9), seed=93),
         "b": dict(g_ch=64, l_ch=64, int_ch=32, shape=(1, 3, 7, 8), seed=94)}


def main():
    torch.set_num_threads(8)
    import_reference()
    from model.dim3.attention_unet import AttentionUNet
    from model.dim3.attention_unet_utils import AttentionBlock
    c = CFG
    net = AttentionUNet(1, c["base"], scale=c["scale"], kernel_size=c["kernel"], num_classes=c["classes"], block=c["block"], norm="in")
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes == oatt.attention_unet_param_shapes(1, c["base"], c["classes"], c["kernel"], c["block"])
    assert list(shapes) == list(oatt.attention_unet_param_shapes(1, c["base"], c["classes"], c["kernel"], c["block"]))
    sd = ounet.make_state_dict(shapes, seed=c["state_seed"])
    net.load_state_dict(sd)
    img, lab = make_volume(*c["shape"], c["classes"], seed=c["data_seed"])
    w = torch.tensor(c["ce_weight"])
    logits = net(img)
    loss = nn.CrossEntropyLoss(weight=w)(logits, lab.squeeze(1)) + olosses.dice_loss(logits, lab)
    loss.backward()
    # the oracle restatement == the reference (forward bit-equal up to reduction order, gradients 1e-5)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = oatt.attention_unet_forward(sdo, img, c["scale"], c["kernel"], c["block"])
    lo_loss = olosses.total_loss(lo, lab, w)
    lo_loss.backward()
    dl = (lo - logits).abs().max().item()
    grads = {k: p.grad for k, p in net.named_parameters()}
    unused = [k for k, g in grads.items() if g is None]
    dg = max((sdo[k].grad - g).abs().max().item() / (g.abs().max().item() + 1e-12) for k, g in grads.items() if g is not None)
    print("attention_unet_small: %d tensors, %d params, loss %.6f | oracle-vs-reference logits %.2e loss %.2e grads %.2e | no grad: %s"
          % (len(shapes), sum(v.numel() for v in sd.values()), loss.item(), dl, abs(lo_loss.item() - loss.item()), dg, unused))
    assert dl < 1e-5 and dg < 1e-4 and all("conv_ch" in k for k in unused)
    torch.save({"cfg": c, "shapes": shapes, "logits": logits.detach().half(), "argmax": logits.argmax(1).to(torch.uint8), "loss": loss.item(),
                "unused": unused, "grad_digest": {k: digest(g) for k, g in grads.items() if g is not None}},
               os.path.join(ROOT, "tests", "golden", "attention_unet_small.pt"))
    for name, gc in GATES.items():
        g = torch.Generator().manual_seed(gc["seed"])
        blk = AttentionBlock(gc["g_ch"], gc["l_ch"], gc["int_ch"])
        B, D, H, W = gc["shape"]
        gin = torch.randn(B, gc["g_ch"], D, H, W, generator=g, requires_grad=True)
        xin = torch.randn(B, gc["l_ch"], D, H, W, generator=g, requires_grad=True)
        dout = torch.randn(B, gc["l_ch"], D, H, W, generator=g)
        out = blk(gin, xin)
        out.backward(dout)
        bsd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
        so = {k: v.clone().requires_grad_(True) for k, v in bsd.items()}
        go, xo = gin.detach().clone().requires_grad_(True), xin.detach().clone().requires_grad_(True)
        oo = oatt.attention_block(so, "", go, xo)
        oo.backward(dout)
        assert (oo - out).abs().max().item() < 1e-6 and (go.grad - gin.grad).abs().max().item() < 1e-5
        torch.save({"cfg": gc, "sd": bsd, "g": gin.detach(), "x": xin.detach(), "dout": dout, "out": out.detach(),
                    "dg": gin.grad, "dx": xin.grad, "dw": {k: p.grad for k, p in blk.named_parameters()}},
                   os.path.join(ROOT, "tests", "golden", "attn_gate_%s.pt" % name))
        print("attn_gate_%s: out %s" % (name, tuple(out.shape)))


if __name__ == "__main__":
    main()
