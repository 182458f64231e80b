"""Pin b200seg.UNetPlusPlus against the UNMODIFIED reference class (model/dim3/unetpp.py) and write
tests/golden/unetpp_small.pt: state_dict keys / shapes, logits, loss and per-parameter gradient digests of one
CPU fp32 forward + CE + Dice + backward with seeded weights.  Runs only where /root/reference exists.
Usage:  python oracle/make_golden_unetpp.py"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import digest, import_reference     # noqa: E402
from oracle import losses as olosses                         # noqa: E402
from oracle import unet3d as ounet                           # noqa: E402
from oracle.synth import make_volume                         # noqa: E402

CFG = dict(base=16, classes=4, scale=[[1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
           kernel=[[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]], block="BasicBlock", shape=(1, 16, 32, 32),
           ce_weight=[0.5, 1.0, 1.0, 2.0], state_seed=81, data_seed=82)


def main():
    torch.set_num_threads(8)
    import_reference()
    from model.dim3.unetpp import UNetPlusPlus
    c = CFG
    net = UNetPlusPlus(1, c["base"], scale=c["scale"], kernel_size=c["kernel"], num_classes=c["classes"], block=c["block"], norm="in")
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = ounet.make_state_dict(shapes, seed=c["state_seed"])
    net.load_state_dict(sd)
    img, lab = make_volume(*c["shape"], c["classes"], seed=c["data_seed"])
    w = torch.tensor(c["ce_weight"])
    logits = net(img)
    loss = nn.CrossEntropyLoss(weight=w)(logits, lab.squeeze(1)) + olosses.dice_loss(logits, lab)
    loss.backward()
    torch.save({"cfg": c, "shapes": shapes, "logits": logits.detach().half(), "argmax": logits.argmax(1).to(torch.uint8), "loss": loss.item(),
                "grad_digest": {k: digest(p.grad) for k, p in net.named_parameters()}},
               os.path.join(ROOT, "tests", "golden", "unetpp_small.pt"))
    print("unetpp_small: %d tensors, %d params, loss %.6f" % (len(shapes), sum(v.numel() for v in sd.values()), loss.item()))


if __name__ == "__main__":
    main()
