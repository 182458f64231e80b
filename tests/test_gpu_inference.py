"""GPU: the evaluation consumers of net(x) against a literal restatement of the reference functions
(inference/inference3d.py:28-92 sliding-window schedule with half-window stride and border snapping;
metric/utils.py:62-82 one-hot Dice) run with the same network."""
import types

import pytest
import torch
import torch.nn.functional as F

from oracle import unet3d as ounet
from oracle.synth import make_volume

pytestmark = pytest.mark.gpu


def _ref_sliding_window(net, img, args):
    """inference/inference3d.py:28-92, restated."""
    B, C, D, H, W = img.shape
    wd, wh, ww = args.window_size
    flag = False
    if D < wd or H < wh or W < ww:
        flag = True
        img = F.pad(img, (0, max(0, ww - W), 0, max(0, wh - H), 0, max(0, wd - D)))
        oD, oH, oW = D, H, W
        B, C, D, H, W = img.shape
    hd, hh, hw = wd // 2, wh // 2, ww // 2
    out = torch.zeros((B, args.classes, D, H, W), device=img.device)
    cnt = torch.zeros((B, 1, D, H, W), device=img.device)

    def split(half, size, i):
        s = half * i
        e = s + 2 * half
        return (size - 2 * half, size) if e > size else (s, e)
    with torch.no_grad():
        for i in range(D // hd):
            for j in range(H // hh):
                for k in range(W // hw):
                    d0, d1 = split(hd, D, i); h0, h1 = split(hh, H, j); w0, w1 = split(hw, W, k)
                    pred = F.softmax(net(img[:, :, d0:d1, h0:h1, w0:w1]).float(), dim=1)
                    out[:, :, d0:d1, h0:h1, w0:w1] += pred
                    cnt[:, :, d0:d1, h0:h1, w0:w1] += 1
    out /= cnt
    return out[:, :, :oD, :oH, :oW] if flag else out


def _ref_dice(pred, target, C):
    """metric/utils.py:62-82, restated."""
    target, pred = target.long(), pred.long()
    N = pred.shape[0]
    tm = target.new_zeros(N, C).scatter_(1, target, 1)
    pm = pred.new_zeros(N, C).scatter_(1, pred, 1)
    inter = (pm * tm).sum(0).float()
    summ = (pm + tm).sum(0).float() + 1e-5
    return 2 * inter / summ, inter, summ


@pytest.mark.parametrize("shape,window", [((1, 24, 40, 48), (16, 32, 32)), ((2, 8, 20, 24), (16, 32, 32))])
@pytest.mark.parametrize("amp", [False, True])
def test_sliding_window_matches_reference(shape, window, amp):
    import b200seg
    kernel, scale, classes = [[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]], [[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]], 5
    net = b200seg.UNet(1, 16, scale=scale, kernel_size=kernel, num_classes=classes, block="BasicBlock", norm="in")
    net.load_state_dict(ounet.make_state_dict(ounet.unet_param_shapes(1, 16, classes, kernel, "BasicBlock"), seed=9))
    net = net.cuda().eval()
    img, lab = make_volume(*shape, classes, seed=21)
    img = img.cuda()
    args = types.SimpleNamespace(window_size=window, classes=classes, dimension="3d", sliding_window=True)
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        prob, label = b200seg.inference_sliding_window(net, img, args, return_label=True)
        ref = _ref_sliding_window(net, img, args)
    assert prob.shape == ref.shape and prob.dtype == torch.float32
    assert (prob - ref).abs().max().item() < 2e-6
    assert torch.equal(label.long(), prob.argmax(1))
    assert b200seg.get_inference(args) is b200seg.inference_sliding_window
    whole = b200seg.inference_whole_image(net, img[:, :, :16, :32, :32].contiguous())
    with torch.no_grad():
        assert (whole - F.softmax(net(img[:, :, :16, :32, :32].contiguous()).float(), 1)).abs().max().item() < 2e-6


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
def test_dice_metric_matches_reference(dtype):
    import b200seg
    torch.manual_seed(5)
    C, N = 14, 96 * 96 * 50 + 13
    pred = torch.randint(0, C, (N, 1), device="cuda").to(dtype)
    tgt = torch.randint(0, C, (N, 1), device="cuda")
    tgt[: N // 3] = pred[: N // 3].long()
    d, i, s = b200seg.calculate_dice(pred, tgt, C)
    rd, ri, rs = _ref_dice(pred, tgt, C)
    assert torch.equal(i, ri) and torch.allclose(s, rs) and torch.allclose(d, rd, rtol=1e-6)
    d2, i2, s2 = b200seg.calculate_dice_split(pred, tgt, C)
    assert torch.allclose(d2, 2 * ri / ((rs - 1e-5) + 1e-5), rtol=1e-6)
