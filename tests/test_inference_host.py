"""CPU: b200seg.inference (sliding-window schedule, pad / crop path, whole-image path, get_inference dispatch, Dice metric)
against fixtures the UNMODIFIED reference functions produced (oracle/make_golden_inference.py), with the three C-ABI
entry points emulated on the host pointers they are handed (numpy views) — the product code runs unchanged."""
import ctypes
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from b200seg import inference as inf
from oracle.make_golden_inference import tiny_net
from util import load_golden


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _emulate(monkeypatch):
    def softmax_accumulate(logits, prob, counter, region):
        d0, h0, w0, wd, wh, ww = region
        p = F.softmax(logits.float(), dim=1)
        prob[:, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww] += p
        counter[:, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww] += 1

    def call(name, *a):
        if name == "b200seg_normalize_argmax":
            prob_p, cnt_p, lab_p, B, C, V, _ = a
            prob, cnt = _view(prob_p, (B, C, V), np.float32), _view(cnt_p, (B, 1, V), np.float32)
            prob /= cnt
            if lab_p:
                _view(lab_p, (B, V), np.uint8)[...] = prob.argmax(1)
        elif name == "b200seg_dice_metric":
            p_p, pb, t_p, tb, N, C, out_p, _ = a
            p = _view(p_p, (N,), np.uint8 if pb == 1 else np.int64).astype(np.int64)
            t = _view(t_p, (N,), np.uint8 if tb == 1 else np.int64).astype(np.int64)
            out = _view(out_p, (C, 2), np.int64)
            for c in range(C):
                out[c, 0] += int(((p == c) & (t == c)).sum())
                out[c, 1] += int((p == c).sum() + (t == c).sum())
        else:
            raise AssertionError("unexpected entry point " + name)
    monkeypatch.setattr(inf, "_softmax_accumulate", softmax_accumulate)
    monkeypatch.setattr(inf._lib, "call", call)
    monkeypatch.setattr(inf, "_need_cuda", lambda t: None)
    monkeypatch.setattr(inf, "_stream", lambda: 0)


@pytest.mark.parametrize("case", ["exact", "snapped", "padded"])
def test_sliding_window_matches_reference(monkeypatch, case):
    _emulate(monkeypatch)
    g = load_golden("inference_ref")
    c = g["cases"][case]
    net = tiny_net(g["net_seed"])
    args = types.SimpleNamespace(window_size=c["window"], classes=g["classes"], dimension="3d", sliding_window=True)
    assert inf.get_inference(args) is inf.inference_sliding_window
    prob, label = inf.inference_sliding_window(net, c["img"], args, return_label=True)
    assert prob.shape == c["sliding"].shape and (prob - c["sliding"]).abs().max().item() < 1e-6
    assert torch.equal(label.long(), c["sliding"].argmax(1))
    whole = inf.inference_whole_image(net, c["img"], args)
    assert (whole - c["whole"]).abs().max().item() < 1e-6
    args.sliding_window = False
    assert inf.get_inference(args) is inf.inference_whole_image


def test_dice_metric_matches_reference(monkeypatch):
    _emulate(monkeypatch)
    d = load_golden("inference_ref")["dice"]
    dice, inter, summ = inf.calculate_dice(d["pred"], d["target"], 5)
    assert torch.allclose(dice, d["dice"], atol=1e-6) and torch.equal(inter, d["intersection"]) and torch.allclose(summ, d["summ"])
    ds, is_, ss = inf.calculate_dice_split(d["pred"], d["target"], 5, block_size=d["block_size"])
    assert torch.allclose(ds, d["split"][0], atol=1e-6) and torch.equal(is_, d["split"][1])
    assert torch.allclose(ss, d["split"][2], atol=1e-3)          # the reference adds 1e-5 once per block


def test_no_cpu_path():
    import b200seg
    with pytest.raises(b200seg.B200SegError):
        inf.calculate_dice(torch.zeros(8, 1, dtype=torch.uint8), torch.zeros(8, 1, dtype=torch.uint8), 2)
