"""Evaluation-side consumers of ``net(x)`` behind the reference's own function names:
  inference_whole_image / inference_sliding_window / get_inference   inference/inference3d.py:8-92, inference/utils.py:1-40
  calculate_dice / calculate_dice_split                               metric/utils.py:33-82
Same signatures and return values (probabilities fp32 [B,classes,D,H,W]; (dice, intersection, summ) tensors); the
window schedule (half-window stride, last window snapped to the border, utils.split_idx) is the reference's.  On the
device the per-window softmax + accumulate + count is ONE kernel, the final division (and an optional argmax label
map) another, and the Dice metric is a single pass over the two label maps — the [N,C] one-hot masks never exist."""
import torch
import torch.nn.functional as F

from . import _lib
from .ops import _dt, _need_cuda, _stream, on_device


def split_idx(half_win, size, i):
    """inference/utils.py:27-40."""
    start = half_win * i
    end = start + half_win * 2
    if end > size:
        start, end = size - half_win * 2, size
    return start, end


def _softmax_accumulate(logits, prob, counter, region):
    """prob[region] += softmax(logits, 1); counter[region] += 1.  logits: [B,C,d,h,w]-shaped view, any strides."""
    B, C = logits.shape[0], logits.shape[1]
    d0, h0, w0, wd, wh, ww = region
    if logits.dtype not in (torch.float16, torch.float32):
        logits = logits.float()
    cl = logits.permute(0, 2, 3, 4, 1)
    if cl.is_contiguous():                       # channels-last view handed out by the b200seg models
        sb, sv, sc = cl.stride(0), C, 1
    else:
        logits = logits.contiguous()
        sb, sv, sc = logits.stride(0), 1, logits.stride(1)
    D, H, W = prob.shape[2:]
    _lib.call("b200seg_softmax_accumulate", logits.data_ptr(), _dt(logits), sb, sv, sc, prob.data_ptr(), counter.data_ptr(),
              B, C, wd, wh, ww, D, H, W, d0, h0, w0, _stream())


def inference_whole_image(net, img, args=None):
    """inference/inference3d.py:8-25."""
    net.eval()
    with torch.no_grad():
        pred = net(img)
        if isinstance(pred, (tuple, list)):
            pred = pred[0]
    _need_cuda(pred)
    with on_device(pred):
        B, C = pred.shape[:2]
        prob = torch.zeros(B, C, *pred.shape[2:], dtype=torch.float32, device=pred.device)
        counter = torch.zeros(B, 1, *pred.shape[2:], dtype=torch.float32, device=pred.device)
        _softmax_accumulate(pred, prob, counter, (0, 0, 0, *pred.shape[2:]))
    return prob


def inference_sliding_window(net, img, args, return_label=False):
    """inference/inference3d.py:28-92.  return_label=True additionally returns the uint8 argmax map produced by the
    normalisation kernel (validation_ddp's `torch.max(pred, 1)`, training/validation.py:120-135, without another pass)."""
    net.eval()
    _need_cuda(img)
    B, C, D, H, W = img.shape
    win_d, win_h, win_w = args.window_size
    flag = False
    if D < win_d or H < win_h or W < win_w:
        flag = True
        img = F.pad(img, (0, max(0, win_w - W), 0, max(0, win_h - H), 0, max(0, win_d - D)))
        origin = (D, H, W)
        B, C, D, H, W = img.shape
    hd, hh, hw = win_d // 2, win_h // 2, win_w // 2
    with on_device(img):
        prob = torch.zeros(B, args.classes, D, H, W, dtype=torch.float32, device=img.device)
        counter = torch.zeros(B, 1, D, H, W, dtype=torch.float32, device=img.device)
        with torch.no_grad():
            for i in range(D // hd):
                for j in range(H // hh):
                    for k in range(W // hw):
                        d0, d1 = split_idx(hd, D, i)
                        h0, h1 = split_idx(hh, H, j)
                        w0, w1 = split_idx(hw, W, k)
                        pred = net(img[:, :, d0:d1, h0:h1, w0:w1])
                        if isinstance(pred, (tuple, list)):
                            pred = pred[0]
                        _softmax_accumulate(pred, prob, counter, (d0, h0, w0, d1 - d0, h1 - h0, w1 - w0))
        label = torch.empty(B, D, H, W, dtype=torch.uint8, device=img.device) if return_label else None
        _lib.call("b200seg_normalize_argmax", prob.data_ptr(), counter.data_ptr(), None if label is None else label.data_ptr(),
                  B, args.classes, D * H * W, _stream())
    if flag:
        prob = prob[:, :, :origin[0], :origin[1], :origin[2]]
        if label is not None:
            label = label[:, :origin[0], :origin[1], :origin[2]]
    return (prob, label) if return_label else prob


def get_inference(args):
    """inference/utils.py:1-23 for the 3D path."""
    if args.dimension == '3d':
        return inference_sliding_window if args.sliding_window else inference_whole_image
    if args.dimension == '2d':
        raise ValueError("2d inference is outside the B200 hot path (SURVEY.md §2); use the reference")
    raise ValueError('Error in image dimension')


def calculate_dice(pred, target, C):
    """metric/utils.py:62-82: pred / target label tensors [N,1] (or any shape with N elements); returns
    (dice [C], intersection [C], summ [C]) float32 like the reference (summ includes its +1e-5)."""
    _need_cuda(pred)
    p = pred.reshape(-1)
    t = target.reshape(-1)
    if p.dtype not in (torch.uint8, torch.int64):
        p = p.long()
    if t.dtype not in (torch.uint8, torch.int64):
        t = t.long()
    p, t = p.contiguous(), t.contiguous()
    with on_device(p):
        out = torch.zeros(C, 2, dtype=torch.int64, device=p.device)
        _lib.call("b200seg_dice_metric", p.data_ptr(), p.element_size(), t.data_ptr(), t.element_size(), p.numel(), C,
                  out.data_ptr(), _stream())
    intersection = out[:, 0].to(torch.float32)
    summ = out[:, 1].to(torch.float32) + 1e-5
    return 2 * intersection / summ, intersection, summ


def calculate_dice_split(pred, target, C, block_size=64 * 64 * 64):
    """metric/utils.py:33-55: the reference splits only to bound the memory of its one-hot masks; the device kernel has
    none, so this is one pass with the reference's final formula."""
    _, inter, summ = calculate_dice(pred, target, C)
    summ = summ - 1e-5
    return 2 * inter / (summ + 1e-5), inter, summ
