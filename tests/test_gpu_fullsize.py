"""GPU: parity at the BENCHMARKED configuration (BASELINE.json configs[1]: ResUNet base 32, literal ACDC kernel/scale
lists, 1x128^3, 4 classes) — the exact kernel / tiling mix bench.py times (resident vs streamed weights, split-K
wgrad, 2-tile N) — against the reference-pinned oracle evaluated on the SAME GPU in fp32 with TF32 off and in fp64.
Also: the mask-flip-free gradient case (1e-3 asserted with NO noise-floor allowance) and the `.data`-mutation
(reference EMA, training/utils.py:99-102) staleness check."""
import pytest
import torch

from oracle import losses as olosses
from oracle import unet3d as ounet
from oracle.synth import make_volume
from util import global_l2, rel_err

pytestmark = pytest.mark.gpu

SCALE = [[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]]
KERNEL = [[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]]
BASE, CLASSES, SHAPE = 32, 4, (1, 128, 128, 128)
CE_W = [0.5, 1.0, 1.0, 1.0]


def _oracle_gpu(sd, img, lab, w, dtype, autocast=False, scale=1.0):
    dev = img.device
    s = {k: v.to(dev, dtype).requires_grad_(True) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
        lo = ounet.unet_forward(s, img.to(dtype), SCALE, KERNEL, "BasicBlock")
        loss = olosses.total_loss(lo, lab, w.to(dev, dtype if dtype == torch.float64 else torch.float32))
    (loss * scale).backward()
    grads = {k: (v.grad / scale).double().cpu() for k, v in s.items()}
    return lo.detach().double().cpu(), loss.item(), grads


@pytest.fixture(scope="module")
def fullsize():
    import b200seg
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = ounet.unet_param_shapes(1, BASE, CLASSES, KERNEL, "BasicBlock")
    sd = ounet.make_state_dict(shapes, seed=7)
    img, lab = make_volume(*SHAPE, CLASSES, seed=2023)
    img, lab = img.cuda(), lab.cuda()
    w = torch.tensor(CE_W)
    net = b200seg.UNet(1, BASE, scale=SCALE, kernel_size=KERNEL, num_classes=CLASSES, block="BasicBlock", norm="in")
    net.load_state_dict(sd)
    net = net.cuda()
    ref = {}
    ref["l32"], ref["loss32"], ref["g32"] = _oracle_gpu(sd, img, lab, w, torch.float32)
    torch.cuda.empty_cache()
    ref["l64"], ref["loss64"], ref["g64"] = _oracle_gpu(sd, img, lab, w, torch.float64)
    torch.cuda.empty_cache()
    yield net, sd, img, lab, w, ref
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _margin_mask(l64, rel=1e-4):
    top2 = l64.topk(2, dim=1).values
    return (top2[:, 0] - top2[:, 1]) > rel * l64.abs().max()


def test_fullsize_fp32_step_matches_oracle(fullsize):
    import b200seg
    net, sd, img, lab, w, ref = fullsize
    net.zero_grad(set_to_none=True)
    logits = net(img)
    loss = b200seg.DiceCELoss(weight=w)(logits, lab)
    loss.backward()
    lg = logits.detach().double().cpu()
    e32, e64 = rel_err(lg, ref["l32"]), rel_err(lg, ref["l64"])
    ours = {k: p.grad.double().cpu() for k, p in net.named_parameters()}
    l2_32, l2_64, floor = global_l2(ours, ref["g32"]), global_l2(ours, ref["g64"]), global_l2(ref["g32"], ref["g64"])
    mask = _margin_mask(ref["l64"])
    am, am_ref = lg.argmax(1), ref["l64"].argmax(1)
    print("full-size fp32: logits rel err vs fp32 oracle %.2e, vs fp64 %.2e; loss %.6f (oracle %.6f); grads global-L2 vs "
          "fp32 oracle %.2e, vs fp64 %.2e (the fp32 oracle's own distance to fp64: %.2e); decided voxels %.5f, label "
          "agreement on all voxels %.6f" % (e32, e64, loss.item(), ref["loss64"], l2_32, l2_64, floor,
                                            mask.float().mean().item(), (am == am_ref).float().mean().item()))
    assert e32 < 1e-3 and e64 < 1e-3
    assert abs(loss.item() - ref["loss64"]) < 1e-4
    # bit-exact label maps wherever the exact (fp64) network is decided by more than 1e-4 of the logit range
    assert mask.float().mean().item() > 0.999
    assert torch.equal(am[mask], am_ref[mask])
    assert l2_64 < max(1e-3, 2 * floor)


def test_fullsize_amp_step_matches_oracle(fullsize):
    """The configuration and precision the headline voxels/s is measured in (autocast fp16 + loss scaling)."""
    import b200seg
    net, sd, img, lab, w, ref = fullsize
    net.zero_grad(set_to_none=True)
    S = 1024.0
    with torch.autocast("cuda", dtype=torch.float16):
        logits = net(img)
        loss = b200seg.DiceCELoss(weight=w)(logits, lab)
    (loss * S).backward()
    lg = logits.detach().double().cpu()
    ours = {k: (p.grad / S).double().cpu() for k, p in net.named_parameters()}
    l_amp, loss_amp, g_amp = _oracle_gpu(sd, img, lab, w, torch.float32, autocast=True, scale=S)
    e, e_stock = rel_err(lg, ref["l64"]), rel_err(l_amp, ref["l64"])
    l2, l2_stock = global_l2(ours, ref["g64"]), global_l2(g_amp, ref["g64"])
    mask = _margin_mask(ref["l64"], rel=2e-2)        # decided by more than the fp16 logit error bar
    am, am_ref = lg.argmax(1), ref["l64"].argmax(1)
    print("full-size AMP: logits rel err vs fp64 oracle %.2e (stock torch autocast: %.2e); loss %.5f (oracle %.5f); grads "
          "global-L2 %.2e (stock autocast %.2e); voxels decided by >2e-2 of range %.4f, agreement there %.6f, overall %.5f"
          % (e, e_stock, loss.item(), ref["loss64"], l2, l2_stock, mask.float().mean().item(),
             (am[mask] == am_ref[mask]).float().mean().item(), (am == am_ref).float().mean().item()))
    assert e < max(3e-2, 2 * e_stock)
    assert abs(loss.item() - ref["loss64"]) < 2e-2
    assert torch.equal(am[mask], am_ref[mask])
    assert l2 < max(0.05, 2 * l2_stock)


@pytest.mark.parametrize("cin,cout,seed", [(32, 32, 101), (16, 32, 130)])
def test_basic_block_mask_flip_free_gradients(cin, cout, seed):
    """A case PROVEN free of ReLU-mask flips: the fp64 oracle's smallest |normalised pre-activation| (1.4e-4 ... 1.9e-4,
    seeds found by search, re-checked here) is two orders above the fp32 forward error, so the discontinuity of
    relu'(0) cannot move any mask and <=1e-3 (max-norm, per tensor) is asserted with NO noise-floor allowance."""
    from b200seg import ops
    from b200seg.unet3d import BasicBlock
    k = [3, 3, 3]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, cin, 4, 16, 16, generator=g, dtype=torch.float64)
    shapes = {"b.conv1.conv.weight": (cout, cin, 3, 3, 3), "b.conv2.conv.weight": (cout, cout, 3, 3, 3)}
    if cin != cout:
        shapes["b.shortcut.conv.weight"] = (cout, cin, 3, 3, 3)
    sd = {n: v.double().requires_grad_(True) for n, v in ounet.make_state_dict(shapes, seed=seed).items()}
    xo = x.clone().requires_grad_(True)
    with ounet.relu_margin_probe() as margins:
        o = ounet.basic_block(sd, "b.", xo, k)
    assert min(margins) > 1e-4, margins
    gdy = torch.Generator().manual_seed(seed + 1)
    dy = torch.randn(o.shape, generator=gdy, dtype=torch.float64)
    o.backward(dy)
    blk = BasicBlock(cin, cout, k)
    blk.load_state_dict({n[2:]: v.detach().float() for n, v in sd.items()})
    blk = blk.cuda()
    xg = x.float().permute(0, 2, 3, 4, 1).contiguous().cuda().requires_grad_(True)
    st = ops.instnorm_stats(xg.detach(), 0, cin)
    out, _ = blk((xg, st))
    out.backward(dy.float().permute(0, 2, 3, 4, 1).contiguous().cuda())
    assert rel_err(out.permute(0, 4, 1, 2, 3), o) < 1e-5
    errs = {n: rel_err(p.grad, sd["b." + n].grad) for n, p in blk.named_parameters()}
    errs["input"] = rel_err(xg.grad.permute(0, 4, 1, 2, 3), xo.grad)
    print("flip-free block %d->%d: min |xhat| %.2e; max-norm rel grad errors %s" % (cin, cout, min(margins),
                                                                                   {n: "%.1e" % e for n, e in errs.items()}))
    assert max(errs.values()) < 1e-3, errs


def test_data_mutation_is_seen_by_the_next_forward():
    """The reference's EMA update writes through `.data` (training/utils.py:99-102), which does not bump
    Tensor._version; the packed weight images must nevertheless follow."""
    import b200seg
    kernel = [[3, 3, 3]] * 5
    scale = [[2, 2, 2]] * 4
    shapes = ounet.unet_param_shapes(1, 16, 3, kernel, "BasicBlock")
    sd = ounet.make_state_dict(shapes, seed=3)
    img, _ = make_volume(1, 16, 32, 32, 3, seed=4)
    img = img.cuda()

    def build(state):
        n = b200seg.UNet(1, 16, scale=scale, kernel_size=kernel, num_classes=3, block="BasicBlock", norm="in")
        n.load_state_dict(state)
        return n.cuda().eval()
    for amp in (False, True):
        net = build(sd)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            a = net(img).float()
            v0 = [p._version for p in net.parameters()]
            for p in net.parameters():                      # exactly the reference's EMA arithmetic
                p.data.mul_(0.5).add_(0.25 * torch.ones_like(p.data))
            assert [p._version for p in net.parameters()] == v0      # the premise: versions did not move
            b = net(img).float()
            fresh = build({k: v.detach().cpu() for k, v in net.state_dict().items()})
            c = fresh(img).float()
        assert not torch.equal(a, b)
        assert torch.equal(b, c), "forward after a .data update used stale packed weights"
