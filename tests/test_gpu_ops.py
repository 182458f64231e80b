"""GPU: each C-ABI op against a plain PyTorch fp32 CPU evaluation of the same maths (oracle side)."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu

TOL32 = 1e-4   # fp32 storage, fp32 accumulate
TOL16 = 4e-3   # fp16 storage (one rounding of inputs/outputs), fp32 accumulate


def ndhwc(t):   # [B,C,D,H,W] -> [B,D,H,W,C] contiguous
    return t.permute(0, 2, 3, 4, 1).contiguous()


def ncdhw(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


def stats_of(x):  # x [B,C,...] -> [B,C,2] double
    xd = x.double().flatten(2)
    return torch.stack([xd.sum(-1), (xd * xd).sum(-1)], -1)


@pytest.fixture(scope="module")
def ops():
    import b200seg
    from b200seg import ops as o
    assert o._lib.load().b200seg_check_device() == 0, "not a B200"
    return o


@pytest.mark.parametrize("dtype,tol", [(torch.float32, TOL32), (torch.float16, TOL16)])
@pytest.mark.parametrize("C", [8, 32, 5])
def test_instnorm_stats(ops, dtype, tol, C):
    torch.manual_seed(0)
    x = (torch.randn(2, C, 6, 10, 12) * 2 + 0.5).to(dtype)
    st = ops.instnorm_stats(ndhwc(x).cuda(), 0, C)
    assert rel_err(st, stats_of(x.float())) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, TOL32), (torch.float16, TOL16)])
@pytest.mark.parametrize("Cin,Cout,k", [(1, 8, (3, 3, 3)), (8, 16, (3, 3, 3)), (16, 8, (1, 3, 3)), (24, 4, (1, 1, 1)), (5, 7, (3, 3, 3))])
def test_conv_fwd_preact_residual_stats(ops, dtype, tol, Cin, Cout, k):
    torch.manual_seed(1)
    B, D, H, W = 2, 6, 9, 11
    x = torch.randn(B, Cin, D, H, W).to(dtype)
    w = torch.randn(Cout, Cin, *k) * 0.2
    res = torch.randn(B, Cout, D, H, W).to(dtype)
    xc, rc = ndhwc(x).cuda(), ndhwc(res).cuda()
    st = ops.instnorm_stats(xc, 0, Cin)
    wp = ops.pack_weight(w.cuda(), dtype)
    y, yst = ops.conv3d_fwd(xc, 0, Cin, st, ops.ACT_RELU, wp, Cout, k, residual=rc, algo=ops._lib.ALGO_DIRECT)
    a = F.relu(F.instance_norm(x.float(), eps=1e-4)).to(dtype).float()
    ref = F.conv3d(a, w.to(dtype).float(), padding=[i // 2 for i in k]).to(dtype).float() + res.float()
    assert rel_err(ncdhw(y.float()), ref) < tol
    assert rel_err(yst, stats_of(ncdhw(y.float().cpu()))) < 1e-5      # sums describe what was STORED


@pytest.mark.parametrize("dtype,tol", [(torch.float32, TOL32), (torch.float16, TOL16)])
@pytest.mark.parametrize("Cin,Cout,k", [(1, 8, (3, 3, 3)), (8, 16, (3, 3, 3)), (16, 8, (1, 3, 3)), (24, 4, (1, 1, 1))])
def test_conv_wgrad_and_dgrad(ops, dtype, tol, Cin, Cout, k):
    torch.manual_seed(2)
    B, D, H, W = 2, 6, 8, 10
    x = torch.randn(B, Cin, D, H, W).to(dtype).float().requires_grad_(True)
    w = (torch.randn(Cout, Cin, *k) * 0.2).to(dtype).float().requires_grad_(True)
    dy = torch.randn(B, Cout, D, H, W).to(dtype).float()
    a = F.relu(F.instance_norm(x, eps=1e-4))
    y = F.conv3d(a, w, padding=[i // 2 for i in k])
    y.backward(dy)
    xc, dyc = ndhwc(x.detach().to(dtype)).cuda(), ndhwc(dy.to(dtype)).cuda()
    st = ops.instnorm_stats(xc, 0, Cin)
    dw, _ = ops.conv3d_wgrad(xc, 0, Cin, st, ops.ACT_RELU, dyc, 0, Cout, k, algo=ops._lib.ALGO_DIRECT)
    assert rel_err(dw, w.grad) < max(tol, 2e-3 if dtype == torch.float16 else tol)
    # data gradient incl. ReLU mask + both InstanceNorm-backward reductions
    wpb = ops.pack_weight(w.detach().cuda(), dtype, transpose_flip=True)
    g, bst = ops.conv3d_fwd(dyc, 0, Cout, None, ops.ACT_NONE, wpb, Cin, k, dgrad_of=(xc, 0, st, ops.ACT_RELU),
                            algo=ops._lib.ALGO_DIRECT)
    dx = ops.in_bwd_apply(g, xc, 0, Cin, st, bst)
    assert rel_err(ncdhw(dx.float()), x.grad) < (2e-2 if dtype == torch.float16 else 1e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("scale", [(2, 2, 2), (1, 2, 2)])
@pytest.mark.parametrize("C", [16, 3])
def test_maxpool(ops, dtype, scale, C):
    torch.manual_seed(3)
    x = torch.randn(2, C, 8, 12, 10).to(dtype).float().requires_grad_(True)
    y = F.max_pool3d(x, scale)
    dy = torch.randn_like(y).to(dtype).float()
    y.backward(dy)
    xc = ndhwc(x.detach().to(dtype)).cuda().requires_grad_(True)
    yo, st = ops.MaxPoolFn.apply(xc, scale, True)
    assert torch.equal(ncdhw(yo.detach().float().cpu()), y.detach())
    assert rel_err(st, stats_of(y.detach())) < 1e-5
    yo.backward(ndhwc(dy.to(dtype)).cuda())
    assert torch.equal(ncdhw(xc.grad.float().cpu()), x.grad)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float16, 2e-3)])
@pytest.mark.parametrize("shp", [((4, 5, 6), (8, 10, 12)), ((8, 3, 5), (8, 6, 10)), ((3, 3, 3), (7, 5, 6))])
def test_upsample_cat(ops, dtype, tol, shp):
    torch.manual_seed(4)
    (Di, Hi, Wi), (Do, Ho, Wo) = shp
    Cl, Cs = 16, 8
    low = torch.randn(2, Cl, Di, Hi, Wi).to(dtype).float().requires_grad_(True)
    skip = torch.randn(2, Cs, Do, Ho, Wo).to(dtype).float().requires_grad_(True)
    up = F.interpolate(low, size=(Do, Ho, Wo), mode="trilinear", align_corners=True)
    cat = torch.cat([skip, up], 1)
    dcat = torch.randn_like(cat).to(dtype).float()
    cat.backward(dcat)
    lc = ndhwc(low.detach().to(dtype)).cuda().requires_grad_(True)
    sc = ndhwc(skip.detach().to(dtype)).cuda().requires_grad_(True)
    co, cst = ops.UpCatFn.apply(lc, sc, None, True)
    assert rel_err(ncdhw(co.detach().float()), cat.detach()) < tol
    assert rel_err(cst, stats_of(ncdhw(co.detach().float().cpu()))) < 1e-5
    co.backward(ndhwc(dcat.to(dtype)).cuda())
    assert rel_err(ncdhw(lc.grad.float()), low.grad) < max(tol, 1e-5) * 2
    assert rel_err(ncdhw(sc.grad.float()), skip.grad) < 1e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.float16, 2e-3)])
@pytest.mark.parametrize("Cin,Cout,k,normed,bias", [(1, 32, (1, 3, 3), False, False), (1, 32, (3, 3, 3), False, False),
                                                    (1, 48, (3, 3, 3), False, False), (1, 48, (1, 1, 1), False, False), (32, 4, (1, 1, 1), False, True),
                                                    (64, 14, (1, 1, 1), True, True)])
def test_small_wgrad_special_cases(ops, dtype, tol, Cin, Cout, k, normed, bias):
    """AUTO routes the Cin=1 stem and the 1x1x1 few-class head to the HBM-bound special kernels (small_conv.cu)."""
    torch.manual_seed(5)
    B, D, H, W = 2, 5, 12, 9
    x = torch.randn(B, Cin, D, H, W).to(dtype).float()
    w = torch.zeros(Cout, Cin, *k, requires_grad=True)
    b = torch.zeros(Cout, requires_grad=True)
    dy = torch.randn(B, Cout, D, H, W).to(dtype).float()
    a = F.relu(F.instance_norm(x, eps=1e-4)).to(dtype).float() if normed else x
    F.conv3d(a, w, b, padding=[i // 2 for i in k]).backward(dy)
    xc, dyc = ndhwc(x.to(dtype)).cuda(), ndhwc(dy.to(dtype)).cuda()
    st = ops.instnorm_stats(xc, 0, Cin) if normed else None
    dw, db = ops.conv3d_wgrad(xc, 0, Cin, st, ops.ACT_RELU if normed else ops.ACT_NONE, dyc, 0, Cout, k, want_bias=bias)
    assert rel_err(dw, w.grad) < tol
    if bias:
        assert rel_err(db, b.grad) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.float16, 3e-3)])
@pytest.mark.parametrize("Cin,Cout,k,bias", [(1, 32, (1, 3, 3), False), (1, 32, (3, 3, 3), False), (32, 4, (1, 1, 1), True),
                                             (4, 32, (1, 1, 1), False), (16, 12, (1, 1, 1), True),
                                             (1, 48, (3, 3, 3), False), (1, 48, (1, 1, 1), False), (1, 64, (1, 3, 3), False)])
def test_small_fwd_special_cases(ops, dtype, tol, Cin, Cout, k, bias):
    """The stem (Cin=1 -> 32) and few-channel 1x1x1 convs (classifier head fwd / dgrad) run HBM-bound special kernels."""
    torch.manual_seed(6)
    B, D, H, W = 2, 5, 13, 9
    x = torch.randn(B, Cin, D, H, W).to(dtype)
    w = torch.randn(Cout, Cin, *k) * 0.3
    b = torch.randn(Cout) if bias else None
    xc = ndhwc(x).cuda()
    wp = ops.pack_weight(w.cuda(), dtype)
    y, yst = ops.conv3d_fwd(xc, 0, Cin, None, ops.ACT_NONE, wp, Cout, k, bias=None if b is None else b.cuda(),
                            want_stats=(Cin == 1), algo=ops._lib.ALGO_DIRECT)
    ref = F.conv3d(x.float(), w.to(dtype).float(), b, padding=[i // 2 for i in k])
    assert rel_err(ncdhw(y.float()), ref) < tol
    if yst is not None:
        assert rel_err(yst, stats_of(ncdhw(y.float().cpu()))) < 1e-5


@pytest.mark.parametrize("cout,cin,k,flip,off,total,layout_tc", [
    (32, 32, (1, 3, 3), False, 0, 32, True), (32, 32, (1, 3, 3), True, 0, 32, True),
    (64, 96, (3, 3, 3), False, 64, 128, True), (64, 96, (3, 3, 3), True, 64, 128, True),     # shortcut half of a fused GEMM
    (256, 384, (3, 3, 3), True, 0, 256, True), (128, 64, (1, 1, 1), False, 0, 128, True),
    (32, 16, (3, 3, 3), False, 0, 32, False), (4, 32, (1, 1, 1), False, 0, 4, False),         # direct layout / element chunks
])
def test_multi_tensor_pack_matches_single(cout, cin, k, flip, off, total, layout_tc):
    """b200seg_pack_weights_multi (TILE chunks: 8 co x tile_ci x taps through shared memory, 16-byte stores; ELEMENT chunks
    for shapes that do not qualify) writes exactly what b200seg_pack_weight writes."""
    from b200seg import _lib, ops
    torch.manual_seed(1)
    w = torch.randn(cout, cin, *k, device="cuda")
    taps = k[0] * k[1] * k[2]
    layout = ops.ALGO_TC if layout_tc else ops.ALGO_DIRECT
    ref = torch.zeros(taps * total * cin, dtype=torch.float16, device="cuda")
    ops.pack_weight(w, torch.float16, transpose_flip=flip, out=ref, co_off=off, co_total=total, layout=layout)
    out = torch.zeros_like(ref)
    lib = _lib.load()
    chunk, tile_ci = lib.b200seg_pack_chunk_elems(), lib.b200seg_pack_tile_ci(taps)
    jobs = [[w.data_ptr(), out.data_ptr(), cout, cin, taps, 1, 1 if flip else 0, off, total, 1 if layout_tc else 0]]
    if tile_ci and cout % 8 == 0 and cin % 8 == 0 and off % 8 == 0:
        chunks = [[0, -(1 + co0 * 65536 + ci0)] for co0 in range(0, cout, 8) for ci0 in range(0, cin, tile_ci)]
    else:
        chunks = [[0, e] for e in range(0, w.numel(), chunk)]
    jt, ct = torch.tensor(jobs, dtype=torch.int64).cuda(), torch.tensor(chunks, dtype=torch.int64).cuda()
    _lib.call("b200seg_pack_weights_multi", jt.data_ptr(), ct.data_ptr(), len(chunks), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
