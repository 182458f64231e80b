"""GPU: the MedFormer kernels (B-MHA core, depthwise conv) through the C-ABI against the reference-pinned oracle
(oracle/medformer_ops.py) and the fixtures captured from the real reference modules (tests/golden/biattn_*,
dwconv_*).  Tolerances: fp32 storage 1e-4 / fp16 storage 4e-3 max-norm relative (same bars as test_gpu_ops)."""
import pytest
import torch

from oracle import medformer_ops as mops
from util import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 2e-4, torch.float16: 5e-3}


def ndhwc(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


def ncdhw(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


@pytest.fixture(scope="module")
def ops():
    import b200seg  # noqa
    from b200seg import ops as o
    assert o._lib.load().b200seg_check_device() == 0, "not a B200"
    return o


def _run_biattn(ops, fqv, mqv, heads, dfo, dmo, dtype):
    f = ndhwc(fqv).to(dtype).cuda().requires_grad_(True)
    m = ndhwc(mqv).to(dtype).cuda().requires_grad_(True)
    fo, mo = ops.BiAttnFn.apply(f, m, heads, 32)
    torch.autograd.backward([fo, mo], [ndhwc(dfo).to(dtype).cuda(), ndhwc(dmo).to(dtype).cuda()])
    torch.cuda.synchronize()
    return ncdhw(fo.float().cpu()), ncdhw(mo.float().cpu()), ncdhw(f.grad.float().cpu()), ncdhw(m.grad.float().cpu())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("name", ["biattn_a", "biattn_b", "biattn_c", "biattn_d"])
def test_biattn_golden(ops, name, dtype):
    g = load_golden(name)
    fo, mo, dfqv, dmqv = _run_biattn(ops, g["fqv"], g["mqv"], g["heads"], g["dfo"], g["dmo"], dtype)
    tol = TOL[dtype]
    assert rel_err(fo, g["fo"]) < tol and rel_err(mo, g["mo"]) < tol
    assert rel_err(dfqv, g["dfqv"]) < 2 * tol and rel_err(dmqv, g["dmqv"]) < 2 * tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("B,heads,fs,ms", [(2, 4, (12, 12, 12), (3, 3, 3)),      # BCV level shape, several blocks
                                            (1, 1, (3, 5, 131), (1, 1, 1)),        # single token, ragged block
                                            (1, 10, (6, 6, 6), (2, 4, 4)),         # 320-channel level, M = 32
                                            (2, 4, (10, 12, 12), (4, 4, 4)),       # 64 tokens (AMOS / KiTS map size)
                                            (1, 8, (3, 3, 5), (3, 4, 4))])         # 48 tokens in the 64-row build
def test_biattn_vs_oracle(ops, dtype, B, heads, fs, ms):
    torch.manual_seed(5)
    inner = 32 * heads
    fqv = (torch.randn(B, 2 * inner, *fs) * 1.5).to(dtype).float()
    mqv = (torch.randn(B, 2 * inner, *ms) * 1.5).to(dtype).float()
    dfo = torch.randn(B, inner, *fs).to(dtype).float()
    dmo = torch.randn(B, inner, *ms).to(dtype).float()
    f64, m64 = fqv.double().requires_grad_(True), mqv.double().requires_grad_(True)
    fo_r, mo_r = mops.bidirection_attention_core(*f64.chunk(2, 1), *m64.chunk(2, 1), heads)
    torch.autograd.backward([fo_r, mo_r], [dfo.double(), dmo.double()])
    fo, mo, dfqv, dmqv = _run_biattn(ops, fqv, mqv, heads, dfo, dmo, dtype)
    tol = TOL[dtype]
    assert rel_err(fo, fo_r) < tol and rel_err(mo, mo_r) < tol
    assert rel_err(dfqv, f64.grad) < 2 * tol and rel_err(dmqv, m64.grad) < 2 * tol


def test_biattn_rejects_unsupported(ops):
    f = torch.zeros(1, 2, 2, 2, 2 * 64, device="cuda")
    m = torch.zeros(1, 1, 1, 1, 2 * 64, device="cuda")
    with pytest.raises(ops._lib.B200SegError):
        ops.biattn_fwd(f, m, 1, dim_head=64)                       # dim_head != 32 -> loud, no fallback
    f = torch.zeros(1, 2, 2, 2, 64, device="cuda")
    m = torch.zeros(1, 2, 6, 6, 64, device="cuda")
    with pytest.raises(ops._lib.B200SegError):
        ops.biattn_fwd(f, m, 1)                                    # 72 map tokens (ACDC YAML) > 64


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("name", ["dwconv_a", "dwconv_b"])
def test_dwconv_golden(ops, name, dtype):
    g = load_golden(name)
    x = ndhwc(g["x"]).to(dtype).cuda().requires_grad_(True)
    w = g["w"].to(dtype).cuda().requires_grad_(True)
    y = ops.DepthwiseConvFn.apply(x, w)
    y.backward(ndhwc(g["gy"]).to(dtype).cuda())
    tol = TOL[dtype]
    assert rel_err(ncdhw(y.float().cpu()), g["y"]) < tol
    assert rel_err(ncdhw(x.grad.float().cpu()), g["dx"]) < tol
    assert rel_err(w.grad.float().cpu(), g["dw"]) < 2 * tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("B,C,sp,k", [(2, 64, (6, 24, 22), (3, 3, 3)), (1, 320, (3, 6, 6), (3, 3, 3)),
                                      (1, 40, (5, 7, 9), (1, 3, 3)), (1, 8, (1, 2, 3), (3, 3, 3))])
def test_dwconv_fused_prologue_and_stats(ops, dtype, B, C, sp, k):
    """a = relu(IN(x)) prologue from producer sums + IN sums of the output, vs the oracle composition."""
    torch.manual_seed(9)
    x = (torch.randn(B, C, *sp) * 2 + 0.3).to(dtype).float()
    w = torch.randn(C, 1, *k) * 0.3
    xs = x.double().flatten(2)
    st = torch.stack([xs.sum(-1), (xs * xs).sum(-1)], -1).cuda()
    a = torch.relu(torch.nn.functional.instance_norm(x.double(), eps=1e-4)).to(dtype).double()
    y_r = mops.depthwise_conv3d(a, w.double())
    wt = w.reshape(C, -1).t().contiguous().float().cuda()
    xg = ndhwc(x).to(dtype).cuda()
    y, yst = ops.dwconv3d(xg, wt, k, x_stats=st, act=ops.ACT_RELU, want_stats=True)
    tol = TOL[dtype]
    assert rel_err(ncdhw(y.float().cpu()), y_r) < tol
    yd = ncdhw(y.double().cpu()).flatten(2)
    ref_st = torch.stack([yd.sum(-1), (yd * yd).sum(-1)], -1)
    assert rel_err(yst.cpu(), ref_st) < 1e-4
    gy = torch.randn(B, C, *sp).to(dtype)
    dw = ops.dwconv3d_wgrad(xg, ndhwc(gy).to(dtype).cuda(), k, x_stats=st, act=ops.ACT_RELU)
    a.requires_grad_(False)
    wr = w.double().requires_grad_(True)
    mops.depthwise_conv3d(a, wr).backward(gy.double())
    assert rel_err(dw.t().reshape(C, 1, *k).cpu(), wr.grad) < 2 * tol
