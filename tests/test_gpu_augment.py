"""GPU: csrc/augment.cu through the C ABI against (a) the fixtures the unmodified reference produced
(tests/golden/augment_*.pt), (b) the numpy oracle on further seeded inputs, (c) size-independent properties at the
full 128^3 training size (identity transform == crop, flips are involutions, statistics == torch reductions)."""
import os

import numpy as np
import pytest
import torch

import b200seg
from b200seg import augmentation as aug
from oracle import augmentation as oaug
from oracle.synth import make_volume

pytestmark = pytest.mark.gpu
TOL = 2e-5          # fp32 images with |x| <= ~3: a few ulp of coordinate / reduction-order difference


@pytest.fixture(scope="module")
def ops_fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "augment_ops.pt"), weights_only=False)


@pytest.fixture(scope="module")
def train_fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "augment_train.pt"), weights_only=False)


def _md(a, b):
    return float((a.detach().double().cpu() - torch.as_tensor(np.asarray(b)).double()).abs().max())


def test_affine_matches_reference_fixture(ops_fx):
    f = ops_fx
    img, lab = f["img"].cuda(), f["lab"].cuda()
    full, full_lab, _ = aug.resample(img, lab, f["sub_origin"], f["sub_size"], f["theta"], (0, 0, 0), f["sub_size"])
    assert _md(full, f["affine_img"]) < TOL
    assert (full_lab.cpu() != f["affine_lab"].long()).float().mean() < 1e-3
    # only the centre patch (what the training branch keeps): same values, computed without the margin
    patch, patch_lab, st = aug.resample(img, lab, f["sub_origin"], f["sub_size"], f["theta"], f["out_origin"], [16, 20, 24],
                                        want_stats=True)
    assert _md(patch, f["patch_img"]) < TOL
    assert (patch_lab.cpu() != f["patch_lab"].long()).float().mean() < 1e-3
    assert patch_lab.dtype == torch.int64
    s = aug.decode_stats(st, patch.numel())
    p = f["patch_img"].double()
    assert abs(s["min"][0] - p.min().item()) < TOL and abs(s["max"][0] - p.max().item()) < TOL
    assert abs(s["mean"][0] - p.mean().item()) < TOL and abs(s["std"][0] - p.std().item()) < TOL


@pytest.mark.parametrize("name", ["brightness_multiply", "gamma", "contrast", "gamma_no_retain", "contrast_no_clamp",
                                  "brightness_additive", "blur_k5", "blur_k7"])
def test_intensity_ops_match_reference_fixture(ops_fx, name):
    c = ops_fx[name]
    x, par = c["in"].cuda(), c["param"]
    if name == "brightness_multiply":
        y = aug._pointwise(x, aug.OP_MUL, a=[par])[0]
    elif name == "brightness_additive":
        y = aug._pointwise(x, aug.OP_ADD, a=[par])[0]
    elif name.startswith("gamma"):
        y = aug._gamma(x, [par], 1, retain_stats=(name == "gamma"))[0]
    elif name.startswith("contrast"):
        y = aug._contrast(x, [par], 1, preserve_range=(name == "contrast"))[0]
    else:
        y = aug._blur(x, par)[0]
    assert _md(y, c["out"]) < TOL, name


def test_public_functions_draw_like_the_reference(ops_fx):
    """Seeding torch reproduces the reference's parameter: the public call == the fixture the reference made from the
    same generator state (the fixture script ran the ops in this order from manual_seed(11))."""
    torch.manual_seed(11)
    x = ops_fx["brightness_multiply"]["in"].cuda()
    y = aug.brightness_multiply(x, multiply_range=[0.7, 1.3])
    assert _md(y, ops_fx["brightness_multiply"]["out"]) < TOL
    y = aug.gamma(y, gamma_range=[0.7, 1.5])
    assert _md(y, ops_fx["gamma"]["out"]) < TOL
    y = aug.contrast(y, contrast_range=[0.65, 1.5])
    assert _md(y, ops_fx["contrast"]["out"]) < TOL


def test_copy_branch_flips_and_label_types():
    img, lab = make_volume(1, 20, 22, 26, 5, seed=5, in_ch=2)
    gi, gl8 = img.cuda(), lab.to(torch.uint8).cuda()
    org, size = [3, 2, 5], [12, 16, 18]
    ref_i = img[:, :, 3:15, 2:18, 5:23]
    ref_l = lab[:, :, 3:15, 2:18, 5:23]
    for mask in range(8):
        flips = (bool(mask & 1), bool(mask & 2), bool(mask & 4))
        dims = [2 + a for a in range(3) if flips[a]]
        oi, ol, st = aug.resample(gi, gl8, org, size, None, (0, 0, 0), size, flips, want_stats=True, per_channel=True)
        assert torch.equal(oi.cpu(), torch.flip(ref_i, dims) if dims else ref_i)
        assert torch.equal(ol.cpu(), torch.flip(ref_l, dims) if dims else ref_l)
        s = aug.decode_stats(st, ref_i[0, 0].numel())
        for c in range(2):
            assert s["min"][c] == ref_i[0, c].min().item() and s["max"][c] == ref_i[0, c].max().item()
            assert abs(s["mean"][c] - ref_i[0, c].double().mean().item()) < 1e-6
            assert abs(s["std"][c] - ref_i[0, c].double().std().item()) < 1e-6
    # mirror() on an image and on label maps of both widths (label-only gather), and that it is an involution
    for axis in range(3):
        assert torch.equal(aug.mirror(gi, axis).cpu(), torch.flip(img, [2 + axis]))
        assert torch.equal(aug.mirror(gl8, axis).cpu(), torch.flip(lab.to(torch.uint8), [2 + axis]))
        l64 = aug.mirror(lab.cuda(), axis)
        assert l64.dtype == torch.int64 and torch.equal(l64.cpu(), torch.flip(lab, [2 + axis]))
        assert torch.equal(aug.mirror(aug.mirror(gi, axis), axis), gi)


def test_crop_functions_follow_numpy_stream():
    img, lab = make_volume(1, 24, 20, 28, 3, seed=6)
    gi, gl = img.cuda(), lab.cuda()
    np.random.seed(3)
    z, y, x = [int(np.random.randint(0, max(d, 1))) for d in (24 - 10, 20 - 12, 28 - 14)]
    np.random.seed(3)
    ci, cl = aug.crop_3d(gi, gl, [10, 12, 14], mode="random")
    assert torch.equal(ci.cpu(), img[:, :, z:z + 10, y:y + 12, x:x + 14]) and torch.equal(cl.cpu(), lab[:, :, z:z + 10, y:y + 12, x:x + 14])
    ci, cl = aug.crop_3d(gi, gl, 8, mode="center")
    assert torch.equal(ci.cpu(), img[:, :, 8:16, 6:14, 10:18])
    ci, cl = aug.crop_around_coordinate_3d(gi, gl, [8, 8, 8], (12, 10, 14), mode="center")
    assert torch.equal(ci.cpu(), img[:, :, 8:16, 6:14, 10:18])
    with pytest.raises(b200seg.B200SegError):           # patch outside the sub-volume: EINVAL, never a wild gather
        aug.resample(gi, gl, (0, 0, 0), [8, 8, 8], None, (4, 4, 4), [8, 8, 8])


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_affine_vs_oracle_random_cases(seed):
    """Odd extents, a two-channel image, strong transforms (large parts of the grid fall outside -> zeros padding)."""
    rng = np.random.RandomState(seed)
    D, H, W = int(rng.randint(9, 20)), int(rng.randint(9, 24)), int(rng.randint(9, 28))
    img, lab = make_volume(1, D, H, W, 4, seed=seed, in_ch=2)
    np.random.seed(seed)
    theta = aug.draw_affine_theta(scale=0.4, rotate=60, translate=0.3, shear=0.1)
    oi, ol = oaug.scale_rotate_translate_3d(img[0].numpy(), lab[0, 0].numpy(), theta.numpy())
    gi, gl, _ = aug.resample(img.cuda(), lab.cuda(), (0, 0, 0), [D, H, W], theta, (0, 0, 0), [D, H, W])
    assert _md(gi[0], oi) < TOL
    assert (gl[0, 0].cpu().numpy() != ol).mean() < 2e-3
    assert float((gi == 0).float().mean()) > 0.01           # the zero-padding region is exercised


def test_blur_ragged_tiles_vs_oracle():
    img, _ = make_volume(1, 11, 13, 37, 2, seed=9, in_ch=2)      # not multiples of the 8x8x32 tile
    for sigma in (0.5, 0.9):
        y, st = aug._blur(img.cuda(), sigma, want_stats=True, per_channel=True)
        o = oaug.gaussian_blur(img[0].numpy(), sigma)
        assert _md(y[0], o) < TOL
        s = aug.decode_stats(st, img[0, 0].numel())
        for c in range(2):
            assert abs(s["mean"][c] - float(o[c].mean(dtype=np.float64))) < 1e-5
            assert abs(s["min"][c] - float(o[c].min())) < TOL


def test_noise_moments_and_determinism():
    x = torch.zeros(1, 1, 48, 64, 64, device="cuda")
    a = aug._pointwise(x, aug.OP_NOISE, a=[0.5], b=[0.1], seed=1234)[0]
    b = aug._pointwise(x, aug.OP_NOISE, a=[0.5], b=[0.1], seed=1234)[0]
    c = aug._pointwise(x, aug.OP_NOISE, a=[0.5], b=[0.1], seed=1235)[0]
    assert torch.equal(a, b) and not torch.equal(a, c)
    z = (a.double() - 0.1) / 0.5
    n = z.numel()
    assert abs(z.mean().item()) < 5 / n ** 0.5 and abs(z.std().item() - 1) < 0.01
    assert abs((z ** 3).mean().item()) < 0.02 and abs((z ** 4).mean().item() - 3) < 0.05
    # neighbouring elements are uncorrelated (they share one Philox block)
    f = z.flatten()
    assert abs((f[:-1] * f[1:]).mean().item()) < 0.01
    torch.manual_seed(0)
    y = aug.gaussian_noise(torch.ones(1, 1, 8, 8, 8, device="cuda"), std=0.05)
    assert abs(y.mean().item() - 1) < 0.02


def test_train_branch_reproduces_reference_fixture(train_fx):
    c = train_fx["cfg"]
    img, lab = make_volume(1, *c["volume"], c["classes"], seed=c["data_seed"])
    gi, gl = img.cuda(), lab.to(torch.uint8).cuda()
    ta = aug.TrainAugment3D(c["training_size"], scale=c["scale"], rotate=c["rotate"], translate=c["translate"])
    for case in train_fx["cases"]:
        np.random.seed(case["seed"])
        torch.manual_seed(case["seed"])
        oi, ol = ta(gi, gl)
        assert (ol.cpu() != case["lab"].long()).float().mean() < 2e-3, case["seed"]
        if case["noise_std"] is None:
            assert _md(oi, case["img"]) < 5e-5, case["seed"]
        else:       # the reference fixture carries no noise; ours has N(0, std): the difference must be exactly that
            d = (oi.cpu().double() - case["img"].double()).flatten()
            assert abs(d.std().item() / case["noise_std"] - 1) < 0.15 and abs(d.mean().item()) < 0.5 * case["noise_std"]


def test_full_size_properties():
    """128^3 patch out of a 188^3 volume (the reference's size+60 crop trick at BASELINE's training size)."""
    g = torch.Generator(device="cuda").manual_seed(4)
    vol = torch.randn(1, 1, 188, 188, 188, device="cuda", generator=g)
    lab = torch.randint(0, 4, (1, 1, 188, 188, 188), device="cuda", dtype=torch.uint8, generator=g)
    ident = torch.tensor([[1., 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]])
    a_i, a_l, st = aug.resample(vol, lab, (0, 0, 0), [188] * 3, ident, (30, 30, 30), [128] * 3, want_stats=True)
    ref = vol[:, :, 30:158, 30:158, 30:158]
    assert (a_i - ref).abs().max().item() < 2e-3            # interpolation weights ~1e-5 off the lattice x |x| <= 5
    assert torch.equal(a_l, lab[:, :, 30:158, 30:158, 30:158].long())
    s = aug.decode_stats(st, ref.numel())
    assert abs(s["mean"][0] - a_i.double().mean().item()) < 1e-6 and abs(s["std"][0] - a_i.double().std().item()) < 1e-6
    assert s["min"][0] == a_i.min().item() and s["max"][0] == a_i.max().item()
    # flips are involutions and commute with the gather
    f_i, f_l, _ = aug.resample(vol, lab, (0, 0, 0), [188] * 3, ident, (30, 30, 30), [128] * 3, flips=(True, False, True))
    assert torch.equal(torch.flip(f_i, [2, 4]), a_i) and torch.equal(torch.flip(f_l, [2, 4]), a_l)
    # linearity of the blur and preservation of the mean away from the border (weights sum to 1)
    b1 = aug._blur(a_i, 0.8)[0]
    b2 = aug._blur(a_i * 2 + 1, 0.8)[0]
    inner = (slice(None), slice(None), slice(8, 120), slice(8, 120), slice(8, 120))
    assert (b2[inner] - (2 * b1[inner] + 1)).abs().max().item() < 1e-4
    # gamma(retain_stats) keeps mean / std; contrast(preserve_range) keeps the range
    st0 = aug.image_stats(a_i)
    y, st1 = aug._gamma(a_i, [1.3], 1, stats=st0, want_stats=True)
    s0, s1 = aug.decode_stats(st0, a_i.numel()), aug.decode_stats(st1, a_i.numel())
    assert abs(s0["mean"][0] - s1["mean"][0]) < 1e-4 and abs(s0["std"][0] - s1["std"][0]) < 1e-4
    yc, stc = aug._contrast(a_i, [1.5], 1, stats=st0, want_stats=True)
    sc = aug.decode_stats(stc, a_i.numel())
    assert sc["min"][0] >= s0["min"][0] and sc["max"][0] <= s0["max"][0]
