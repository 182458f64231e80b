"""CPU: the augmentation oracle against the fixtures the unmodified reference produced, and the product's host logic
(random draws / plans of b200seg.augmentation) against the same fixtures — no kernel runs here."""
import os

import numpy as np
import pytest
import torch

import b200seg
from b200seg import augmentation as aug
from oracle import augmentation as oaug
from oracle.synth import make_volume


@pytest.fixture(scope="module")
def ops_fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "augment_ops.pt"), weights_only=False)


@pytest.fixture(scope="module")
def train_fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "augment_train.pt"), weights_only=False)


def _maxdiff(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


def test_oracle_affine_matches_reference(ops_fx):
    f = ops_fx
    ci, cl = oaug.crop(f["img"][0].numpy(), f["lab"][0, 0].numpy().astype(np.int64), f["sub_origin"], f["sub_size"])
    theta = oaug.theta_from_draws(*f["affine_draws"])
    assert np.array_equal(theta, f["theta"].numpy())
    oi, ol = oaug.scale_rotate_translate_3d(ci, cl, theta)
    assert _maxdiff(oi, f["affine_img"][0]) < 2e-5
    assert (ol != f["affine_lab"][0, 0].numpy()).mean() < 1e-3
    pi, pl = oaug.crop(oi, ol, f["out_origin"], [16, 20, 24])
    assert _maxdiff(pi, f["patch_img"][0]) < 2e-5


@pytest.mark.parametrize("name,fn", [
    ("brightness_multiply", oaug.brightness_multiply), ("gamma", oaug.gamma), ("contrast", oaug.contrast),
    ("gamma_no_retain", lambda a, g: oaug.gamma(a, g, retain_stats=False)),
    ("contrast_no_clamp", lambda a, f: oaug.contrast(a, f, preserve_range=False)),
    ("brightness_additive", oaug.brightness_additive), ("blur_k5", oaug.gaussian_blur), ("blur_k7", oaug.gaussian_blur)])
def test_oracle_intensity_ops_match_reference(ops_fx, name, fn):
    c = ops_fx[name]
    assert _maxdiff(fn(c["in"][0].numpy(), c["param"]), c["out"][0]) < 2e-5


def test_product_theta_draws_follow_the_reference(ops_fx):
    """Same numpy stream -> the same matrix the reference handed to F.affine_grid (augmentation.py:244-286)."""
    np.random.seed(11)
    for d in (36 - 24, 40 - 28, 44 - 32):          # the crop_3d draws that precede the affine in the fixture's stream
        np.random.randint(0, max(d, 1))
    theta = aug.draw_affine_theta(**ops_fx["affine_args"])
    assert torch.equal(theta, ops_fx["theta"])


def test_gaussian_kernel_is_separable():
    """The product blurs with the 1-D factor; its outer cube is the reference's dense kernel (augmentation.py:31-44)."""
    for sigma in (0.5, 0.61, 0.85, 1.0):
        k = 2 * int(np.ceil(3 * sigma)) + 1
        w = aug.gaussian_kernel_1d(k, sigma).numpy().astype(np.float64)
        dense = oaug.gaussian_kernel_3d(k, sigma)
        assert _maxdiff(np.einsum("i,j,k->ijk", w, w, w), dense) < 5e-7      # fp32 rounding of the centre weight (0.49)


def _plan_to_oracle(p):
    f = lambda t: None if t is None else float(t.reshape(-1)[0])     # noqa: E731
    return dict(sub_origin=p["sub_origin"], sub_size=p["sub_size"], theta=None if p["theta"] is None else p["theta"].numpy(),
                out_origin=p["out_origin"], brightness=f(p["brightness"]), gamma=f(p["gamma"]), contrast=f(p["contrast"]),
                flips=p["flips"], blur_sigma=p["blur_sigma"])


def test_train_plan_reproduces_the_reference_branch(train_fx):
    """TrainAugment3D.plan consumes the numpy / torch streams exactly like dataset_kits.py:116-153 does: executing its
    plan with the oracle gives the patches the reference's own functions produced under the same seeds."""
    c = train_fx["cfg"]
    img, lab = make_volume(1, *c["volume"], c["classes"], seed=c["data_seed"])
    assert abs(float(img.double().sum()) - train_fx["img_digest"]) < 1e-6 and int(lab.sum()) == train_fx["lab_digest"]
    ta = aug.TrainAugment3D(c["training_size"], scale=c["scale"], rotate=c["rotate"], translate=c["translate"])
    seen_affine = 0
    for case in train_fx["cases"]:
        np.random.seed(case["seed"])
        torch.manual_seed(case["seed"])
        p = ta.plan(c["volume"])
        assert (p["theta"] is not None) == case["affine"]
        assert (p["noise_std"] is None) == (case["noise_std"] is None)
        if case["noise_std"] is not None:
            assert p["noise_std"] == case["noise_std"]
        oi, ol = oaug.train_branch(img[0].numpy(), lab[0, 0].numpy(), _plan_to_oracle(p), c["training_size"])
        assert _maxdiff(oi, case["img"][0]) < 5e-5, case["seed"]
        assert (ol != case["lab"][0, 0].numpy()).mean() < 2e-3, case["seed"]
        seen_affine += case["affine"]
    assert seen_affine >= 2


def test_augmentation_has_no_cpu_path():
    x = torch.zeros(1, 1, 4, 4, 4)
    with pytest.raises(b200seg.B200SegError):
        aug.brightness_multiply(x)
    with pytest.raises(b200seg.B200SegError):
        aug.TrainAugment3D([2, 2, 2])(x, x.long())
