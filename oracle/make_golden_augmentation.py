"""Pin oracle/augmentation.py against the UNMODIFIED reference module (training/augmentation.py) and write
tests/golden/augment_ops.pt (per-function cases with the random parameters the reference drew) and
tests/golden/augment_train.pt (the training branch of dataset_kits.py:116-153 driven with the reference's own functions
under fixed seeds).  Runs only where /root/reference exists.
Usage:  python oracle/make_golden_augmentation.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import augmentation as oaug                       # noqa: E402
from oracle.make_golden import REF                            # noqa: E402
from oracle.synth import make_volume                          # noqa: E402


def import_reference_aug():
    sys.path.insert(0, REF)
    import training.augmentation as ref_aug
    return ref_aug


def replay(np_state, fn):
    """Run fn() from a saved numpy RNG state and put the generator back where the reference left it."""
    after = np.random.get_state()
    np.random.set_state(np_state)
    out = fn()
    np.random.set_state(after)
    return out


def replay_torch(t_state, fn):
    after = torch.get_rng_state()
    torch.set_rng_state(t_state)
    out = fn()
    torch.set_rng_state(after)
    return out


def draws_affine(scale, rotate, translate, shear):
    """The 15 numpy draws of random_scale_rotate_translate_3d in the reference's order (augmentation.py:244-267)."""
    diag = [np.random.uniform(low=1 - s, high=1 / (1 - s)) for s in scale]
    off = [np.random.uniform(-shear[i // 2], shear[i // 2]) for i in range(6)]
    tr = [np.random.uniform(-t, t) for t in translate]
    ang = [float(np.random.randint(-r, max(r, 1))) for r in rotate]
    return diag, off, tr, ang


def maxdiff(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max())


def ops_fixture(ref):
    img, lab = make_volume(1, 36, 40, 44, 4, seed=301)
    fx = {"img": img.clone(), "lab": lab.to(torch.uint8)}
    np.random.seed(11)
    torch.manual_seed(11)
    # --- crop trick + affine + centre crop
    s0 = np.random.get_state()
    ci, cl = ref.crop_3d(img, lab, [24, 28, 32], mode="random")
    org = replay(s0, lambda: [int(np.random.randint(0, max(d, 1))) for d in (36 - 24, 40 - 28, 44 - 32)])
    oi, ol = oaug.crop(img[0].numpy(), lab[0, 0].numpy(), org, [24, 28, 32])
    assert maxdiff(oi, ci[0]) == 0 and maxdiff(ol, cl[0, 0]) == 0
    args = dict(scale=0.3, rotate=30, translate=0.1, shear=0.05)
    s0 = np.random.get_state()
    ai, al = ref.random_scale_rotate_translate_3d(ci, cl, **args)
    diag, off, tr, ang = replay(s0, lambda: draws_affine([0.3] * 3, [30] * 3, [0.1] * 3, [0.05] * 3))
    theta = oaug.theta_from_draws(diag, off, tr, ang)
    oi2, ol2 = oaug.scale_rotate_translate_3d(oi, ol, theta)
    d_img, lab_mis = maxdiff(oi2, ai[0]), float((ol2 != al[0, 0].numpy()).mean())
    print("affine: image max diff %.3e, label mismatch %.3e" % (d_img, lab_mis))
    assert d_img < 2e-5 and lab_mis < 1e-3
    pi, pl = ref.crop_3d(ai, al, [16, 20, 24], mode="center")
    fx.update(sub_origin=org, sub_size=[24, 28, 32], affine_args=args, affine_draws=(diag, off, tr, ang),
              theta=torch.from_numpy(theta.copy()), affine_img=ai.clone(), affine_lab=al.to(torch.uint8),
              out_origin=[4, 4, 4], patch_img=pi.clone(), patch_lab=pl.to(torch.uint8))
    # --- intensity ops on the patch, each with the parameter the reference drew
    x = pi.clone()
    xo = pi[0].numpy().copy()

    def one(name, call, draw, orc, tol=2e-5):
        nonlocal x, xo
        ts = torch.get_rng_state()
        y = call(x.clone())
        par = replay_torch(ts, draw)
        yo = orc(xo, par)
        d = maxdiff(yo, y[0])
        print("%-22s param %.6f  max diff %.3e" % (name, par, d))
        assert d < tol, name
        fx[name] = {"param": par, "in": x.clone(), "out": y.clone()}
        x, xo = y, y[0].numpy().copy()
    one("brightness_multiply", lambda t: ref.brightness_multiply(t, multiply_range=[0.7, 1.3]),
        lambda: float(torch.rand(size=(1, 1, 1, 1, 1)) * (1.3 - 0.7) + 0.7), oaug.brightness_multiply)
    one("gamma", lambda t: ref.gamma(t, gamma_range=[0.7, 1.5]), lambda: float(torch.rand(1, 1) * (1.5 - 0.7) + 0.7), oaug.gamma)
    one("contrast", lambda t: ref.contrast(t, contrast_range=[0.65, 1.5]), lambda: float(torch.rand(1, 1) * (1.5 - 0.65) + 0.65),
        oaug.contrast)
    one("gamma_no_retain", lambda t: ref.gamma(t, gamma_range=[0.5, 2], retain_stats=False),
        lambda: float(torch.rand(1, 1) * (2 - 0.5) + 0.5), lambda a, g: oaug.gamma(a, g, retain_stats=False))
    one("contrast_no_clamp", lambda t: ref.contrast(t, contrast_range=[1.2, 1.5], preserve_range=False),
        lambda: float(torch.rand(1, 1) * (1.5 - 1.2) + 1.2), lambda a, f: oaug.contrast(a, f, preserve_range=False))
    one("brightness_additive", lambda t: ref.brightness_additive(t, std=0.2),
        lambda: float(torch.normal(0, 0.2, size=(1, 1, 1, 1, 1))), oaug.brightness_additive)
    one("blur_k5", lambda t: ref.gaussian_blur(t, sigma_range=[0.5, 0.6]), lambda: float(torch.rand(1) * 0.1 + 0.5), oaug.gaussian_blur)
    one("blur_k7", lambda t: ref.gaussian_blur(t, sigma_range=[0.8, 1.0]), lambda: float(torch.rand(1) * 0.2 + 0.8), oaug.gaussian_blur)
    for ax in (0, 1, 2):
        assert maxdiff(oaug.mirror(xo, ax), ref.mirror(x, axis=ax)[0]) == 0
    torch.save(fx, os.path.join(ROOT, "tests", "golden", "augment_ops.pt"))


TRAIN = dict(volume=(72, 70, 76), classes=3, data_seed=302, training_size=[10, 12, 14], scale=0.3, rotate=30, translate=0.1, shear=0.05)


def reference_train_branch(ref, tensor_img, tensor_lab, c):
    """dataset_kits.py:116-153 statement by statement, on the reference's own functions (the dataset class itself needs
    SimpleITK, which this image lacks); the noise step returns its std instead of adding host noise."""
    d, h, w = c["training_size"]
    if np.random.random() < 0.2:
        tensor_img, tensor_lab = ref.crop_3d(tensor_img, tensor_lab, [d + 60, h + 60, w + 60], mode="random")
        tensor_img, tensor_lab = ref.random_scale_rotate_translate_3d(tensor_img, tensor_lab, c["scale"], c["rotate"], c["translate"])
        tensor_img, tensor_lab = ref.crop_3d(tensor_img, tensor_lab, c["training_size"], mode="center")
    else:
        tensor_img, tensor_lab = ref.crop_3d(tensor_img, tensor_lab, c["training_size"], mode="random")
    tensor_img, tensor_lab = tensor_img.contiguous(), tensor_lab.contiguous()
    if np.random.random() < 0.2:
        tensor_img = ref.brightness_multiply(tensor_img, multiply_range=[0.7, 1.3])
    if np.random.random() < 0.2:
        tensor_img = ref.gamma(tensor_img, gamma_range=[0.7, 1.5])
    if np.random.random() < 0.2:
        tensor_img = ref.contrast(tensor_img, contrast_range=[0.65, 1.5])
    for axis in (2, 1, 0):
        if np.random.random() < 0.3:
            tensor_img = ref.mirror(tensor_img, axis=axis)
            tensor_lab = ref.mirror(tensor_lab, axis=axis)
    if np.random.random() < 0.2:
        tensor_img = ref.gaussian_blur(tensor_img, sigma_range=[0.5, 1.0])
    noise_std = None
    if np.random.random() < 0.2:
        noise_std = np.random.random() * 0.1
    return tensor_img, tensor_lab, noise_std


def train_fixture(ref):
    c = TRAIN
    img, lab = make_volume(1, *c["volume"], c["classes"], seed=c["data_seed"])
    cases = []
    for seed in range(400):
        np.random.seed(seed)
        torch.manual_seed(seed)
        gates = np.random.random()          # first gate only, to pick seeds cheaply
        np.random.seed(seed)
        want_affine = gates < 0.2
        if len([k for k in cases if k["affine"]]) >= 3 and want_affine:
            continue
        if len([k for k in cases if not k["affine"]]) >= 3 and not want_affine:
            continue
        oi, ol, nstd = reference_train_branch(ref, img, lab, c)
        cases.append({"seed": seed, "affine": want_affine, "img": oi.clone(), "lab": ol.to(torch.uint8), "noise_std": nstd})
        if len(cases) == 6:
            break
    print("train-branch seeds:", [(k["seed"], k["affine"]) for k in cases])
    torch.save({"cfg": c, "img_digest": float(img.double().sum()), "lab_digest": int(lab.sum()), "cases": cases},
               os.path.join(ROOT, "tests", "golden", "augment_train.pt"))


def main():
    torch.set_num_threads(8)
    ref = import_reference_aug()
    ops_fixture(ref)
    train_fixture(ref)


if __name__ == "__main__":
    main()
