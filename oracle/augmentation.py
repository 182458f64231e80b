"""TEST INFRASTRUCTURE — CPU (numpy) restatement of the reference's GPU-augmentation functions
(training/augmentation.py) with every random quantity passed in explicitly, so the CUDA kernels can be compared on
identical parameters.  Pinned against the unmodified reference by oracle/make_golden_augmentation.py
(tests/golden/augment_*.pt).  Nothing in the product imports this file.

Array conventions: image float32 [C, D, H, W]; label integer [D, H, W]."""
import math

import numpy as np


# ---- geometry -------------------------------------------------------------------------------------------------------
def affine_grid(theta, size):
    """F.affine_grid(theta[None], (1,C,D,H,W), align_corners=True) — augmentation.py:287: the base grid holds
    linspace(-1, 1, n) per axis (x fastest), the output is base @ theta^T, channels (x, y, z)."""
    D, H, W = size
    th = np.asarray(theta, dtype=np.float32).reshape(3, 4)

    def lin(n):
        return np.linspace(-1.0, 1.0, n, dtype=np.float32) if n > 1 else np.zeros(1, np.float32)
    z, y, x = np.meshgrid(lin(D), lin(H), lin(W), indexing="ij")
    base = np.stack([x, y, z, np.ones_like(x)], axis=-1)            # [D,H,W,4]
    return (base @ th.T).astype(np.float32)                           # [D,H,W,3]


def grid_sample(vol, grid, mode):
    """F.grid_sample(vol[None], grid[None], mode, padding_mode='zeros', align_corners=True) — augmentation.py:288-289.
    vol [C,D,H,W]; mode 'bilinear' (trilinear) or 'nearest' (round half to even)."""
    C, D, H, W = vol.shape
    ix = (grid[..., 0] + 1) * np.float32(0.5) * np.float32(W - 1)
    iy = (grid[..., 1] + 1) * np.float32(0.5) * np.float32(H - 1)
    iz = (grid[..., 2] + 1) * np.float32(0.5) * np.float32(D - 1)
    out = np.zeros((C,) + grid.shape[:3], dtype=np.float32)

    def fetch(zi, yi, xi):
        ok = (zi >= 0) & (zi < D) & (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
        v = vol[:, np.clip(zi, 0, D - 1), np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)]
        return np.where(ok[None], v, 0).astype(np.float32)
    if mode == "nearest":
        return fetch(np.rint(iz).astype(np.int64), np.rint(iy).astype(np.int64), np.rint(ix).astype(np.int64))
    x0, y0, z0 = np.floor(ix), np.floor(iy), np.floor(iz)
    tx, ty, tz = ix - x0, iy - y0, iz - z0
    x0, y0, z0 = x0.astype(np.int64), y0.astype(np.int64), z0.astype(np.int64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                w = (tx if dx else 1 - tx) * (ty if dy else 1 - ty) * (tz if dz else 1 - tz)
                out += fetch(z0 + dz, y0 + dy, x0 + dx) * w[None].astype(np.float32)
    return out


def scale_rotate_translate_3d(img, lab, theta):
    """random_scale_rotate_translate_3d (augmentation.py:226-291) for a given 3x4 theta."""
    grid = affine_grid(theta, img.shape[1:])
    out = grid_sample(img, grid, "bilinear")
    olab = grid_sample(lab[None].astype(np.float32), grid, "nearest")[0].astype(np.int64)
    return out, olab


def theta_from_draws(diag, off, tr, ang_deg):
    """The matrix of augmentation.py:257-286 from its 15 random numbers: S = [[sx, sxy, sxz, tx], [syx, sy, syz, ty],
    [szx, szy, sz, tz]], theta = (Rx @ Ry @ Rz @ S)[:3] in float32."""
    S = np.array([[diag[0], off[0], off[1], tr[0]], [off[2], diag[1], off[3], tr[1]], [off[4], off[5], diag[2], tr[2]],
                  [0, 0, 0, 1]], dtype=np.float32)
    ax, ay, az = [(float(a) / 180.0) * math.pi for a in ang_deg]
    Rx = np.array([[1, 0, 0, 0], [0, math.cos(ax), -math.sin(ax), 0], [0, math.sin(ax), math.cos(ax), 0], [0, 0, 0, 1]], np.float32)
    Ry = np.array([[math.cos(ay), 0, -math.sin(ay), 0], [0, 1, 0, 0], [math.sin(ay), 0, math.cos(ay), 0], [0, 0, 0, 1]], np.float32)
    Rz = np.array([[math.cos(az), -math.sin(az), 0, 0], [math.sin(az), math.cos(az), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    return (((Rx @ Ry) @ Rz) @ S)[:3]


def crop(img, lab, origin, size):
    """crop_3d / crop_around_coordinate_3d slicing (augmentation.py:338-341)."""
    z, y, x = origin
    d, h, w = size
    return img[:, z:z + d, y:y + h, x:x + w].copy(), None if lab is None else lab[z:z + d, y:y + h, x:x + w].copy()


def mirror(a, axis):
    """torch.flip(dims=[2+axis]) (augmentation.py:197) on [C,D,H,W] (image) or [D,H,W] (label)."""
    return np.flip(a, axis=(a.ndim - 3) + axis).copy()


# ---- intensity ------------------------------------------------------------------------------------------------------
def brightness_multiply(img, r):
    return (img * np.float32(r)).astype(np.float32)                    # augmentation.py:101


def brightness_additive(img, r):
    return (img + np.float32(r)).astype(np.float32)                    # augmentation.py:85


def gamma(img, g, retain_stats=True):
    """augmentation.py:115-131 on one statistics row (per_channel=False with C == 1, or one channel of per_channel)."""
    x = img.reshape(-1).astype(np.float32)
    mn, mx = x.min(), x.max()
    rng = mx - mn
    mean, std = x.mean(dtype=np.float64), x.std(ddof=1, dtype=np.float64)
    y = np.power((x - mn) / rng, np.float32(g)) * rng + mn
    if retain_stats:
        y = y - np.float32(y.mean(dtype=np.float64))
        y = y / np.float32(y.std(ddof=1, dtype=np.float64)) * np.float32(std) + np.float32(mean)
    return y.reshape(img.shape).astype(np.float32)


def contrast(img, f, preserve_range=True):
    """augmentation.py:150-166."""
    x = img.reshape(-1).astype(np.float32)
    mn, mx = x.min(), x.max()
    mean = np.float32(x.mean(dtype=np.float64))
    y = (x - mean) * np.float32(f) + mean
    if preserve_range:
        y = np.clip(y, mn, mx)
    return y.reshape(img.shape).astype(np.float32)


def gaussian_kernel_3d(kernel_size, sigma):
    """generate_3d_gaussian_kernel (augmentation.py:31-44): dense, normalised to sum 1."""
    r = np.arange(-kernel_size // 2 + 1, kernel_size // 2 + 1, dtype=np.float32)
    x, y, z = np.meshgrid(r, r, r, indexing="ij")
    k = np.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * sigma ** 2)) / (2 * math.pi * sigma ** 2) ** 1.5
    return (k / k.sum()).astype(np.float32)


def gaussian_blur(img, sigma):
    """gaussian_blur (augmentation.py:46-58): dense k^3 cross-correlation with zero padding k//2, k = 2*ceil(3 sigma)+1."""
    from scipy import ndimage
    k = 2 * math.ceil(3 * sigma) + 1
    ker = gaussian_kernel_3d(k, sigma).astype(np.float64)
    return np.stack([ndimage.correlate(c.astype(np.float64), ker, mode="constant", cval=0.0) for c in img]).astype(np.float32)


# ---- the training branch (dataset_kits.py:116-153) for a given plan -----------------------------------------------------
def train_branch(img, lab, plan, training_size):
    """plan: dict with sub_origin, sub_size, theta (or None), out_origin, brightness, gamma, contrast (scalars or None),
    flips (D, H, W), blur_sigma — in the reference's order; noise is excluded (random by construction)."""
    i, l = crop(img, lab, plan["sub_origin"], plan["sub_size"])
    if plan["theta"] is not None:
        i, l = scale_rotate_translate_3d(i, l, plan["theta"])
    i, l = crop(i, l, plan["out_origin"], training_size)
    if plan.get("brightness") is not None:
        i = brightness_multiply(i, plan["brightness"])
    if plan.get("gamma") is not None:
        i = gamma(i, plan["gamma"])
    if plan.get("contrast") is not None:
        i = contrast(i, plan["contrast"])
    fd, fh, fw = plan["flips"]
    for axis, f in ((2, fw), (1, fh), (0, fd)):
        if f:
            i, l = mirror(i, axis), mirror(l, axis)
    if plan.get("blur_sigma") is not None:
        i = gaussian_blur(i, plan["blur_sigma"])
    return i, l
