"""GPU augmentation on the B200 behind the reference's ``training/augmentation.py`` interface.

The reference's ``aug_device: gpu`` path runs every augmentation as a chain of stock PyTorch ops per sample
(``training/dataset/dim3/dataset_kits.py:116-153``): slice + ``.contiguous()``, ``F.affine_grid`` + two
``F.grid_sample`` over a (size+60)^3 sub-volume, a second crop, three ``torch.flip`` copies, min/max/mean/std reductions
per intensity op, a dense 5^3/7^3 ``F.conv3d`` blur and a host-generated noise volume copied to the device.  Here:

  * every function of that module has a same-named counterpart with the same signature, the same random draws (same
    generators, same order — seeding ``numpy``/``torch`` reproduces the reference's decisions) and kernels from
    ``csrc/augment.cu`` instead of library calls;
  * :class:`TrainAugment3D` is the whole training branch as a *plan + 2..7 launches*: all random numbers are drawn first
    (they never depend on data), then ONE gather produces the final patch (crop -> affine -> centre crop -> mirrors), and
    each intensity op reads the statistics its predecessor left on the device — no reduction passes, no host sync.

Images are ``[1, C, D, H, W]`` fp32 CUDA tensors, label maps ``[1, 1, D, H, W]`` uint8 or int64.  There is no CPU path."""
import math

import numpy as np
import torch

from ._lib import B200SegError, call
from .ops import _need_cuda, _stream

OP_MUL, OP_ADD, OP_GAMMA_POW, OP_RENORM, OP_CONTRAST, OP_NOISE, OP_STATS = range(7)

_I3 = torch.int32


# ----------------------------------------------------------------------------- device statistics rows
_STAT_TEMPLATES = {}


def new_stats(rows, device):
    """[rows, 4] int64 = {min key, max key, sum (double bits), sumsq (double bits)} initialised empty (include/b200seg.h)."""
    key = (rows, str(device))
    t = _STAT_TEMPLATES.get(key)
    if t is None:
        t = torch.tensor([[0xFFFFFFFF, 0, 0, 0]] * rows, dtype=torch.int64).to(device)
        _STAT_TEMPLATES[key] = t
    return t.clone()


def decode_stats(stats, n):
    """Host view of a statistics tensor: dict of per-row min / max / mean / std (unbiased).  Synchronises — tests only."""
    s = stats.cpu()
    keys = s[:, :2].numpy().astype(np.uint64).astype(np.uint32)
    bits = np.where(keys & np.uint32(0x80000000), keys & np.uint32(0x7FFFFFFF), ~keys).astype(np.uint32)
    mnmx = bits.view(np.float32)
    sums = s[:, 2:].contiguous().view(torch.float64).numpy()
    mean = sums[:, 0] / n
    var = (sums[:, 1] - n * mean * mean) / max(n - 1, 1)
    return {"min": mnmx[:, 0], "max": mnmx[:, 1], "mean": mean, "std": np.sqrt(np.maximum(var, 0))}


def _img(t):
    _need_cuda(t)
    if t.dim() != 5 or t.shape[0] != 1:
        raise ValueError("expected a [1, C, D, H, W] volume (2D augmentation is outside the B200 hot path)")
    if t.dtype != torch.float32:
        raise TypeError("images are fp32 (the reference augments before autocast), got %s" % t.dtype)
    return t.contiguous()


def _lab(t):
    _need_cuda(t)
    if t.dtype not in (torch.uint8, torch.int64):
        t = t.long()
    return t.contiguous()


def _rows(C, per_channel):
    return C if per_channel else 1


def _host_f(vals):
    return (torch.as_tensor(vals, dtype=torch.float32).reshape(-1).contiguous())


def _pointwise(x, op, a=None, b=None, rows=1, stats_in=None, stats_in2=None, want_stats=False, seed=0, out=True):
    x = _img(x)
    n = x.numel() // rows
    y = torch.empty_like(x) if out else None
    so = new_stats(rows, x.device) if want_stats else None
    ah = None if a is None else _host_f(a)
    bh = None if b is None else _host_f(b)
    call("b200seg_aug_pointwise", x.data_ptr(), None if y is None else y.data_ptr(), rows, n, op,
         None if ah is None else ah.data_ptr(), None if bh is None else bh.data_ptr(),
         None if stats_in is None else stats_in.data_ptr(), None if stats_in2 is None else stats_in2.data_ptr(),
         None if so is None else so.data_ptr(), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream())
    return y, so


def image_stats(tensor_img, per_channel=False):
    """{min, max, sum, sum^2} of an image as a device statistics tensor (one pass)."""
    rows = _rows(tensor_img.shape[1], per_channel)
    return _pointwise(tensor_img, OP_STATS, rows=rows, want_stats=True, out=False)[1]


# ----------------------------------------------------------------------------- geometry
def resample(tensor_img, tensor_lab, sub_origin, sub_size, theta, out_origin, out_size, flips=(False, False, False),
             want_stats=False, per_channel=False, out_label_dtype=torch.int64):
    """The fused gather (``b200seg_aug_resample``).  theta: [3,4] float tensor / array (affine branch) or None (copy)."""
    img = None if tensor_img is None else _img(tensor_img)
    lab = None if tensor_lab is None else _lab(tensor_lab)
    ref = img if img is not None else lab
    C = 0 if img is None else img.shape[1]
    src = torch.tensor(list(ref.shape[2:]), dtype=_I3)
    so, ss = torch.tensor(list(sub_origin), dtype=_I3), torch.tensor(list(sub_size), dtype=_I3)
    oo, os_ = torch.tensor(list(out_origin), dtype=_I3), torch.tensor(list(out_size), dtype=_I3)
    th = None if theta is None else torch.as_tensor(theta, dtype=torch.float32).reshape(12).contiguous()
    out_img = None if img is None else torch.empty(1, C, *out_size, dtype=torch.float32, device=ref.device)
    out_lab = None if lab is None else torch.empty(1, 1, *out_size, dtype=out_label_dtype, device=ref.device)
    rows = _rows(C, per_channel)
    st = new_stats(rows, ref.device) if want_stats else None
    mask = (1 if flips[0] else 0) | (2 if flips[1] else 0) | (4 if flips[2] else 0)
    call("b200seg_aug_resample", None if img is None else img.data_ptr(), None if lab is None else lab.data_ptr(),
         0 if lab is None else lab.element_size(), C, src.data_ptr(), so.data_ptr(), ss.data_ptr(),
         None if th is None else th.data_ptr(), oo.data_ptr(), os_.data_ptr(), mask,
         None if out_img is None else out_img.data_ptr(),
         None if out_lab is None else out_lab.data_ptr(), 0 if out_lab is None else out_lab.element_size(),
         None if st is None else st.data_ptr(), rows, _stream())
    return out_img, out_lab, st


def _triple(v):
    return [v] * 3 if isinstance(v, (int, float)) else list(v)


def draw_affine_theta(scale=0.3, rotate=45, translate=0.1, shear=0.05):
    """The random 3x4 matrix of ``random_scale_rotate_translate_3d`` (augmentation.py:226-286): same numpy draws in the
    same order (3 scales U[1-s, 1/(1-s)], 6 shears, 3 translations, 3 integer angles), same fp32 products
    Rx @ Ry @ Rz @ S.  Axis convention of the reference: arguments in [z, y, x] order, matrix rows in (x, y, z)."""
    scale, translate, rotate, shear = _triple(scale), _triple(translate), _triple(rotate), _triple(shear)
    diag = [np.random.uniform(low=1 - s, high=1 / (1 - s)) for s in scale]
    off = [np.random.uniform(-shear[i // 2], shear[i // 2]) for i in range(6)]     # xy, xz, yx, yz, zx, zy
    tr = [np.random.uniform(-t, t) for t in translate]
    S = torch.tensor([[diag[0], off[0], off[1], tr[0]],
                      [off[2], diag[1], off[3], tr[1]],
                      [off[4], off[5], diag[2], tr[2]],
                      [0, 0, 0, 1]]).float()
    ang = [(float(np.random.randint(-r, max(r, 1))) / 180.) * math.pi for r in rotate]
    mats = []
    for axis, a in enumerate(ang):
        c, s = math.cos(a), math.sin(a)
        i, j = [(1, 2), (0, 2), (0, 1)][axis]          # plane the rotation acts in: about x -> (y,z); y -> (x,z); z -> (x,y)
        R = [[1.0 if r == q else 0.0 for q in range(4)] for r in range(4)]
        R[i][i], R[i][j], R[j][i], R[j][j] = c, -s, s, c
        mats.append(torch.tensor(R).float())
    theta = torch.mm(torch.mm(torch.mm(mats[0], mats[1]), mats[2]), S)
    return theta[0:3, :].contiguous()


def random_scale_rotate_translate_3d(tensor_img, tensor_lab, scale=0.3, rotate=45, translate=0.1, shear=0.05):
    """augmentation.py:226-291 on the whole input volume."""
    theta = draw_affine_theta(scale, rotate, translate, shear)
    size = list(tensor_img.shape[2:])
    img, lab, _ = resample(tensor_img, tensor_lab, (0, 0, 0), size, theta, (0, 0, 0), size)
    return img, lab


def _crop_origin_random(shape, crop_size):
    diffs = [s - c for s, c in zip(shape, crop_size)]
    return [int(np.random.randint(0, max(d, 1))) for d in diffs]      # z, y, x — the reference's draw order


def crop_3d(tensor_img, tensor_lab, crop_size, mode):
    """augmentation.py:320-343 (slicing clamps at the volume border, as Python slices do)."""
    assert mode in ['random', 'center'], "Invalid Mode, should be 'random' or 'center'"
    crop_size = _triple(crop_size) if isinstance(crop_size, int) else list(crop_size)
    shape = list(tensor_img.shape[2:])
    if mode == 'random':
        org = _crop_origin_random(shape, crop_size)
    else:
        org = [(s - c) // 2 for s, c in zip(shape, crop_size)]
    if min(org) < 0:
        raise ValueError("centre crop larger than the volume (the reference's negative slice start is not supported)")
    size = [min(c, s - o) for c, s, o in zip(crop_size, shape, org)]
    img, lab, _ = resample(tensor_img, tensor_lab, org, size, None, (0, 0, 0), size)
    return img, lab


def crop_around_coordinate_3d(tensor_img, tensor_lab, crop_size, coordinate, mode):
    """augmentation.py:346-383."""
    assert mode in ['random', 'center'], "Invalid Mode, should be 'random' or 'center'"
    crop_size = _triple(crop_size) if isinstance(crop_size, int) else list(crop_size)
    shape = list(tensor_img.shape[2:])
    org = []
    for c, s, k in zip(coordinate, shape, crop_size):
        if mode == 'random':
            lo, hi = max(0, c - k), min(s - k, c + k)
            org.append(int(np.random.randint(lo, hi)))
        else:
            org.append(min(max(0, c - math.ceil(k / 2)), s - k))
    size = [min(c, s - o) for c, s, o in zip(crop_size, shape, org)]
    img, lab, _ = resample(tensor_img, tensor_lab, org, size, None, (0, 0, 0), size)
    return img, lab


def mirror(tensor_img, axis=0):
    """torch.flip(dims=[2+axis]) (augmentation.py:176-197) for an image or a label map."""
    assert axis in [0, 1, 2], "axis should be either 0, 1 or 2 for volume images"
    flips = [axis == 0, axis == 1, axis == 2]
    size = list(tensor_img.shape[2:])
    if tensor_img.dtype == torch.float32:
        return resample(tensor_img, None, (0, 0, 0), size, None, (0, 0, 0), size, flips)[0]
    lab = _lab(tensor_img)      # label map: the label slot of the gather alone
    return resample(None, lab, (0, 0, 0), size, None, (0, 0, 0), size, flips, out_label_dtype=lab.dtype)[1]


# ----------------------------------------------------------------------------- intensity
def gaussian_noise(tensor_img, std, mean=0):
    """augmentation.py:14-16.  The reference draws the noise volume on the host and copies it; here the normals come from
    a counter-based Philox generator in the kernel, keyed by one integer drawn from torch's CPU generator."""
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return _pointwise(tensor_img, OP_NOISE, a=[std], b=[mean], rows=1, seed=seed)[0]


def gaussian_kernel_1d(kernel_size, sigma):
    """1-D factor of generate_3d_gaussian_kernel (augmentation.py:31-44): the normalised dense kernel is its outer cube."""
    r = torch.arange(-kernel_size // 2 + 1, kernel_size // 2 + 1, dtype=torch.float32)
    w = torch.exp(-(r ** 2) / (2 * float(sigma) ** 2))
    return (w / w.sum()).contiguous()


def _blur(tensor_img, sigma, want_stats=False, per_channel=False):
    x = _img(tensor_img)
    kernel_size = 2 * math.ceil(3 * sigma) + 1
    w = gaussian_kernel_1d(kernel_size, sigma)
    y = torch.empty_like(x)
    rows = _rows(x.shape[1], per_channel)
    so = new_stats(rows, x.device) if want_stats else None
    _, C, D, H, W = x.shape
    call("b200seg_aug_gaussian_blur", x.data_ptr(), y.data_ptr(), C, D, H, W, w.data_ptr(), kernel_size,
         None if so is None else so.data_ptr(), rows, _stream())
    return y, so


def gaussian_blur(tensor_img, sigma_range=[0.5, 1.0]):
    """augmentation.py:46-64."""
    sigma = float(torch.rand(1) * (sigma_range[1] - sigma_range[0]) + sigma_range[0])
    return _blur(tensor_img, sigma)[0]


def brightness_additive(tensor_img, std, mean=0, per_channel=False):
    """augmentation.py:66-85."""
    C = tensor_img.shape[1] if per_channel else 1
    r = torch.normal(mean, std, size=(1, C, 1, 1, 1))
    return _pointwise(tensor_img, OP_ADD, a=r, rows=C)[0]


def brightness_multiply(tensor_img, multiply_range=[0.7, 1.3], per_channel=False):
    """augmentation.py:88-101."""
    assert multiply_range[1] > multiply_range[0], 'Invalid range'
    C = tensor_img.shape[1] if per_channel else 1
    r = torch.rand(size=(1, C, 1, 1, 1)) * (multiply_range[1] - multiply_range[0]) + multiply_range[0]
    return _pointwise(tensor_img, OP_MUL, a=r, rows=C)[0]


def _check_rows(C, per_channel):
    if C > 1 and not per_channel:
        # the reference broadcasts a [C,1] random vector against a [1,N] view and then fails to reshape
        raise ValueError("multi-channel images need per_channel=True (the reference's own view() fails otherwise)")
    return C


def _gamma(tensor_img, g, rows, stats=None, retain_stats=True, want_stats=False):
    if stats is None:
        stats = image_stats(tensor_img, per_channel=rows > 1)
    y, s1 = _pointwise(tensor_img, OP_GAMMA_POW, a=g, rows=rows, stats_in=stats, want_stats=retain_stats or want_stats)
    if retain_stats:
        y, s1 = _pointwise(y, OP_RENORM, rows=rows, stats_in=s1, stats_in2=stats, want_stats=want_stats)
    return y, s1


def gamma(tensor_img, gamma_range=(0.5, 2), per_channel=False, retain_stats=True):
    """augmentation.py:104-137."""
    C = _check_rows(tensor_img.shape[1], per_channel)
    g = torch.rand(C, 1) * (gamma_range[1] - gamma_range[0]) + gamma_range[0]
    return _gamma(tensor_img, g, C, retain_stats=retain_stats)[0]


def _contrast(tensor_img, f, rows, stats=None, preserve_range=True, want_stats=False):
    if stats is None:
        stats = image_stats(tensor_img, per_channel=rows > 1)
    return _pointwise(tensor_img, OP_CONTRAST, a=f, b=[1.0 if preserve_range else 0.0] * rows, rows=rows, stats_in=stats,
                      want_stats=want_stats)


def contrast(tensor_img, contrast_range=(0.65, 1.5), per_channel=False, preserve_range=True):
    """augmentation.py:139-173."""
    C = _check_rows(tensor_img.shape[1], per_channel)
    f = torch.rand(C, 1) * (contrast_range[1] - contrast_range[0]) + contrast_range[0]
    return _contrast(tensor_img, f, C, preserve_range=preserve_range)[0]


# ----------------------------------------------------------------------------- the training branch as one plan
class TrainAugment3D:
    """The ``mode == 'train'`` branch of the 3D datasets' ``__getitem__`` (dataset_kits.py:116-153; the other 3D
    datasets use the same sequence): crop trick + affine (p=0.2) or random crop, brightness / gamma / contrast (p=0.2
    each), mirrors about W, H, D (p=0.3 each), blur (p=0.2), noise (p=0.2).

    ``plan()`` consumes ``np.random`` / ``torch`` CPU random numbers exactly as the reference does (same calls, same
    order), ``apply()`` executes a plan on the device.  Mirrors are folded into the gather: the intensity ops between
    the geometry and the flips are voxelwise with whole-image statistics, the blur kernel is symmetric, so moving the
    flips forward changes nothing but floating-point summation order."""

    def __init__(self, training_size, scale=0.3, rotate=45, translate=0.1, shear=0.05, margin=60):
        self.size = list(training_size)
        self.scale, self.rotate, self.translate, self.shear, self.margin = scale, rotate, translate, shear, margin

    def plan(self, volume_shape, channels=1):
        shape = list(volume_shape)
        p = {}
        if np.random.random() < 0.2:
            big = [s + self.margin for s in self.size]
            org = _crop_origin_random(shape, big)
            sub = [min(b, s - o) for b, s, o in zip(big, shape, org)]
            p["sub_origin"], p["sub_size"] = org, sub
            p["theta"] = draw_affine_theta(self.scale, self.rotate, self.translate, self.shear)
            p["out_origin"] = [(s - c) // 2 for s, c in zip(sub, self.size)]
        else:
            org = _crop_origin_random(shape, self.size)
            p["sub_origin"], p["sub_size"], p["theta"], p["out_origin"] = org, list(self.size), None, [0, 0, 0]
        if min(p["out_origin"]) < 0 or any(o + c > s for o, c, s in zip(p["out_origin"], self.size, p["sub_size"])):
            raise ValueError("volume %s is smaller than the training size %s" % (shape, self.size))
        C = channels
        p["brightness"] = (torch.rand(size=(1, 1, 1, 1, 1)) * 0.6 + 0.7) if np.random.random() < 0.2 else None
        p["gamma"] = (torch.rand(C, 1) * (1.5 - 0.7) + 0.7) if np.random.random() < 0.2 else None
        p["contrast"] = (torch.rand(C, 1) * (1.5 - 0.65) + 0.65) if np.random.random() < 0.2 else None
        fw = np.random.random() < 0.3      # axis=2
        fh = np.random.random() < 0.3      # axis=1
        fd = np.random.random() < 0.3      # axis=0
        p["flips"] = (fd, fh, fw)
        p["blur_sigma"] = float(torch.rand(1) * 0.5 + 0.5) if np.random.random() < 0.2 else None
        if np.random.random() < 0.2:
            p["noise_std"] = np.random.random() * 0.1
            p["noise_seed"] = int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            p["noise_std"] = None
        return p

    def apply(self, tensor_img, tensor_lab, p):
        C = tensor_img.shape[1]
        need_stats = p["gamma"] is not None or p["contrast"] is not None
        img, lab, st = resample(tensor_img, tensor_lab, p["sub_origin"], p["sub_size"], p["theta"], p["out_origin"],
                                self.size, p["flips"], want_stats=need_stats and p["brightness"] is None)
        if p["brightness"] is not None:
            img, st = _pointwise(img, OP_MUL, a=p["brightness"], rows=1, want_stats=need_stats)
        if p["gamma"] is not None:
            _check_rows(C, False)
            img, st = _gamma(img, p["gamma"], 1, stats=st, want_stats=p["contrast"] is not None)
        if p["contrast"] is not None:
            _check_rows(C, False)
            img, _ = _contrast(img, p["contrast"], 1, stats=st)
        if p["blur_sigma"] is not None:
            img, _ = _blur(img, p["blur_sigma"])
        if p["noise_std"] is not None:
            img, _ = _pointwise(img, OP_NOISE, a=[p["noise_std"]], b=[0.0], rows=1, seed=p["noise_seed"])
        return img, lab

    def __call__(self, tensor_img, tensor_lab):
        if not tensor_img.is_cuda:
            raise B200SegError("b200seg.augmentation runs on a B200 only — there is no CPU fallback")
        return self.apply(tensor_img, tensor_lab, self.plan(tensor_img.shape[2:], tensor_img.shape[1]))
