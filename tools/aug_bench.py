"""Time the GPU augmentation path (csrc/augment.cu) at BASELINE's training size — a 128^3 patch out of a (128+60)^3
crop of a 256^3 volume — against the HBM roofline and against the UNMODIFIED reference functions (baseline/_ref,
training/augmentation.py with `aug_device: gpu`, i.e. stock PyTorch CUDA ops) on the same GPU in the same run.
Usage (GPU box):  python tools/aug_bench.py > gpurun_out/aug_bench.json"""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200seg                                    # noqa: E402
from b200seg import augmentation as aug           # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def main():
    torch.cuda.set_device(0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6570.0))
    g = torch.Generator(device="cuda").manual_seed(1)
    vol = torch.randn(1, 1, 256, 256, 256, device="cuda", generator=g)
    lab = torch.randint(0, 4, (1, 1, 256, 256, 256), device="cuda", dtype=torch.uint8, generator=g)
    size, big = [128] * 3, [188] * 3
    np.random.seed(3)
    theta = aug.draw_affine_theta(0.3, 30, 0.1, 0.05)
    org = [30, 30, 30]
    Vo = 128 ** 3
    patch, _, st0 = aug.resample(vol, lab, org, big, theta, [30] * 3, size, want_stats=True)
    rows = {}

    def rec(name, us, nbytes, stock_us=None):
        rows[name] = {"us": round(us, 1), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / us / 1e3, 1),
                      "frac_of_hbm": round(nbytes / us / 1e3 / hbm, 3)}
        if stock_us is not None:
            rows[name]["stock_torch_us"] = round(stock_us, 1)
            rows[name]["speedup"] = round(stock_us / us, 2)

    # ---- ours
    t_aff = timed(lambda: aug.resample(vol, lab, org, big, theta, [30] * 3, size, (True, False, True), want_stats=True))
    t_copy = timed(lambda: aug.resample(vol, lab, org, size, None, [0] * 3, size, (True, False, True), want_stats=True))
    t_mul = timed(lambda: aug._pointwise(patch, aug.OP_MUL, a=[1.1], want_stats=True))
    t_gamma = timed(lambda: aug._gamma(patch, [1.2], 1, stats=st0, want_stats=True))
    t_con = timed(lambda: aug._contrast(patch, [1.3], 1, stats=st0))
    t_blur5 = timed(lambda: aug._blur(patch, 0.6))
    t_blur7 = timed(lambda: aug._blur(patch, 0.9))
    t_noise = timed(lambda: aug._pointwise(patch, aug.OP_NOISE, a=[0.05], b=[0.0], seed=7))

    # ---- the unmodified reference functions on the same GPU
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    stock = {}
    try:
        import training.augmentation as ref
        labl = lab.long()

        def ref_geom():
            i, l = ref.crop_3d(vol, labl, big, mode="center")
            i, l = ref.random_scale_rotate_translate_3d(i, l, 0.3, 30, 0.1)
            i, l = ref.crop_3d(i, l, size, mode="center")
            i, l = ref.mirror(i, 2), ref.mirror(l, 2)
            return ref.mirror(i, 0), ref.mirror(l, 0)

        def ref_copy():
            i, l = ref.crop_3d(vol, labl, size, mode="center")
            i, l = ref.mirror(i, 2), ref.mirror(l, 2)
            return ref.mirror(i, 0), ref.mirror(l, 0)
        stock["affine"] = timed(ref_geom, iters=5, warm=2)
        stock["copy"] = timed(ref_copy, iters=5, warm=2)
        stock["mul"] = timed(lambda: ref.brightness_multiply(patch))
        stock["gamma"] = timed(lambda: ref.gamma(patch.clone(), gamma_range=[0.7, 1.5]))
        stock["contrast"] = timed(lambda: ref.contrast(patch, contrast_range=[0.65, 1.5]))
        stock["blur5"] = timed(lambda: ref.gaussian_blur(patch, sigma_range=[0.55, 0.6]), iters=5, warm=2)
        stock["blur7"] = timed(lambda: ref.gaussian_blur(patch, sigma_range=[0.9, 0.95]), iters=5, warm=2)
        stock["noise"] = timed(lambda: ref.gaussian_noise(patch, std=0.05), iters=3, warm=1)
    except Exception as e:      # baseline/_ref absent
        stock = {"unavailable": repr(e)}

    g_ = stock.get
    rec("crop+affine+crop+2 mirrors (one gather, +stats)", t_aff, Vo * (4 + 4) + Vo * (1 + 8), g_("affine"))
    rec("crop+2 mirrors (copy branch, +stats)", t_copy, Vo * (4 + 4) + Vo * (1 + 8), g_("copy"))
    rec("brightness_multiply (+stats)", t_mul, Vo * 8, g_("mul"))
    rec("gamma retain_stats (2 passes, +stats)", t_gamma, Vo * 16, g_("gamma"))
    rec("contrast", t_con, Vo * 8, g_("contrast"))
    rec("gaussian_blur k=5", t_blur5, Vo * 8, g_("blur5"))
    rec("gaussian_blur k=7", t_blur7, Vo * 8, g_("blur7"))
    rec("gaussian_noise (Philox in kernel)", t_noise, Vo * 8, g_("noise"))
    ours = t_aff + t_mul + t_gamma + t_con + t_blur7 + t_noise
    out = {"what": "GPU augmentation of one 128^3 training patch (all ops applied), B200", "hbm_peak_GBps": hbm,
           "ops": rows, "all_ops_us": round(ours, 1)}
    if "affine" in stock:
        st_total = stock["affine"] + stock["mul"] + stock["gamma"] + stock["contrast"] + stock["blur7"] + stock["noise"]
        out["stock_all_ops_us"] = round(st_total, 1)
        out["speedup_all_ops"] = round(st_total / ours, 2)
    else:
        out["stock"] = stock
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
