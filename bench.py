#!/usr/bin/env python
"""bench.py — training voxels/s of the 3D ResUNet hot path on synthetic 128^3 volumes (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = the body of the reference's train_epoch (train_ddp.py:171-215): zero_grad -> autocast fwd ->
CE+Dice -> GradScaler backward (+DDP all-reduce) -> AdamW -> EMA.  Prints ONE JSON line (rank 0).
  value     : voxels/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       : same through the public API from pinned HOST buffers (H2D inside the timed region) plus the
              per-step loss.item() D2H the reference loop does (train_ddp.py:213)
  roofline  : dominant kernel (conv3d forward of the costliest layer shape) timed alone with CUDA events
  cpu_baseline : the reference-pinned oracle (oracle/) on the host cores, bounded sample
--impl reference times the oracle port on the host CPU (the reference is pure Python/PyTorch).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scale list, kernel list, classes, ce weight, (B, D, H, W))
    "resunet_acdc_128": ([[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]],
                         [[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]], 4, [0.5, 1, 1, 1], (1, 128, 128, 128)),
    "resunet_iso_128": ([[2, 2, 2]] * 4, [[3, 3, 3]] * 5, 4, [0.5, 1, 1, 1], (1, 128, 128, 128)),
    "resunet_kits_160": ([[2, 2, 2]] * 4, [[3, 3, 3]] * 5, 3, [0.5, 1, 2], (2, 160, 160, 80)),
    # BASELINE.json configs[2]: MedFormer, config/bcv/medformer_3d.yaml:9-28,38-39,49 on a 96^3 crop, AMP
    "medformer_bcv_96": (dict(map_size=[3, 3, 3], conv_num=[2, 0, 0, 0, 0, 0, 2, 2], trans_num=[0, 2, 4, 6, 4, 2, 0, 0],
                              num_heads=[1, 4, 8, 10, 8, 4, 1, 1], fusion_depth=2, fusion_dim=320, fusion_heads=10,
                              kernel_size=[[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
                              scale=[[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]], aux_loss=True),
                         None, 14, [0.5] + [1.0] * 13, (1, 96, 96, 96)),
    # BASELINE.json configs[3]: SwinUNETR([128]*3, 1, 16, feature_size=48) (model/utils.py:111 with base_chan 48 as in
    # config/{bcv,kits,lits}/swin_unetr_3d.yaml; 16 classes per config/amos_ct/resunet_3d.yaml:3), AMP
    "swin_unetr_amos_128": ("swin", 48, 16, [0.5] + [1.0] * 15, (1, 128, 128, 128)),
}
BASE = 32
METRIC = "3D-UNet (ResBasicBlock) training voxels/sec, synthetic 128^3"
METRIC_MEDFORMER = "3D MedFormer training voxels/sec, synthetic 96^3"
METRIC_SWIN = "SwinUNETR training voxels/sec, synthetic 128^3"
AUX_WEIGHT = [0.5, 0.5]


def is_medformer(wl):
    return isinstance(wl[0], dict)


def is_swin(wl):
    return wl[0] == "swin"


def metric_of(wl):
    return METRIC_SWIN if is_swin(wl) else (METRIC_MEDFORMER if is_medformer(wl) else METRIC)


def model_name(wl):
    if is_swin(wl):
        return "SwinUNETR (reference SwinUNETR([128,128,128], 1, 16, feature_size=48), depths (2,2,2,0), window 7^3)"
    if is_medformer(wl):
        return "MedFormer-3D BCV config (reference MedFormer(1,14,32,...,norm='in',act='relu',aux_loss=True))"
    return "ResUNet-3D base32 BasicBlock IN (reference UNet(1,32,...,block='BasicBlock',norm='in'))"


def oracle_state(wl):
    """Seeded synthetic weights keyed like the reference's state_dict."""
    import torch
    from oracle import unet3d as ounet
    if is_swin(wl):
        from oracle import swin_unetr as osw
        sd = ounet.make_state_dict(osw.swin_unetr_param_shapes(1, wl[2], wl[1]), seed=7)
        for k in sd:
            if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k.endswith("norm.weight"):
                sd[k] = 1.0 + 0.1 * sd[k] / sd[k].abs().max()
        return sd
    if is_medformer(wl):
        import b200seg
        with torch.device("meta"):      # parameter names / shapes straight from the module tree (no allocation)
            probe = b200seg.MedFormer(1, wl[2], BASE, conv_block="BasicBlock", expansion=4, attn_drop=0, proj_drop=0,
                                      proj_type="depthwise", norm="in", act="relu", **wl[0])
        shapes = {k: tuple(v.shape) for k, v in probe.state_dict().items()}
        sd = ounet.make_state_dict(shapes, seed=7)
        for k in sd:
            if k.endswith("norm.weight"):
                sd[k] = 1.0 + 0.1 * sd[k] / sd[k].abs().max()
        return sd
    return ounet.make_state_dict(ounet.unet_param_shapes(1, BASE, wl[2], wl[1], "BasicBlock"), seed=7)


def oracle_loss(wl, sd, img, lab, w):
    from oracle import losses as olosses
    if is_swin(wl):
        from oracle import swin_unetr as osw
        return olosses.total_loss(osw.swin_unetr_forward(sd, img), lab, w)
    if is_medformer(wl):
        from oracle import medformer as omed
        return olosses.total_loss(omed.medformer_forward(sd, img, wl[0]), lab, w, AUX_WEIGHT)
    from oracle import unet3d as ounet
    return olosses.total_loss(ounet.unet_forward(sd, img, wl[0], wl[1], "BasicBlock"), lab, w)


def conv_layers(scale, kernel, classes, B, D, H, W, base=BASE):
    """(Cin, Cout, k, D, H, W, count) of every conv3d of the ResUNet, with fused conv1+shortcut listed apart."""
    ch = [base, 2 * base, 4 * base, 8 * base, 10 * base]
    dims = [(D, H, W)]
    for s in scale:
        d = dims[-1]
        dims.append((d[0] // s[0], d[1] // s[1], d[2] // s[2]))
    L = []

    def block(ci, co, k, dm):
        L.append((ci, co, k, dm))
        L.append((co, co, k, dm))
        if ci != co:
            L.append((ci, co, k, dm))
    L.append((1, base, kernel[0], dims[0]))
    block(base, base, kernel[0], dims[0])
    for i in range(4):
        block(ch[i], ch[i + 1], kernel[i + 1], dims[i + 1])
        block(ch[i + 1], ch[i + 1], kernel[i + 1], dims[i + 1])
    for j in range(4):
        ci, co = ch[4 - j], ch[3 - j]
        block(ci + co, co, kernel[3 - j], dims[3 - j])
        block(co, co, kernel[3 - j], dims[3 - j])
    L.append((base, classes, [1, 1, 1], dims[0]))
    return L


def swin_conv_layers(fs, classes, D, H, W):
    """3x3x3 / 1x1x1 conv3d layers of SwinUNETR's monai blocks (call sites swin_unetr.py:129-228) as (Cin, Cout, k, dims)."""
    L = []
    k3, k1 = [3, 3, 3], [1, 1, 1]

    def res(ci, co, dm):
        L.append((ci, co, k3, dm)); L.append((co, co, k3, dm))
        if ci != co:
            L.append((ci, co, k1, dm))
    dims = [(D >> i, H >> i, W >> i) for i in range(6)]
    res(1, fs, dims[0]); res(fs, fs, dims[1]); res(2 * fs, 2 * fs, dims[2]); res(4 * fs, 4 * fs, dims[3]); res(16 * fs, 16 * fs, dims[5])
    for i, (ci, co) in enumerate(((16 * fs, 8 * fs), (8 * fs, 4 * fs), (4 * fs, 2 * fs), (2 * fs, fs), (fs, fs))):
        res(2 * co, co, dims[4 - i])
    L.append((fs, classes, k1, dims[0]))
    return L


def conv_flops(L, B):
    return sum(2.0 * B * d[0] * d[1] * d[2] * ci * co * k[0] * k[1] * k[2] for ci, co, k, d in L)


def config_of(args, wl, world, peak_mem=None):
    """The workload description both arms print (the reference arm times a bounded sample of exactly this workload;
    what the sample was is stated in its `cpu_baseline.sample`)."""
    B, D, H, W = wl[4]
    cfg = {"workload": args.workload, "model": model_name(wl), "per_gpu_batch": B, "volume": [D, H, W], "classes": wl[2],
           "parallelism": "dp%d" % world, "amp": "autocast fp16 + GradScaler", "optimizer": "AdamW(fused)+EMA"}
    if peak_mem is not None:
        cfg["l2"] = "working set (%.1f GiB activations/step) >> 126 MB L2, no explicit flush" % peak_mem
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md).  The poller is started
    ahead of time (nvidia-smi needs ~0.2 s before its first line) and every line is stamped on arrival; only lines that
    arrived inside a `window()` — the device-resident and the end-to-end timed loops — are summarised."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.windows = [], None, index, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self):
        sampler = self

        class _W:
            def __enter__(self_w):
                self_w.t0 = time.time()

            def __exit__(self_w, *a):
                sampler.windows.append((self_w.t0, time.time()))
        return _W()

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for stamp, r in self.rows:
            # a line printed at time t describes the GPU a few ms earlier: accept up to 60 ms past the window's end
            if self.windows and not any(a <= stamp <= b + 0.06 for a, b in self.windows):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class quiet_stdout:
    """Temporarily point the process-level stdout (fd 1) at /dev/null — for native libraries that print banners."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def usable_cores():
    """Host threads this process can really use: CPU affinity capped by the cgroup CPU quota (a container that
    sees 128 logical CPUs but owns a 16-CPU quota thrashes with 128 threads)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return model, usable_cores()


# --------------------------------------------------------------------------- the UNMODIFIED reference, when it travelled
_REF = None


def reference_classes():
    """The reference's own classes from baseline/_ref (baseline/install_ref.py copies the files there byte for byte in
    the build container; git-ignored, travels to the GPU box).  None when absent -> the arms use the reference-pinned
    oracle port instead and say so (`kind: port`)."""
    global _REF
    if _REF is not None:
        return _REF or None
    root = os.path.join(ROOT, "baseline", "_ref")
    _REF = False
    if os.path.exists(os.path.join(root, "model", "dim3", "unet.py")):
        import types
        try:
            sys.path.insert(0, root)
            for pkg, sub in (("model", "model"), ("model.dim3", "model/dim3"), ("training", "training")):
                m = types.ModuleType(pkg)
                m.__path__ = [os.path.join(root, sub)]
                sys.modules[pkg] = m
            from model.dim3.unet import UNet
            from model.dim3.medformer import MedFormer
            from training.losses import DiceLoss
            _REF = {"UNet": UNet, "MedFormer": MedFormer, "DiceLoss": DiceLoss}
        except Exception as e:      # noqa
            sys.stderr.write("reference classes not importable from baseline/_ref: %r\n" % (e,))
            _REF = False
    return _REF or None


def reference_net(wl):
    """(net, loss_fn) built from the reference's classes exactly as model/utils.py:80-95 and train_ddp.py:93-94,186-191
    do, loaded with the seeded synthetic weights every arm uses; None for workloads whose reference class cannot be
    imported here (SwinUNETR needs monai)."""
    import torch
    import torch.nn as nn
    ref = reference_classes()
    if ref is None or is_swin(wl):
        return None
    if is_medformer(wl):
        net = ref["MedFormer"](1, wl[2], BASE, conv_block="BasicBlock", expansion=4, attn_drop=0, proj_drop=0,
                               proj_type="depthwise", norm="in", act="relu", **wl[0])
    else:
        net = ref["UNet"](1, BASE, scale=wl[0], kernel_size=wl[1], num_classes=wl[2], block="BasicBlock", norm="in")
    missing = net.load_state_dict(oracle_state(wl), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    dl = ref["DiceLoss"]()

    def loss_fn(result, label, ce):
        if isinstance(result, (tuple, list)):                        # deep supervision, train_ddp.py:186-189
            return sum(AUX_WEIGHT[j] * (ce(result[j], label.squeeze(1)) + dl(result[j], label)) for j in range(len(result)))
        return ce(result, label.squeeze(1)) + dl(result, label)
    return net, loss_fn


# --------------------------------------------------------------------------- CPU (reference) arm
def oracle_step_fn(wl):
    import torch
    rn = reference_net(wl)
    if rn is not None:                  # the unmodified reference modules on the host cores
        import torch.nn as nn
        net, loss_fn = rn
        net.train()
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5)
        ce = nn.CrossEntropyLoss(weight=torch.tensor(wl[3], dtype=torch.float32))

        def ref_step(img, lab):
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(net(img), lab, ce)
            loss.backward()
            opt.step()
            return loss.item()
        ref_step.kind = "reference"
        return ref_step
    return _oracle_port_step_fn(wl)


def _oracle_port_step_fn(wl):
    import torch
    sd = {k: v.requires_grad_(True) for k, v in oracle_state(wl).items()}
    params = list(sd.values())
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5)
    w = torch.tensor(wl[3], dtype=torch.float32)

    def step(img, lab):
        opt.zero_grad(set_to_none=True)
        loss = oracle_loss(wl, sd, img, lab, w)
        loss.backward()
        opt.step()
        return loss.item()
    step.kind = "port"
    return step


def cpu_sample_shape(wl, step):
    """Smallest legal depth-crop of the workload and the measured CPU rate on it (voxels/s)."""
    from oracle.synth import make_volume
    classes, (_, _, H, W) = wl[2], wl[4]
    if is_swin(wl):
        dmin = 64                      # img_size % 32 == 0 and >= 2 voxels at the deepest stage
    else:
        scale = wl[0]["scale"] if is_medformer(wl) else wl[0]
        dmin = 1
        for s in scale:
            dmin *= s[0]
        dmin = max(dmin * 2, 8)
    img, lab = make_volume(1, dmin, H, W, classes, seed=1)
    step(img, lab)
    t0 = time.time(); step(img, lab); t = time.time() - t0
    return dmin, dmin * H * W / t


def run_reference(args, wl):
    import torch
    from oracle.synth import make_volume
    classes, (B, D, H, W) = wl[2], wl[4]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    torch.set_num_threads(cores)
    step = oracle_step_fn(wl)
    dmin, rate = cpu_sample_shape(wl, step)
    total = args.steps + args.warmup
    budget = 150.0
    d = dmin
    while d * 2 <= D and (d * 2) * H * W * total / rate <= budget:
        d *= 2
    img, lab = make_volume(1, d, H, W, classes, seed=2023)
    for _ in range(args.warmup):
        step(img, lab)
    t0 = time.time()
    for _ in range(args.steps):
        step(img, lab)
    dt = time.time() - t0
    vps = args.steps * d * H * W / dt
    model, _ = host_info()
    how = ("the UNMODIFIED reference modules (baseline/_ref: model/dim3/*.py, training/losses.py, copied byte for byte by "
           "baseline/install_ref.py)" if step.kind == "reference" else
           "the reference-pinned oracle port of the reference modules (baseline/_ref absent or not importable for this "
           "workload; oracle/make_golden*.py prove the port equal to the reference)")
    sample = ("every step = one full train step (fwd + CE/Dice + bwd + AdamW) on a %dx%dx%d depth-crop of the %dx%dx%d volume, "
              "batch 1, fp32, %d host threads, %s" % (d, H, W, D, H, W, cores, how))
    out = {"impl": "reference", "metric": metric_of(wl), "value": vps, "unit": "voxels/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_of(args, wl, max(1, args.gpus)),
           "cpu_baseline": {"value": vps, "unit": "voxels/s", "cores": cores, "kind": step.kind, "sample": sample,
                            "cpu": model},
           "e2e": {"value": vps, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- B200 arm
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    import b200seg
    from b200seg import _lib, ops
    from b200seg.train import TrainStep
    from oracle.synth import make_volume

    scale, kernel, classes, weight, (B, D, H, W) = wl
    med = is_medformer(wl)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the ONE JSON line: communicator creation prints an "NCCL version ..." banner on fd 1
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/b200seg_nccl.%h.%p.log")
        with quiet_stdout():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    assert _lib.load().b200seg_check_device() == 0, "not an sm_100 device"

    def make_net():
        if is_swin(wl):
            n = b200seg.SwinUNETR((D, H, W), 1, classes, feature_size=wl[1])
            n.load_state_dict(oracle_state(wl), strict=False)            # relative_position_index buffers are derived
            return n.to(dev)
        if med:
            n = b200seg.MedFormer(1, classes, BASE, conv_block="BasicBlock", expansion=4, attn_drop=0, proj_drop=0,
                                  proj_type="depthwise", norm="in", act="relu", **wl[0])
        else:
            n = b200seg.UNet(1, BASE, scale=scale, kernel_size=kernel, num_classes=classes, block="BasicBlock", norm="in")
        n.load_state_dict(oracle_state(wl))
        return n.to(dev)
    net, ema = make_net(), make_net()
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        # train_ddp.py:353 wraps with find_unused_parameters=True; every parameter of these models receives a gradient
        # (SURVEY.md §8a "No unused parameters", tests/test_ddp_gloo.py), so the extra autograd-graph walk is dropped
        # here, and gradients live directly in the all-reduce buckets (no copy into them).  tests/test_gpu_ddp_nccl.py
        # keeps the reference's own flags.
        net = DDP(net, device_ids=[local], find_unused_parameters=False, gradient_as_bucket_view=True)
        if args.grad_compress:
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            net.register_comm_hook(None, default_hooks.fp16_compress_hook)      # half the all-reduce bytes
    for p in ema.parameters():
        p.requires_grad_(False)                                             # train_ddp.py:360-361
    ts = TrainStep(net, ema, ce_weight=torch.tensor(weight), amp=True, aux_weight=AUX_WEIGHT if med else None)

    img_h, lab_h = make_volume(B, D, H, W, classes, seed=2023 + rank)
    img_h, lab_h = img_h.pin_memory(), lab_h.pin_memory()
    img, lab = img_h.to(dev), lab_h.to(dev)
    vox = B * D * H * W

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def e2e_step():
        i = img_h.to(dev, non_blocking=True)                 # train_ddp.py:173-174
        l = lab_h.to(dev, non_blocking=True)
        return ts(i, l).item()                               # train_ddp.py:213

    with ClockSampler(local) as cs:                          # polling starts during the warm-up; only the two timed loops count
        for _ in range(max(args.warmup, 3)):
            ts(img, lab)
        torch.cuda.synchronize()
        c0 = _lib.launch_count
        with cs.window():
            ms = timed(lambda: ts(img, lab), args.steps)
        launches = (_lib.launch_count - c0) // args.steps
        e2e_step()
        with cs.window():
            ms_e2e = timed(e2e_step, args.steps)
    clocks = cs.summary()
    peak_mem = torch.cuda.max_memory_allocated() / 2**30

    out = {"metric": metric_of(wl), "value": world * vox * args.steps / (ms / 1e3), "unit": "voxels/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
           "config": config_of(args, wl, world, peak_mem),
           "clocks": clocks,
           "e2e": {"value": world * vox * args.steps / (ms_e2e / 1e3), "unit": "voxels/s",
                   "h2d_bytes_per_step": img_h.numel() * 4 + lab_h.numel() * 8, "d2h_bytes_per_step": 4,
                   "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": launches, "peak_mem_gib": peak_mem}

    if rank == 0 and med:
        out["config"]["algorithmic_gflop_per_step"] = 2128.0      # 3*709.5 - 0.5 (conv + bmm + mm)
        out["achieved_tflops_step"] = 2128.0e9 / (ms / args.steps / 1e3) / 1e12
        out["roofline"] = biattn_roofline(torch, ops, _lib, dev, B)
    if rank == 0 and not med:
        # ---- roofline of the dominant kernel: conv3d forward on the costliest layer shape
        if is_swin(wl):
            L = swin_conv_layers(wl[1], classes, D, H, W)
            fl_step = 4544.0e9          # fwd 1516.6 (conv 1392.5 + bmm 61.4 + addmm 58.3 + mm 4.5) + bwd 3027.4 GFLOP
        else:
            L = conv_layers(scale, kernel, classes, B, D, H, W)
            fl_step = 3 * conv_flops(L, B) - conv_flops(L[:1], B)
        out["config"]["algorithmic_gflop_per_step"] = fl_step / 1e9
        out["achieved_tflops_step"] = fl_step / (ms / args.steps / 1e3) / 1e12
        groups = {}
        for ci, co, k, d in L:
            groups[(ci, co, tuple(k), d)] = groups.get((ci, co, tuple(k), d), 0) + conv_flops([(ci, co, k, d)], B)
        (ci, co, k, d), _ = max(groups.items(), key=lambda kv: kv[1])
        x = torch.randn(B, d[0], d[1], d[2], ci, device=dev).half()
        st = ops.instnorm_stats(x, 0, ci)
        algo = ops.conv_algo(ci, co, k, torch.float16, B)
        wp = (ops.pack_weight(torch.randn(co, ci, *k, device=dev) * 0.05, torch.float16, layout=algo), algo)
        # launch-overhead-free timing: preallocated outputs, the C entry point called back to back
        y = torch.empty(B, d[0], d[1], d[2], co, device=dev, dtype=torch.float16)
        yst = torch.zeros(B, co, 2, device=dev, dtype=torch.float64)
        fn = _lib.load().b200seg_conv3d_fwd
        cargs = (x.data_ptr(), ci, 0, st.data_ptr(), 1e-4, 1, wp[0].data_ptr(), None, None, 0, 0, y.data_ptr(), co, 0,
                 yst.data_ptr(), None, 0, 0, None, 1e-4, 0, B, d[0], d[1], d[2], ci, co, k[0], k[1], k[2], 1, algo,
                 torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            assert fn(*cargs) == 0
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn(*cargs)
        e1.record()
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / reps
        kfl = conv_flops([(ci, co, list(k), d)], B)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops", 1590.0)
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        # (profiles/r1_ncu_full_conv_wgrad_128.txt: 68.09 MB read + 5.81 MB written per launch, incl. a residual read)
        traffic = 75.7e6 if (ci, co, tuple(k), tuple(d)) == (128, 128, (3, 3, 3), (128, 32, 32)) else None
        traffic_src = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at this shape (round 2: 68.1 + 7.6 MB; "
                       "the 126 MB L2 still holds most of the 33.5 MB output at kernel end), profiles/r2_ncu_full_128_128_k333_fwd.txt; not re-measured in-run")
        out["roofline"] = {"bound": "tensor", "achieved": kfl / (kms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                           "frac": kfl / (kms / 1e3) / 1e12 / peak, "traffic": traffic, "traffic_source": traffic_src if traffic else None,
                           "peak_source": "measured (MEASURED_PEAKS.json, burst)" if peaks else "fallback 1.59 PF",
                           "kernel": "conv3d_fwd[%s] %d->%d k%s @%s (IN+ReLU loader, IN-sums epilogue)" % ("tcgen05" if algo == 2 else "direct", ci, co, list(k), list(d)),
                           "ms": kms, "step_frac_of_sustained_peak": out["achieved_tflops_step"] / peaks.get("bf16_tflops_sustained", 1400.0)}
        del x, wp
    if rank == 0:
        # ---- CPU baseline: oracle port on the host cores, bounded sample
        if world == 1 and not args.no_cpu:
            cores = usable_cores()
            torch.set_num_threads(cores)
            step = oracle_step_fn(wl)
            dmin, rate = cpu_sample_shape(wl, step)
            dd = dmin
            while dd * 2 <= D and (dd * 2) * H * W * 2 / rate <= 25.0:
                dd *= 2
            ci_, cl_ = make_volume(1, dd, H, W, classes, seed=3)
            step(ci_, cl_)
            t0 = time.time(); step(ci_, cl_); dt = time.time() - t0
            out["cpu_baseline"] = {"value": dd * H * W / dt, "unit": "voxels/s", "cores": cores, "kind": step.kind,
                                   "cpu": host_info()[0],
                                   "sample": "1 timed step (after 1 warm-up) on a %dx%dx%d depth-crop, batch 1, fp32, all host threads, %s"
                                             % (dd, H, W, "unmodified reference modules (baseline/_ref)" if step.kind == "reference"
                                                else "reference-pinned oracle port")}
        # ---- the bar to beat: the same algorithm through stock PyTorch + cuDNN on this GPU (AMP)
        if world == 1 and not args.no_cudnn:
            try:
                out["torch_cudnn_same_gpu"] = cudnn_baseline(wl, img, lab, args.steps)
            except Exception as e:       # noqa
                out["torch_cudnn_same_gpu"] = {"error": str(e)[:200]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def biattn_roofline(torch, ops, _lib, dev, B):
    """HBM roofline of the B-MHA forward at MedFormer's largest attention level (down2/up2: N = 96*24*24 voxels,
    4 heads x 32): algorithmic bytes = read Q_f,V_f + write O_f = 3*B*N*inner*2."""
    heads, N = 4, 96 * 24 * 24
    inner = 32 * heads
    f = torch.randn(B, 96, 24, 24, 2 * inner, device=dev).half()
    m = torch.randn(B, 3, 3, 3, 2 * inner, device=dev).half()
    for _ in range(3):
        ops.biattn_fwd(f, m, heads)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.biattn_fwd(f, m, heads)
    e1.record()
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / reps
    nbytes = 3.0 * B * N * inner * 2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 7700.0)
    ach = nbytes / (kms / 1e3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of biattn_fwd_kernel at this shape from the committed ncu --set full
    # capture (profiles/r1_ncu_full_medformer_ops.txt: 31.22 MB read + 2.84 MB written per launch, B = 1)
    traffic = 34.06e6 if B == 1 else None
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 7.7 TB/s",
            "kernel": "biattn_fwd N=%d heads=%d dim_head=32 M=27 (B-MHA core, both softmax directions)" % (N, heads),
            "ms": kms}


def cudnn_baseline(wl, img, lab, steps):
    """The bar to beat: the reference's algorithm (the reference-pinned oracle modules: F.conv3d / F.instance_norm /
    F.relu / max_pool3d / interpolate, the reference's own losses restated) through stock PyTorch + cuDNN on the SAME
    GPU under autocast fp16, same step (AMP backward, fused AdamW, EMA).  Three library configurations are timed and the
    BEST is reported as `value`: default flags, cudnn.benchmark=True, and benchmark=True + channels_last_3d weights /
    activations (the layout cuDNN's fp16 tensor-core kernels prefer)."""
    import torch
    dev = img.device
    w = torch.tensor(wl[3], device=dev)

    use_ref = reference_net(wl) is not None

    def run_ref(benchmark, channels_last):
        """the UNMODIFIED reference modules (baseline/_ref) driven exactly like train_ddp.py:171-215"""
        import torch.nn as nn
        old = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = benchmark
        try:
            net, loss_fn = reference_net(wl)
            net = net.to(dev).train()
            if channels_last:
                net = net.to(memory_format=torch.channels_last_3d)
            x = img.contiguous(memory_format=torch.channels_last_3d) if channels_last else img
            params = list(net.parameters())
            ema = [v.detach().clone() for v in params]
            ce = nn.CrossEntropyLoss(weight=w)
            opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5, fused=True)
            scaler = torch.amp.GradScaler("cuda")

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.float16):
                    loss = loss_fn(net(x), lab, ce)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                torch._foreach_mul_(ema, 0.99)
                torch._foreach_add_(ema, [v.detach() for v in params], alpha=0.01)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps
        finally:
            torch.backends.cudnn.benchmark = old
            del net
            torch.cuda.empty_cache()

    def run(benchmark, channels_last):
        if use_ref:
            return run_ref(benchmark, channels_last)
        old = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = benchmark
        try:
            sd = {}
            for k, v in oracle_state(wl).items():
                v = v.to(dev)
                if channels_last and v.dim() == 5:
                    v = v.contiguous(memory_format=torch.channels_last_3d)
                sd[k] = v.requires_grad_(True)
            x = img.contiguous(memory_format=torch.channels_last_3d) if channels_last else img
            ema = [v.detach().clone() for v in sd.values()]
            opt = torch.optim.AdamW(list(sd.values()), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5, fused=True)
            scaler = torch.amp.GradScaler("cuda")

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.float16):
                    loss = oracle_loss(wl, sd, x, lab, w)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                torch._foreach_mul_(ema, 0.99)
                torch._foreach_add_(ema, [v.detach() for v in sd.values()], alpha=0.01)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps
        finally:
            torch.backends.cudnn.benchmark = old
            torch.cuda.empty_cache()
    variants = {}
    for name, (bm, cl) in (("default", (False, False)), ("cudnn_benchmark", (True, False)),
                           ("cudnn_benchmark+channels_last_3d", (True, True))):
        try:
            variants[name] = run(bm, cl)
        except Exception as e:       # noqa
            variants[name] = "error: " + str(e)[:120]
    ok = {k: v for k, v in variants.items() if isinstance(v, float)}
    best = min(ok, key=ok.get)
    ms = ok[best]
    return {"value": img.shape[0] * img[0, 0].numel() / (ms / 1e3), "unit": "voxels/s", "ms_per_step": ms, "best_variant": best,
            "ms_per_step_by_variant": variants,
            "modules": "unmodified reference (baseline/_ref)" if use_ref else "reference-pinned oracle port",
            "what": "%s via stock torch %s + cuDNN %s, autocast fp16 + GradScaler + fused AdamW + EMA, same GPU, same step; best "
                    "of three library configurations" % ("the UNMODIFIED reference modules (baseline/_ref, train_ddp.py:171-215 loop)"
                                                         if use_ref else "reference algorithm (reference-pinned oracle modules)",
                                                         torch.__version__, torch.backends.cudnn.version())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="resunet_acdc_128", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cudnn", action="store_true")
    ap.add_argument("--grad-compress", action="store_true", help="fp16 gradient all-reduce (DDP comm hook); default fp32 like the reference")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
