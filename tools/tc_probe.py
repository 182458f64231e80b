"""Diagnostic probes for the tcgen05 conv kernel: each probe isolates one mechanism (K-steps, K-chunks,
w/h/d tap shifts, N tiling, ragged tiles, epilogue modes) and compares against the CUDA-core kernel.
Run one probe per process (a protocol bug traps and kills the CUDA context):
    python tools/tc_probe.py <probe-index>     |    python tools/tc_probe.py all   (spawns subprocesses)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PROBES = [
    # name, Cin, Cout, k, (B,D,H,W), mode
    ("1x1 K16 N16", 16, 16, (1, 1, 1), (1, 1, 16, 8), "plain"),
    ("1x1 K64 N64", 64, 64, (1, 1, 1), (1, 2, 16, 8), "plain"),
    ("1x1 K128(2 chunks) N32", 128, 32, (1, 1, 1), (1, 2, 32, 16), "plain"),
    ("1x1x3 w-shift", 32, 32, (1, 1, 3), (1, 2, 16, 16), "plain"),
    ("1x3x1 h-shift", 32, 32, (1, 3, 1), (1, 2, 32, 8), "plain"),
    ("3x1x1 d-stages", 32, 32, (3, 1, 1), (1, 4, 16, 8), "plain"),
    ("3x3x3 32->32", 32, 32, (3, 3, 3), (1, 4, 16, 16), "plain"),
    ("3x3x3 ragged 20x12", 32, 64, (3, 3, 3), (2, 3, 20, 12), "plain"),
    ("1x3x3 96->64 (KC48)", 96, 64, (1, 3, 3), (1, 2, 32, 16), "norm"),
    ("3x3x3 64->128 norm+res", 64, 128, (3, 3, 3), (1, 3, 16, 16), "normres"),
    ("3x3x3 192->64 (3 chunks)", 192, 64, (3, 3, 3), (1, 2, 16, 8), "norm"),
    ("3x3x3 64->320 (2 ntiles)", 64, 320, (3, 3, 3), (1, 2, 16, 8), "norm"),
    ("3x3x3 128->256 (single acc)", 128, 256, (3, 3, 3), (1, 2, 16, 8), "norm"),
    ("dgrad 64->32", 64, 32, (3, 3, 3), (2, 3, 16, 16), "dgrad"),
    ("many tiles 32->32 1x3x3", 32, 32, (1, 3, 3), (1, 8, 64, 64), "normres"),
]


WPROBES = [
    # name, Cin, Cout, k, (B,D,H,W), normalised
    ("wgrad 1x1 16->16", 16, 16, (1, 1, 1), (1, 1, 16, 8), False),
    ("wgrad 1x1 64->128", 64, 128, (1, 1, 1), (1, 2, 16, 16), False),
    ("wgrad 1x1x3 32->32", 32, 32, (1, 1, 3), (1, 2, 16, 16), False),
    ("wgrad 1x3x1 32->32", 32, 32, (1, 3, 1), (1, 2, 32, 8), False),
    ("wgrad 3x1x1 32->32", 32, 32, (3, 1, 1), (1, 4, 16, 8), False),
    ("wgrad 3x3x3 32->64 norm ragged", 32, 64, (3, 3, 3), (2, 3, 20, 12), True),
    ("wgrad 1x3x3 96->64 norm", 96, 64, (1, 3, 3), (1, 2, 32, 16), True),
    ("wgrad 3x3x3 128->128 norm", 128, 128, (3, 3, 3), (1, 4, 32, 32), True),
    ("wgrad 3x3x3 192->256", 192, 256, (3, 3, 3), (1, 2, 16, 16), True),
    ("wgrad 3x3x3 64->320", 64, 320, (3, 3, 3), (1, 2, 16, 8), True),
]


def run_wprobe(i):
    import torch
    from b200seg import ops, _lib
    name, Cin, Cout, k, (B, D, H, W), normed = WPROBES[i]
    torch.manual_seed(200 + i)
    x = torch.randn(B, D, H, W, Cin, device="cuda").half()
    dy = torch.randn(B, D, H, W, Cout, device="cuda").half()
    st = ops.instnorm_stats(x, 0, Cin) if normed else None
    act = ops.ACT_RELU if normed else ops.ACT_NONE
    out = {}
    for algo in (_lib.ALGO_DIRECT, _lib.ALGO_TC):
        dw, _ = ops.conv3d_wgrad(x, 0, Cin, st, act, dy, 0, Cout, k, algo=algo)
        torch.cuda.synchronize()
        out[algo] = dw.double().cpu()
    a, b = out[_lib.ALGO_TC], out[_lib.ALGO_DIRECT]
    err = ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
    ok = err < 2e-3
    print("wprobe %2d %-32s dw_rel_err %.3e %s" % (i, name, err, "OK" if ok else "FAIL"), flush=True)
    if not ok:
        e = (a - b).abs()
        print("   tc absmax %.4f direct absmax %.4f nonfinite=%d" % (a.abs().max(), b.abs().max(), (~torch.isfinite(a)).sum()))
        print("   err by tap :", [round(v, 2) for v in e.flatten(2).amax(dim=(0, 1)).tolist()])
        print("   err by co  :", [round(v, 2) for v in e.amax(dim=(1, 2, 3, 4)).tolist()][:40])
        print("   err by ci  :", [round(v, 2) for v in e.amax(dim=(0, 2, 3, 4)).tolist()][:40])
        print("   tc[0,:8,0,0,0]    :", [round(v, 2) for v in a[0, :8].flatten(1)[:, 0].tolist()])
        print("   direct[0,:8,...]  :", [round(v, 2) for v in b[0, :8].flatten(1)[:, 0].tolist()])
    return ok


def run_probe(i):
    import torch
    from b200seg import ops, _lib
    name, Cin, Cout, k, (B, D, H, W), mode = PROBES[i]
    torch.manual_seed(100 + i)
    dev = "cuda"
    x = torch.randn(B, D, H, W, Cin, device=dev).half()
    w = (torch.randn(Cout, Cin, *k, device=dev) * (1.0 / (Cin * k[0] * k[1] * k[2]) ** 0.5))
    st = ops.instnorm_stats(x, 0, Cin) if mode in ("norm", "normres") else None
    act = ops.ACT_RELU if st is not None else ops.ACT_NONE
    res = torch.randn(B, D, H, W, Cout, device=dev).half() if mode == "normres" else None
    dg = None
    if mode == "dgrad":
        gx = torch.randn(B, D, H, W, Cout, device=dev).half()
        dg = (gx, 0, ops.instnorm_stats(gx, 0, Cout), ops.ACT_RELU)
    out = {}
    for algo in (_lib.ALGO_DIRECT, _lib.ALGO_TC):
        wp = ops.pack_weight(w, torch.float16, layout=algo)
        y, ys = ops.conv3d_fwd(x, 0, Cin, st, act, wp, Cout, k, residual=res, dgrad_of=dg, algo=algo)
        torch.cuda.synchronize()
        out[algo] = (y.float().cpu(), ys.cpu())
    yd, sd = out[_lib.ALGO_DIRECT]
    yt, stc = out[_lib.ALGO_TC]
    err = ((yt - yd).abs().max() / (yd.abs().max() + 1e-30)).item()
    serr = ((stc - sd).abs().max() / (sd.abs().max() + 1e-30)).item()
    ok = err < 3e-3 and serr < 1e-3
    print("probe %2d %-32s y_rel_err %.3e  stats_rel_err %.3e  %s" % (i, name, err, serr, "OK" if ok else "FAIL"), flush=True)
    if not ok:
        d = (yt - yd).abs()
        idx = d.flatten().argmax().item()
        print("   worst at flat %d: tc %.5f direct %.5f | tc absmax %.4f direct absmax %.4f nonfinite=%d" %
              (idx, yt.flatten()[idx], yd.flatten()[idx], yt.abs().max(), yd.abs().max(), (~torch.isfinite(yt)).sum()))
        # per-row / per-channel error structure of the first (b=0,d=0) slice helps localise descriptor bugs
        e = (yt - yd)[0, 0].abs()
        print("   err by h-row :", [round(v, 3) for v in e.amax(dim=(1, 2)).tolist()][:20])
        print("   err by w-col :", [round(v, 3) for v in e.amax(dim=(0, 2)).tolist()][:20])
        print("   err by chan  :", [round(v, 3) for v in e.amax(dim=(0, 1)).tolist()][:32])
        print("   ratio tc/direct sample:", [round((yt.flatten()[j] / (yd.flatten()[j] + 1e-9)).item(), 3) for j in range(0, 64, 4)])
    return ok


if __name__ == "__main__":
    arg = sys.argv[1] if len(sys.argv) > 1 else "all"
    if arg == "wall":
        bad = 0
        for i in range(len(WPROBES)):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "w%d" % i], capture_output=True, text=True, timeout=180)
                sys.stdout.write(r.stdout)
                if r.returncode not in (0, 1):
                    bad += 1
                    print("wprobe %2d %-32s CRASHED rc=%d: %s" % (i, WPROBES[i][0], r.returncode, (r.stderr or "")[-600:].replace("\n", " | ")), flush=True)
                elif r.returncode == 1:
                    bad += 1
            except subprocess.TimeoutExpired:
                bad += 1
                print("wprobe %2d %-32s TIMEOUT" % (i, WPROBES[i][0]), flush=True)
        sys.exit(1 if bad else 0)
    if arg.startswith("w"):
        sys.exit(0 if run_wprobe(int(arg[1:])) else 1)
    if arg == "all":
        bad = 0
        for i in range(len(PROBES)):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True, timeout=180)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    bad += 1
                    print("probe %2d %-32s CRASHED rc=%d: %s" % (i, PROBES[i][0], r.returncode, (r.stderr or "")[-600:].replace("\n", " | ")), flush=True)
            except subprocess.TimeoutExpired:
                bad += 1
                print("probe %2d %-32s TIMEOUT" % (i, PROBES[i][0]), flush=True)
        sys.exit(1 if bad else 0)
    sys.exit(0 if run_probe(int(arg)) else 1)
