"""Timeline of block 0's warp roles for one conv layer (B200SEG_TC_DEBUG=1).  python tools/tc_trace.py [layer-index]"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200SEG_TC_DEBUG", "1")
from b200seg import ops, _lib  # noqa
LAYERS = [(32, 32, (1, 3, 3), (128, 128, 128)), (128, 128, (3, 3, 3), (128, 32, 32))]
lib = _lib.load()
buf = (ctypes.c_longlong * 1024)()
for ci, co, k, (D, H, W) in LAYERS:
    x = torch.randn(1, D, H, W, ci, device="cuda").half()
    r = torch.randn(1, D, H, W, co, device="cuda").half()
    st = ops.instnorm_stats(x, 0, ci)
    algo = ops.conv_algo(ci, co, k, torch.float16, 1)
    wp = (ops.pack_weight(torch.randn(co, ci, *k, device="cuda") * 0.05, torch.float16, layout=algo), algo)
    for _ in range(2):
        ops.conv3d_fwd(x, 0, ci, st, ops.ACT_RELU, wp, co, k, residual=r)
    lib.b200seg_debug_tc_trace(buf)
    t = list(buf)
    base = min(v for v in t if v > 0)
    print("layer %d->%d k%s: events (cycles since first stamp)" % (ci, co, k))
    names = {0: ("mma", ["start", "A_ready", "issued", "committed"]), 1: ("load0", ["start", "slot_free", "stored", "published"]),
             2: ("load1", ["start", "slot_free", "stored", "published"]), 3: ("epi", ["start", "acc_ready", "drained"])}
    for role in range(4):
        nm, ev = names[role]
        row = t[role * 256:(role + 1) * 256]
        n = len(ev)
        for i in range(0, 8 * n, n):
            if row[i] == 0:
                break
            print("  %-6s #%d: " % (nm, i // n) + "  ".join("%s %d" % (ev[j], row[i + j] - base) for j in range(n) if row[i + j] > 0))
