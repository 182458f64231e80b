"""CPU: lane-level emulation of csrc/swin_mma.cu's register algebra (TEST INFRASTRUCTURE).

The tensor-core window-attention kernels keep scores / probabilities in mma.sync m16n8k16 accumulator registers and
re-pack them as the A operand of the next MMA; a wrong fragment index there gives plausible-looking garbage.  This
file restates the PTX fragment layouts (A row-major 16x16, B col-major 16x8, C 16x8 — PTX ISA "Matrix Fragments for
mma.m16n8k16") as a 32-lane emulator and runs the kernels' exact indexing (same shared-memory layouts, same loops,
same re-packing, same bias index L_i - L_j + K0) for one (window, head) against a direct numpy attention, forward
and both backward passes.  It cannot prove the hardware layout, but it proves the kernel's index algebra is
self-consistent with the documented one; the GPU tests then compare with the reference fixtures."""
import numpy as np
import pytest

F16 = np.float16


def mma16816(d, a, b0, b1):
    """d[lane][4] += A(16x16) @ B(16x8); a[lane][4][2], b0/b1[lane][2] hold the halves each lane owns."""
    A = np.zeros((16, 16), np.float32)
    B = np.zeros((16, 8), np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2] = a[lane][0]
        A[g + 8, 2 * t:2 * t + 2] = a[lane][1]
        A[g, 2 * t + 8:2 * t + 10] = a[lane][2]
        A[g + 8, 2 * t + 8:2 * t + 10] = a[lane][3]
        B[2 * t:2 * t + 2, g] = b0[lane]
        B[2 * t + 8:2 * t + 10, g] = b1[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        d[lane][0] += D[g, 2 * t]; d[lane][1] += D[g, 2 * t + 1]
        d[lane][2] += D[g + 8, 2 * t]; d[lane][3] += D[g + 8, 2 * t + 1]


class Layout:
    def __init__(self, n, DH):
        self.n, self.DH = n, DH
        self.DK = max(16, DH); self.RS = self.DK + 8; self.KS = self.DK // 16; self.NT = DH // 8
        self.NP = (n + 15) // 16 * 16; self.VS = self.NP + 8

    def rowmajor(self, x):                 # [n][DH] -> flat [NP][RS] halves, zero padded
        a = np.zeros((self.NP, self.RS), F16)
        a[:self.n, :self.DH] = x.astype(F16)
        return a.reshape(-1)

    def transposed(self, x):               # [n][DH] -> flat [DH][VS]
        a = np.zeros((self.DH, self.VS), F16)
        a[:, :self.n] = x.astype(F16).T
        return a.reshape(-1)

    def load_a(self, arr, r0):
        a = [[None] * 4 for _ in range(self.KS)]
        out = []
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            per = []
            for ks in range(self.KS):
                o = ks * 16 + 2 * t
                per.append([arr[(r0 + g) * self.RS + o:(r0 + g) * self.RS + o + 2], arr[(r0 + g + 8) * self.RS + o:(r0 + g + 8) * self.RS + o + 2],
                            arr[(r0 + g) * self.RS + o + 8:(r0 + g) * self.RS + o + 10],
                            arr[(r0 + g + 8) * self.RS + o + 8:(r0 + g + 8) * self.RS + o + 10]])
            out.append(per)
        return out                          # [lane][ks][4][2]

    def mma_rowmajor(self, acc, afr, arr, c0):
        for ks in range(self.KS):
            b0 = [arr[(c0 + (l >> 2)) * self.RS + ks * 16 + 2 * (l & 3):][:2] for l in range(32)]
            b1 = [arr[(c0 + (l >> 2)) * self.RS + ks * 16 + 2 * (l & 3) + 8:][:2] for l in range(32)]
            mma16816(acc, [afr[l][ks] for l in range(32)], b0, b1)

    def mma_transposed(self, acc, p, tr, k0):
        for nt in range(self.NT):
            b0 = [tr[(nt * 8 + (l >> 2)) * self.VS + k0 + 2 * (l & 3):][:2] for l in range(32)]
            b1 = [tr[(nt * 8 + (l >> 2)) * self.VS + k0 + 2 * (l & 3) + 8:][:2] for l in range(32)]
            mma16816([acc[l][nt] for l in range(32)], p, b0, b1)


def pack(sc):
    """accumulator registers sc[h][4] of a 16x16 block -> A fragment [4][2] (the kernel's `p` array)"""
    f = lambda x, y: np.array([x, y], np.float32).astype(F16)      # noqa: E731
    return [f(sc[0][0], sc[0][1]), f(sc[0][2], sc[0][3]), f(sc[1][0], sc[1][1]), f(sc[1][2], sc[1][3])]


def setup(n, DH, masked, seed):
    rng = np.random.RandomState(seed)
    f = (3, 4, 2) if n <= 24 else (4, 4, 4)
    q, k, v, dO = [rng.randn(n, DH).astype(np.float32).astype(F16).astype(np.float32) for _ in range(4)]
    S2 = 2 * f[2] - 1; S1 = (2 * f[1] - 1) * S2; TBL = (2 * f[0] - 1) * S1
    K0 = (f[0] - 1) * S1 + (f[1] - 1) * S2 + (f[2] - 1)
    coords = [(t // (f[2] * f[1]), (t // f[2]) % f[1], t % f[2]) for t in range(n)]
    L = np.array([a * S1 + b * S2 + c for a, b, c in coords])
    table = rng.randn(TBL).astype(np.float32)
    rid = rng.randint(0, 3, n) if masked else np.zeros(n, int)
    valid = rng.rand(n) > 0.15
    valid[0] = True
    scale = DH ** -0.5
    # direct evaluation (fp32 on the fp16-rounded operands)
    qs = (q * scale).astype(F16).astype(np.float32)
    idx = L[:, None] - L[None, :] + K0
    # the closed form equals the reference's 3-D offset index (relative_position_index, swin_unetr.py:417-459)
    for i in (0, n // 2, n - 1):
        for j in (0, n // 3, n - 1):
            (a1, b1, c1), (a2, b2, c2) = coords[i], coords[j]
            assert idx[i, j] == ((a1 - a2 + f[0] - 1) * (2 * f[1] - 1) + (b1 - b2 + f[1] - 1)) * (2 * f[2] - 1) + (c1 - c2 + f[2] - 1)
    s = qs @ k.T + table[idx] + np.where(rid[:, None] != rid[None, :], -100.0, 0.0)
    m = s.max(1, keepdims=True)
    p = np.exp(s - m); l = p.sum(1, keepdims=True)
    P = p / l
    O = P @ v
    lse = (m + np.log(l))[:, 0]
    dOv = dO * valid[:, None]
    delta = (dOv * O).sum(1)
    dP = dOv @ v.T
    dS = P * (dP - delta[:, None]) * valid[:, None]
    ref = dict(O=O, lse=lse, dQ=(dS @ k) * scale, dK=dS.T @ qs, dV=(P * valid[:, None]).T @ dOv, dT=np.bincount(idx.reshape(-1), dS.reshape(-1), TBL),
               delta=delta)
    return dict(q=q, k=k, v=v, dO=dOv, qs=qs, L=L, rid=rid, table=table, K0=K0, valid=valid, scale=scale, masked=masked, TBL=TBL), ref


@pytest.mark.parametrize("n,DH,masked", [(24, 8, True), (40, 16, False), (27, 32, True)])
def test_fragment_algebra(n, DH, masked):
    x, ref = setup(n, DH, masked, seed=n + DH)
    lo = Layout(n, DH)
    NP = lo.NP
    Lp = np.zeros(NP, int); Lp[:n] = x["L"]
    ridp = np.zeros(NP, int); ridp[:n] = x["rid"]
    vox = -np.ones(NP, int); vox[:n] = np.where(x["valid"], 1, -1)
    sQ, sK, sV, sDO = lo.rowmajor(x["qs"]), lo.rowmajor(x["k"]), lo.rowmajor(x["v"]), lo.rowmajor(x["dO"])
    sVt, sKt, sQt, sDOt = lo.transposed(x["v"]), lo.transposed(x["k"]), lo.transposed(x["qs"]), lo.transposed(x["dO"])
    table, K0 = x["table"], x["K0"]
    lse_p = np.zeros(NP, np.float32); lse_p[:n] = ref["lse"]
    dl_p = np.zeros(NP, np.float32); dl_p[:n] = ref["delta"]

    # ------------------------------------------------ forward (win_attn_fwd_mma_kernel)
    O = np.zeros((NP, DH), np.float32); lse = np.zeros(NP, np.float32)
    for r0 in range(0, NP, 16):
        aq = lo.load_a(sQ, r0)
        m = np.full((32, 2), -np.inf, np.float32); l = np.zeros((32, 2), np.float32)
        o = [[[0.0] * 4 for _ in range(lo.NT)] for _ in range(32)]
        for kb in range(0, NP, 16):
            sc = [[[0.0] * 4 for _ in range(2)] for _ in range(32)]
            for h in range(2):
                lo.mma_rowmajor([sc[la][h] for la in range(32)], aq, sK, kb + h * 8)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    for e in range(2):
                        j = kb + h * 8 + 2 * t + e
                        for half, row in ((0, r0 + g), (1, r0 + g + 8)):
                            val = sc[lane][h][2 * half + e] + table[Lp[row] + K0 - Lp[j]]
                            if masked and ridp[j] != ridp[row]:
                                val -= 100.0
                            if j >= n:
                                val = -np.inf
                            sc[lane][h][2 * half + e] = val
            # quad max, online softmax
            for q4 in range(8):
                lanes = range(q4 * 4, q4 * 4 + 4)
                for half in range(2):
                    mx = max(sc[la][h][2 * half + e] for la in lanes for h in range(2) for e in range(2))
                    for la in lanes:
                        mn = max(m[la][half], mx)
                        corr = np.exp(m[la][half] - mn) if np.isfinite(m[la][half]) else 0.0
                        m[la][half] = mn
                        ssum = 0.0
                        for h in range(2):
                            for e in range(2):
                                sc[la][h][2 * half + e] = np.exp(sc[la][h][2 * half + e] - mn)
                                ssum += sc[la][h][2 * half + e]
                        l[la][half] = l[la][half] * corr + ssum
                        for nt in range(lo.NT):
                            o[la][nt][2 * half] *= corr; o[la][nt][2 * half + 1] *= corr
            lo.mma_transposed(o, [pack(sc[la]) for la in range(32)], sVt, kb)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for half, row in ((0, r0 + g), (1, r0 + g + 8)):
                lt = sum(l[la][half] for la in range((lane >> 2) * 4, (lane >> 2) * 4 + 4))
                for nt in range(lo.NT):
                    O[row, nt * 8 + 2 * t] = o[lane][nt][2 * half] / lt
                    O[row, nt * 8 + 2 * t + 1] = o[lane][nt][2 * half + 1] / lt
                lse[row] = m[lane][half] + np.log(lt)
    assert np.abs(O[:n] - ref["O"]).max() < 5e-3 * max(1.0, np.abs(ref["O"]).max())
    assert np.abs(lse[:n] - ref["lse"]).max() < 2e-3

    # ------------------------------------------------ backward pass A (win_attn_bwd_q_mma_kernel)
    dQ = np.zeros((NP, DH), np.float32); dT = np.zeros(x["TBL"], np.float64)
    for r0 in range(0, NP, 16):
        aq, ado = lo.load_a(sQ, r0), lo.load_a(sDO, r0)
        dq = [[[0.0] * 4 for _ in range(lo.NT)] for _ in range(32)]
        for kb in range(0, NP, 16):
            sc = [[[0.0] * 4 for _ in range(2)] for _ in range(32)]
            dp = [[[0.0] * 4 for _ in range(2)] for _ in range(32)]
            for h in range(2):
                lo.mma_rowmajor([sc[la][h] for la in range(32)], aq, sK, kb + h * 8)
                lo.mma_rowmajor([dp[la][h] for la in range(32)], ado, sV, kb + h * 8)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    for e in range(2):
                        j = kb + h * 8 + 2 * t + e
                        for half, row in ((0, r0 + g), (1, r0 + g + 8)):
                            ix = Lp[row] + K0 - Lp[j]
                            val = sc[lane][h][2 * half + e] + table[ix]
                            if masked and ridp[j] != ridp[row]:
                                val -= 100.0
                            live = j < n and vox[row] >= 0
                            ds = np.exp(val - lse_p[row]) * (dp[lane][h][2 * half + e] - dl_p[row]) if live else 0.0
                            if live:
                                dT[ix] += ds
                            sc[lane][h][2 * half + e] = ds
            lo.mma_transposed(dq, [pack(sc[la]) for la in range(32)], sKt, kb)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for half, row in ((0, r0 + g), (1, r0 + g + 8)):
                for nt in range(lo.NT):
                    dQ[row, nt * 8 + 2 * t] = dq[lane][nt][2 * half] * x["scale"]
                    dQ[row, nt * 8 + 2 * t + 1] = dq[lane][nt][2 * half + 1] * x["scale"]
    sc_ = max(1.0, np.abs(ref["dQ"]).max())
    assert np.abs(dQ[:n] - ref["dQ"]).max() < 1e-2 * sc_
    assert np.abs(dT - ref["dT"]).max() < 1e-2 * max(1.0, np.abs(ref["dT"]).max())

    # ------------------------------------------------ backward pass B (win_attn_bwd_kv_mma_kernel)
    dK = np.zeros((NP, DH), np.float32); dV = np.zeros((NP, DH), np.float32)
    for r0 in range(0, NP, 16):
        ak, av = lo.load_a(sK, r0), lo.load_a(sV, r0)
        dk = [[[0.0] * 4 for _ in range(lo.NT)] for _ in range(32)]
        dv = [[[0.0] * 4 for _ in range(lo.NT)] for _ in range(32)]
        for qb in range(0, NP, 16):
            sc = [[[0.0] * 4 for _ in range(2)] for _ in range(32)]
            dp = [[[0.0] * 4 for _ in range(2)] for _ in range(32)]
            pr = [[[0.0] * 4 for _ in range(2)] for _ in range(32)]
            for h in range(2):
                lo.mma_rowmajor([sc[la][h] for la in range(32)], ak, sQ, qb + h * 8)
                lo.mma_rowmajor([dp[la][h] for la in range(32)], av, sDO, qb + h * 8)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    for e in range(2):
                        i = qb + h * 8 + 2 * t + e
                        for half, row in ((0, r0 + g), (1, r0 + g + 8)):
                            val = sc[lane][h][2 * half + e] + table[Lp[i] - (Lp[row] - K0)]
                            if masked and ridp[i] != ridp[row]:
                                val -= 100.0
                            pv = np.exp(val - lse_p[i]) if vox[i] >= 0 else 0.0
                            pr[lane][h][2 * half + e] = pv
                            sc[lane][h][2 * half + e] = pv * (dp[lane][h][2 * half + e] - dl_p[i])
            lo.mma_transposed(dv, [pack(pr[la]) for la in range(32)], sDOt, qb)
            lo.mma_transposed(dk, [pack(sc[la]) for la in range(32)], sQt, qb)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for half, row in ((0, r0 + g), (1, r0 + g + 8)):
                for nt in range(lo.NT):
                    for e in range(2):
                        dK[row, nt * 8 + 2 * t + e] = dk[lane][nt][2 * half + e]
                        dV[row, nt * 8 + 2 * t + e] = dv[lane][nt][2 * half + e]
    assert np.abs(dK[:n] - ref["dK"]).max() < 1e-2 * max(1.0, np.abs(ref["dK"]).max())
    assert np.abs(dV[:n] - ref["dV"]).max() < 1e-2 * max(1.0, np.abs(ref["dV"]).max())
