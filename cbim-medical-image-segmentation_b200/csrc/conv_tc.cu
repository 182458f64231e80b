// conv_tc.cu — conv3d forward / data-gradient as a tcgen05 implicit GEMM (sm_100a), fp16 operands,
// fp32 accumulation in TMEM.  Replaces cuDNN fprop/dgrad behind nn.Conv3d (conv_layers.py:29-38) and
// fuses the surrounding InstanceNorm+ReLU (conv_layers.py:40-43), residual add (:92) and the next
// layer's InstanceNorm reduction.
//
// GEMM view (per CTA tile):  D[128 voxels][NT cout] += A[128 voxels][KC cin] * B[KC cin][NT cout]
// for every filter tap and every KC-chunk of Cin.
//   * M tile  = 16(h) x 8(w) output voxels of one depth slice; GEMM row r = hl*8 + wl.
//   * A operand = a HALO tile (16+kh-1)x(8+kw-1) voxels x KC channels of ONE input depth slice, staged
//     once in shared memory as [KC/8][halo voxel][8 ch] (UMMA K-major, no-swizzle core matrices:
//     8 consecutive w-voxels x 16 B).  Every (kh,kw) tap reads the SAME staged tile through a shifted
//     matrix descriptor (start += (zh*HALO_W + zw)*16 B, SBO = HALO_W*16 B) — im2col is never formed.
//   * Staging (round 2): the loader warps no longer hold the tile in registers across the global-memory
//     latency.  They issue 16-byte cp.async copies (LDGSTS, zero-fill for padding / ragged tiles) for up to three
//     stages AHEAD, and when a stage has landed apply InstanceNorm-normalise + ReLU IN PLACE in shared memory
//     (skipped for raw inputs: every data-gradient launch), then publish it to the tensor core.  The normalised
//     activation tensor never exists in HBM.
//   * B operand = weights pre-packed on device into the exact shared-memory image per (ntile,tap,kchunk)
//     ([KC/8][NT][8] fp16): resident for the CTA's lifetime when the layer's weights fit (<=112 KB), else streamed
//     by 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) through an mbarrier ring.
//   * D lives in TMEM (double-buffered when 2*NT <= 512 columns); EIGHT epilogue warps (two per TMEM lane
//     quadrant, alternating 16-column chunks) drain it with tcgen05.ld while the MMA warp works on the next tile.
//     InstanceNorm sums of the stored tile: for NT <= 64 per-thread register accumulators carried across all the
//     tiles of the persistent CTA (2 FP ops per value instead of a 16-shuffle butterfly per chunk) and reduced
//     once at the end; for wider tiles the butterfly (the MMA time hides it there).
// Warp roles (640 threads = 5 warpgroups, 1 CTA/SM, persistent over tiles; `setmaxnreg` moves registers from the
// loader / MMA warpgroups to the two epilogue warpgroups, whose statistics accumulators need them):
//   warps 0-7  epilogue (TMEM lane quadrant == warp id & 3): bias / residual / dgrad ReLU-mask, fp16 store, IN sums
//   warps 8-15 A loaders (cp.async prefetch + in-place transform; all 256 threads share every stage)
//   warp  16   weight producer (bulk TMA)
//   warp  17   TMEM alloc + tcgen05.mma issue (highest working warp id = highest issue priority)
// Roofline: tensor pipe (dense fp16) for NT>=128; for NT<128 the MMA is bound by the shared-memory read of
// A (SS mode: 4 KB per 128xNTx16 instruction, ~52 cycles), see DESIGN.md.
#include "common.cuh"
#include "conv_args.h"
#include "tc_common.cuh"
#include "tmap.h"
#include <string.h>
#include <stdlib.h>

namespace {

using namespace tc;

constexpr int TH = 16, TW = 8;            // output tile (h, w); M = 128
constexpr int kEpiWarps = 8;
constexpr int kStatCopies = 4;            // per TMEM lane quadrant (the two warps of a quadrant own disjoint columns)
// The SM arbitrates highest-warp-id-first inside a sub-partition (B300_MICROARCH.md): the single MMA-issuing warp
// must never queue behind ALU-heavy loader / epilogue warps, so it gets the highest id (measured: 4-5x faster issue).
constexpr int kLoadWarp0 = 8;
constexpr int kLoadThreads = 256;
constexpr int kWgtWarp = 16;
constexpr int kMmaWarp = 17;
constexpr int kThreads = 20 * 32;   // 640: five complete warpgroups (setmaxnreg is a warpgroup-wide instruction);
                                    // warps 18-19 only take part in the block-wide barriers
// Registers per thread after the role dispatch.  setmaxnreg trades registers inside the CTA's OWN pool — what the
// launch allotted: 640 threads x 96 = 61440, not the SM's 65536 (an over-subscribed split makes setmaxnreg.inc spin
// forever: measured the hard way).  Split: 2 epilogue warpgroups x 136 (64 statistics accumulators at NT = 64) +
// 2 loader warpgroups x 80 (three transform chains interleaved: with 64 the compiler serialised them, -15..25 % on the
// Cin <= 64 forward layers) + {weights, MMA, 2 idle} x 40 (their loops live in uniform registers) = 60416.
constexpr int kRegsLaunch = 96, kRegsEpi = 136, kRegsLoad = 80, kRegsMma = 40;
static_assert(2 * 128 * kRegsEpi + 2 * 128 * kRegsLoad + 128 * kRegsMma <= kThreads * kRegsLaunch, "register split exceeds the CTA pool");

struct TcParams {
  ConvArgs a;
  const void* wimg;        // weight image [ntile][tap][kchunk][KC/8][NT][8]
  int KC, NKC, NT, NTILES;
  int HALO_H, HALO_W, nvox_h, plane_stride;   // plane_stride in bytes (odd multiple of 16)
  int a_stage_bytes, b_stage_bytes, SA, SB;
  int tiles_h, tiles_w, n_tiles;
  int tmem_cols, acc_stages;
  int w_resident;          // all weights of the layer live in shared memory for the CTA's lifetime (no B ring)
  int prefetch;            // A stages the loaders keep in flight (1..3, < SA)
  int use_tma;             // halo tiles are staged by ONE tensor-TMA box per stage (else 16-byte cp.async copies)
  int row;                 // ROW image: one `pitch`-byte swizzled row per halo voxel (K-major SWIZZLE_64B/128B), TMA only
  int pitch;               // bytes per row (KC * 2 = 64 or 128)
  int smem_a_off, smem_b_off, smem_bar_off, smem_norm_off, smem_gnorm_off, smem_stat_off;
  alignas(64) CUtensorMap tm_x;      // x as {8 ch, w, h, channel plane, b*D + d}, or (row image) {channel, w, h, b*D + d}
};

// Straight-line issue of one staged halo tile against RESIDENT weights: every tap and K step unrolled, descriptor low
// words are `base + compile-time-shaped offsets`, high words constant.  ~5 uniform instructions per tcgen05.mma
// and no loop / constant-bank traffic between them (the 32-channel layers are issue-bound otherwise).
template <int KS, int KH, int KW>
__device__ __forceinline__ void issue_stage_resident(uint32_t tmem_d, uint64_t da_stage, uint64_t db_stage, uint32_t idesc,
                                                     uint32_t& accumulate, uint32_t a_kstep, uint32_t b_kstep,
                                                     uint32_t a_rowstep, uint32_t a_tapstep, uint32_t b_tap_step, uint32_t elected) {
  const uint32_t a_lo = (uint32_t)da_stage, a_hi = (uint32_t)(da_stage >> 32);
  const uint32_t b_lo = (uint32_t)db_stage, b_hi = (uint32_t)(db_stage >> 32);
#pragma unroll
  for (int zh = 0; zh < KH; ++zh) {
#pragma unroll
    for (int zw = 0; zw < KW; ++zw) {
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)zh * a_rowstep + (uint32_t)zw * a_tapstep + (uint32_t)j * a_kstep);
        const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(zh * KW + zw) * b_tap_step + (uint32_t)j * b_kstep);
        if (elected) umma_f16(tmem_d, da, db, idesc, accumulate);
        accumulate = 1;
      }
    }
  }
}

struct TileCoord { int b, d, h0, w0, ntile; };
// Persistent tile walk t = blockIdx.x, +gridDim.x, ... as a mixed-radix counter (ntile, w-tile, h-tile, d, b):
// one set of divisions per kernel instead of four per tile per warp role (~1000 cycles/tile measured).
// TileWalk = the constants of the walk (radices, digits of the stride); TileIter = one position on it.
struct TileWalk {
  int s0, s1, s2, s3, s4;           // digits of the stride
  int r0, r1, r2, r3;               // radices
  int n_tiles, stride;
  __device__ __forceinline__ void init(const TcParams& p) {
    r0 = p.NTILES; r1 = p.tiles_w; r2 = p.tiles_h; r3 = p.a.D;
    n_tiles = p.n_tiles; stride = gridDim.x;
    int x = stride;
    s0 = x % r0; x /= r0; s1 = x % r1; x /= r1; s2 = x % r2; x /= r2; s3 = x % r3; s4 = x / r3;
  }
};
struct TileIter {
  int ntile, wi, hi, d, b, t;       // current digits, linear index
  __device__ __forceinline__ void init(const TileWalk& k) {
    t = blockIdx.x;
    int x = t;
    ntile = x % k.r0; x /= k.r0; wi = x % k.r1; x /= k.r1; hi = x % k.r2; x /= k.r2; d = x % k.r3; b = x / k.r3;
  }
  __device__ __forceinline__ bool valid(const TileWalk& k) const { return t < k.n_tiles; }
  __device__ __forceinline__ TileCoord coord() const { TileCoord c; c.b = b; c.d = d; c.h0 = hi * TH; c.w0 = wi * TW; c.ntile = ntile; return c; }
  __device__ __forceinline__ void next(const TileWalk& k) {
    t += k.stride;
    int c;
    ntile += k.s0; c = ntile >= k.r0; if (c) ntile -= k.r0;
    wi += k.s1 + c; c = wi >= k.r1; if (c) wi -= k.r1;
    hi += k.s2 + c; c = hi >= k.r2; if (c) hi -= k.r2;
    d += k.s3 + c; c = d >= k.r3; if (c) d -= k.r3;
    b += k.s4 + c;
  }
};

// One A stage = (tile, K chunk, depth tap) with an in-volume input slice; the loaders walk them in exactly the order
// the MMA warp consumes them: for tile { for kc { for zd { skip if din outside the volume } } }.
struct StageCursor {
  TileIter ti; int kc, zd, din;
  __device__ __forceinline__ void init(const TileWalk& k, const TcParams& p) {
    ti.init(k); kc = 0; zd = -1;
    if (ti.valid(k)) next(k, p);
  }
  __device__ __forceinline__ bool valid(const TileWalk& k) const { return ti.valid(k); }
  __device__ __forceinline__ void next(const TileWalk& k, const TcParams& p) {
    for (;;) {
      if (++zd == p.a.kd) { zd = 0; if (++kc == p.NKC) { kc = 0; ti.next(k); if (!ti.valid(k)) return; } }
      din = ti.d + zd - p.a.kd / 2;
      if ((unsigned)din < (unsigned)p.a.D) return;
    }
  }
};

// ------------------------------------------------------------------ A loaders
// cp.async (LDGSTS) prefetch of P stages + in-place InstanceNorm/ReLU once a stage has landed.  Each thread owns ONE
// 8-channel plane and a fixed set of (at most kMaxChunks) halo voxels of it, copies exactly those 16-byte chunks and
// later transforms exactly those chunks, so no cross-thread synchronisation is needed between copy and transform.
// Everything that does not depend on the tile (voxel slot, offset from the tile origin, halo row / column) is computed
// once per thread; per stage the work is one pointer add per chunk, and the bounds tests vanish for interior tiles.
constexpr int kMaxChunks = 6;      // ceil(180 halo voxels / (256 threads / (KC/8) planes)) for KC <= 64

// normalise + activate three 8-channel chunks held in registers.  Straight-line on purpose: the three dependency chains
// interleave (the bounds tests only predicate the shared-memory load and store around this)
template <bool RELU>
__device__ __forceinline__ void transform3(uint4 (&raw)[3], const float (&sc)[8], const float (&sf)[8], float slope) {
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    __half2* hv = reinterpret_cast<__half2*>(&raw[u]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(hv[j]);
      f.x = fmaf(f.x, sc[2 * j], sf[2 * j]); f.y = fmaf(f.y, sc[2 * j + 1], sf[2 * j + 1]);
      if (RELU) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
      else { f.x = act_apply_s(f.x, slope); f.y = act_apply_s(f.y, slope); }
      hv[j] = __floats2half2_rn(f.x, f.y);
    }
  }
}

template <int P, bool TMA>
__device__ __forceinline__ void loader_role(const TcParams& p, uint8_t* smem, const float2* s_norm, uint32_t bar0) {
  const ConvArgs& a = p.a;
  const int lt = threadIdx.x - kLoadWarp0 * 32;
  const int cpv = p.KC / 8;
  const int vstep = kLoadThreads / cpv;
  const bool active = lt < vstep * cpv;
  const int c8 = lt % cpv, v0 = lt / cpv;
  const int ph = a.kh / 2, pw = a.kw / 2;
  const bool xform = (a.x_stats != nullptr) || (a.act != 0);
  const int act = a.act;
  const __half* xbase = reinterpret_cast<const __half*>(a.x);
  const uint32_t smem_a = smem_u32(smem + p.smem_a_off) + (uint32_t)(c8 * p.plane_stride);
  uint8_t* smem_a_gen = smem + p.smem_a_off + c8 * p.plane_stride;
  auto A_FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto A_EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.SA + i); };
  auto A_LAND = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + 2 * p.SB + 4 + 1 + i); };
  // TMA mode: thread 0 stages the halo tile with ONE tensor-TMA box per stage; every thread then transforms exactly
  // the chunks it would have copied (same table, same code) once the box has landed.  Nobody but thread 0 walks the
  // issue cursor.
  const uint32_t stage_tx = (uint32_t)(cpv * p.nvox_h * 16);

  // per-thread chunk table
  int rel[kMaxChunks];            // element offset of the chunk's voxel from the (possibly out-of-volume) tile origin
  uint32_t hw[kMaxChunks];        // halo row << 16 | halo column ; 0xffffffff = no such chunk
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int v = v0 + i * vstep;
    const bool have = active && v < p.nvox_h;
    const int hh = v / p.HALO_W, ww = v % p.HALO_W;
    if constexpr (TMA) {
      // byte offset of the chunk inside a stage: plane image [c8][v][8 ch], or row image (one pitch-byte row per voxel,
      // 16-byte chunks XOR-swizzled with address bits 7.. exactly as the TMA unit wrote them; stages are 1024-aligned)
      rel[i] = p.row ? v * p.pitch + ((c8 ^ (p.pitch == 128 ? (v & 7) : ((v >> 1) & 3))) << 4) : c8 * p.plane_stride + v * 16;
    } else {
      rel[i] = (hh * a.W + ww) * a.x_ld;
    }
    hw[i] = have ? ((uint32_t)hh << 16) | (uint32_t)ww : 0xffffffffu;
  }

  const int nch = (p.nvox_h + vstep - 1) / vstep;       // chunks per thread actually present (warp-uniform loop bound)

  TileWalk tw; tw.init(p);
  StageCursor ci, cd;
  ci.init(tw, p); cd.init(tw, p);
  Ring ri, rd; ri.init(p.SA); rd.init(p.SA);

  auto issue = [&]() {
    if constexpr (TMA) {
      if (lt == 0) {
        mbar_wait(A_EMPTY(ri.idx), ri.phase ^ 1, 1);
        mbar_arrive_expect_tx(A_LAND(ri.idx), stage_tx);
        const uint32_t dst = smem_u32(smem + p.smem_a_off) + (uint32_t)(ri.idx * p.a_stage_bytes);
        if (p.row) tma_load_4d(dst, &p.tm_x, A_LAND(ri.idx), ci.kc * p.KC, ci.ti.wi * TW - pw, ci.ti.hi * TH - ph, ci.ti.b * a.D + ci.din);
        else tma_load_5d(dst, &p.tm_x, A_LAND(ri.idx), 0, ci.ti.wi * TW - pw, ci.ti.hi * TH - ph, ci.kc * cpv, ci.ti.b * a.D + ci.din);
        ri.advance(); ci.next(tw, p);
      }
    } else {
    mbar_wait(A_EMPTY(ri.idx), ri.phase ^ 1, 1);
    const uint32_t dst = smem_a + (uint32_t)(ri.idx * p.a_stage_bytes) + (uint32_t)v0 * 16u;
    const int hb = ci.ti.hi * TH - ph, wb = ci.ti.wi * TW - pw;
    const bool interior = hb >= 0 && wb >= 0 && hb + p.HALO_H <= a.H && wb + p.HALO_W <= a.W;
    const __half* xs = xbase + (((int64_t)(ci.ti.b * a.D + ci.din) * a.H + hb) * a.W + wb) * a.x_ld + a.x_coff + ci.kc * p.KC + c8 * 8;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      if (i < nch && hw[i] != 0xffffffffu) {
        const bool ok = interior || (((unsigned)(hb + (int)(hw[i] >> 16)) < (unsigned)a.H) && ((unsigned)(wb + (int)(hw[i] & 0xffffu)) < (unsigned)a.W));
        cp_async16(dst + (uint32_t)(i * vstep) * 16u, ok ? (const void*)(xs + rel[i]) : (const void*)xbase, ok ? 16u : 0u);
      }
    }
    ri.advance(); ci.next(tw, p);
    }
  };

  const float slope = act_slope(act);
  float sc[8], sf[8];                              // x*sc + sf == (x - mean) * rstd for this thread's 8 channels
  int norm_key = -1;                               // (b, kc) the constants belong to
#pragma unroll
  for (int i = 0; i < P; ++i) { if (ci.valid(tw)) issue(); if constexpr (!TMA) cp_async_commit(); }
  while (cd.valid(tw)) {
    if constexpr (TMA) mbar_wait(A_LAND(rd.idx), rd.phase, 7);
    else { TC_PROF(11); cp_async_wait<P - 1>(); }      // this thread's copies of the oldest stage have landed
    if (xform && active) {
      TC_PROF(12);
      const int key = cd.ti.b * p.NKC + cd.kc;
      if (key != norm_key) {
        norm_key = key;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 mr = s_norm[cd.ti.b * a.Cin + cd.kc * p.KC + c8 * 8 + j];
          sc[j] = mr.y; sf[j] = -mr.x * mr.y;
        }
      }
      uint8_t* sp = TMA ? smem + p.smem_a_off + rd.idx * p.a_stage_bytes : smem_a_gen + rd.idx * p.a_stage_bytes + v0 * 16;
      auto chunk = [&](int i) { return TMA ? sp + rel[i] : sp + (i * vstep) * 16; };
      const int hb = cd.ti.hi * TH - ph, wb = cd.ti.wi * TW - pw;
      const bool interior = hb >= 0 && wb >= 0 && hb + p.HALO_H <= a.H && wb + p.HALO_W <= a.W;
      // two batches of three chunks: all loads of a batch are issued before the first use (ILP for the one loader
      // warp each scheduler has), zero-filled padding voxels are left untouched (the conv pads the NORMALISED tensor)
#pragma unroll
      for (int i0 = 0; i0 < kMaxChunks; i0 += 3) {
        if (i0 >= nch) break;
        uint4 raw[3]; bool ok[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int i = i0 + u;
          ok[u] = i < nch && hw[i] != 0xffffffffu &&
                  (interior || (((unsigned)(hb + (int)(hw[i] >> 16)) < (unsigned)a.H) && ((unsigned)(wb + (int)(hw[i] & 0xffffu)) < (unsigned)a.W)));
          if constexpr (TMA) raw[u] = make_uint4(0, 0, 0, 0);
          if (ok[u]) raw[u] = *reinterpret_cast<const uint4*>(chunk(i));
        }
        if constexpr (TMA) {
          if (act == B200SEG_ACT_RELU) transform3<true>(raw, sc, sf, slope);
          else transform3<false>(raw, sc, sf, slope);
#pragma unroll
          for (int u = 0; u < 3; ++u)
            if (ok[u]) *reinterpret_cast<uint4*>(chunk(i0 + u)) = raw[u];
        } else {      // 64 registers: one chunk at a time
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            if (!ok[u]) continue;
            __half2* hv = reinterpret_cast<__half2*>(&raw[u]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f = __half22float2(hv[j]);
              f.x = fmaf(f.x, sc[2 * j], sf[2 * j]); f.y = fmaf(f.y, sc[2 * j + 1], sf[2 * j + 1]);
              if (act) { f.x = act_apply(f.x, act); f.y = act_apply(f.y, act); }
              hv[j] = __floats2half2_rn(f.x, f.y);
            }
            *reinterpret_cast<uint4*>(chunk(i0 + u)) = raw[u];
          }
        }
      }
    }
    {
      TC_PROF(17);
      fence_proxy_async();          // generic-proxy / cp.async writes -> visible to the tensor core (async proxy)
      mbar_arrive(A_FULL(rd.idx));
    }
    { TC_PROF(18); rd.advance(); cd.next(tw, p); }
    if (ci.valid(tw)) { TC_PROF(15); issue(); }
    if constexpr (!TMA) cp_async_commit();
  }
  if constexpr (!TMA) cp_async_wait<0>();
}

// ---- TMA staging of RAW inputs (every data-gradient launch): one elected loader thread issues ONE tensor-TMA box
// {8 ch, HALO_W, HALO_H, KC/8 planes} per stage — the TMA unit writes the [plane][halo voxel][8 ch] image and zero-fills
// conv padding / ragged tiles — and the MMA warp consumes the stage straight off the TMA's transaction barrier: no
// loader instruction touches the data.  (Inputs that need InstanceNorm / activation go through loader_role<P>, whose TMA
// mode lands the same box and then transforms it in place with the cp.async path's per-thread chunk table.)
__device__ __forceinline__ void loader_role_tma(const TcParams& p, uint8_t* smem, uint32_t bar0) {
  const ConvArgs& a = p.a;
  const int lt = threadIdx.x - kLoadWarp0 * 32;
  const int ph = a.kh / 2, pw = a.kw / 2;
  const uint32_t smem_a = smem_u32(smem + p.smem_a_off);
  auto A_EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.SA + i); };
  auto A_LAND = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + 2 * p.SB + 4 + 1 + i); };
  TileWalk tw; tw.init(p);
  const uint32_t stage_tx = (uint32_t)((p.KC / 8) * p.nvox_h * 16);
  auto issue = [&](const StageCursor& c, int slot) {
    mbar_arrive_expect_tx(A_LAND(slot), stage_tx);
    if (p.row)
      tma_load_4d(smem_a + (uint32_t)(slot * p.a_stage_bytes), &p.tm_x, A_LAND(slot), c.kc * p.KC, c.ti.wi * TW - pw, c.ti.hi * TH - ph,
                  c.ti.b * a.D + c.din);
    else
      tma_load_5d(smem_a + (uint32_t)(slot * p.a_stage_bytes), &p.tm_x, A_LAND(slot), 0, c.ti.wi * TW - pw, c.ti.hi * TH - ph,
                  c.kc * (p.KC / 8), c.ti.b * a.D + c.din);
  };
  if (lt == 0) {
    StageCursor c; c.init(tw, p);
    Ring r; r.init(p.SA);
    for (; c.valid(tw); c.next(tw, p)) {
      mbar_wait(A_EMPTY(r.idx), r.phase ^ 1, 1);
      issue(c, r.idx);
      r.advance();
    }
  }
}

// ------------------------------------------------------------------ epilogue
// One 16-column chunk of one accumulator row: bias / residual / dgrad mask, fp16 rounding, store.  On return r[] holds
// the STORED values (0 for rows outside the volume) and s2[] the second statistics operand (r^2, or g*xhat for dgrad).
__device__ __forceinline__ void epi_chunk(const ConvArgs& a, bool valid, bool dgrad, const uint32_t (&v)[16], const uint4& cur0,
                                          const uint4& cur1, bool has_side, bool is_res, const float2* gnorm, const float* bias,
                                          __half* yout, float (&r)[16], float (&s2)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(v[j]);
  if (!valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { r[j] = 0.f; s2[j] = 0.f; }
    return;
  }
  if (bias) {
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] += bias[j];
  }
  float sv[16];
  if (has_side) {
    const __half2* h0 = reinterpret_cast<const __half2*>(&cur0);
    const __half2* h1 = reinterpret_cast<const __half2*>(&cur1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
      sv[2 * j] = f0.x; sv[2 * j + 1] = f0.y; sv[8 + 2 * j] = f1.x; sv[8 + 2 * j + 1] = f1.y;
    }
  }
  if (dgrad) {
    const float gslope = act_slope(a.g_act);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 mr = gnorm[j];
      const float hx = (sv[j] - mr.x) * mr.y;
      float g = r[j] * act_grad_s(hx, gslope);
      g = __half2float(__float2half_rn(g));
      r[j] = g; s2[j] = g * hx;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = __half2float(__float2half_rn(r[j]));
    if (is_res) {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __half2float(__float2half_rn(r[j] + sv[j]));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s2[j] = r[j] * r[j];
  }
  st8<__half>(yout, reinterpret_cast<const float(&)[8]>(r[0]));
  st8<__half>(yout + 8, reinterpret_cast<const float(&)[8]>(r[8]));
}

// Eight output columns of this thread's row on the register-statistics path: same arithmetic (and the same fp16 rounding
// points) as epi_chunk, but the statistics are accumulated at once and nothing but the eight values is kept live —
// two 16-column chunks per warp (NT = 64) need 64 accumulators, which leaves ~70 registers for everything else.
__device__ __forceinline__ void epi_piece8(const ConvArgs& a, bool valid, bool dgrad, const uint32_t* v, const uint4& side, bool has_side,
                                           const float2* gnorm, const float* bias, __half* yout, float gslope, bool want_stats,
                                           float* as, float* aq) {
  if (!valid) return;                                   // rows outside the volume: nothing stored, nothing counted
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = __uint_as_float(v[j]);
  if (bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias)), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 1);
    r[0] += b0.x; r[1] += b0.y; r[2] += b0.z; r[3] += b0.w; r[4] += b1.x; r[5] += b1.y; r[6] += b1.z; r[7] += b1.w;
  }
  const __half2* hs = reinterpret_cast<const __half2*>(&side);
  if (dgrad) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 mr = *reinterpret_cast<const float4*>(gnorm + 2 * j);      // {mean, rstd} of two channels
      const float2 sv = __half22float2(hs[j]);
      const float h0 = (sv.x - mr.x) * mr.y, h1 = (sv.y - mr.z) * mr.w;
      const float g0 = __half2float(__float2half_rn(r[2 * j] * act_grad_s(h0, gslope)));
      const float g1 = __half2float(__float2half_rn(r[2 * j + 1] * act_grad_s(h1, gslope)));
      r[2 * j] = g0; r[2 * j + 1] = g1;
      if (want_stats) { as[2 * j] += g0; as[2 * j + 1] += g1; aq[2 * j] = fmaf(g0, h0, aq[2 * j]); aq[2 * j + 1] = fmaf(g1, h1, aq[2 * j + 1]); }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y0 = __half2float(__float2half_rn(r[2 * j])), y1 = __half2float(__float2half_rn(r[2 * j + 1]));
      if (has_side) {
        const float2 sv = __half22float2(hs[j]);
        y0 = __half2float(__float2half_rn(y0 + sv.x)); y1 = __half2float(__float2half_rn(y1 + sv.y));
      }
      r[2 * j] = y0; r[2 * j + 1] = y1;
      if (want_stats) { as[2 * j] += y0; as[2 * j + 1] += y1; aq[2 * j] = fmaf(y0, y0, aq[2 * j]); aq[2 * j + 1] = fmaf(y1, y1, aq[2 * j + 1]); }
    }
  }
  st8<__half>(yout, r);
}

// REGSTATS: NT <= 64 (at most two chunks per warp); the InstanceNorm sums live in registers across all tiles.
struct EpiCtx {
  const TcParams* p; int q, half, lane, nchunks; bool dgrad, want_stats;
  float* s_stat; const float2* s_gnorm;
};

// flush one chunk's register accumulators into the quadrant's shared partial sums (once per CTA, or on a batch change)
__device__ __forceinline__ void flush_chunk(const EpiCtx& e, int c, int b, float (&as)[16], float (&aq)[16]) {
  if (c >= e.nchunks) return;
  const float u = column_sum16(as, e.lane), q2 = column_sum16(aq, e.lane);
  if ((e.lane & 1) == 0) {
    float* ws = e.s_stat + ((e.q * e.p->a.B + b) * e.p->a.Cout + 16 * c + column_of_lane16(e.lane)) * 2;
    ws[0] += u; ws[1] += q2;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) { as[j] = 0.f; aq[j] = 0.f; }
}

// one chunk of one tile on the register-statistics path
__device__ __forceinline__ void reg_chunk(const EpiCtx& e, int c, uint32_t trow, uint32_t t_empty, bool valid, const uint4& sd0,
                                          const uint4& sd1, bool has_side, const float2* gn, const float* bias, __half* yp,
                                          float (&as)[16], float (&aq)[16]) {
  if (c >= e.nchunks) return;
  uint32_t v[16];
  { TC_PROF(19); tmem_ld16(trow + (uint32_t)(16 * c), v); tmem_ld_wait(); }
  if (c + 2 >= e.nchunks) {       // last chunk of this warp is in registers: hand the TMEM buffer back now
    tc_fence_before();
    mbar_arrive(t_empty);
  }
  const float gslope = act_slope(e.p->a.g_act);
  epi_piece8(e.p->a, valid, e.dgrad, v, sd0, has_side, gn + 16 * c, bias ? bias + 16 * c : nullptr, yp + 16 * c, gslope, e.want_stats,
             as, aq);
  asm volatile("" ::: "memory");        // keep the two halves sequential (register pressure, see epi_piece8)
  epi_piece8(e.p->a, valid, e.dgrad, v + 8, sd1, has_side, gn + 16 * c + 8, bias ? bias + 16 * c + 8 : nullptr, yp + 16 * c + 8, gslope,
             e.want_stats, as + 8, aq + 8);
}

// Per-row addresses of one tile for this thread
struct EpiTile { bool valid; int b, co_base; __half* yp; const __half* side; };
__device__ __forceinline__ EpiTile epi_tile(const TcParams& p, const TileCoord& tc, int hl, int wl, bool dgrad) {
  const ConvArgs& a = p.a;
  EpiTile t;
  const int h = tc.h0 + hl, w = tc.w0 + wl;
  t.valid = (h < a.H) && (w < a.W);
  t.b = tc.b;
  const int64_t vox = ((int64_t)(tc.b * a.D + tc.d) * a.H + h) * a.W + w;
  t.co_base = tc.ntile * p.NT;
  t.yp = reinterpret_cast<__half*>(a.y) + vox * a.y_ld + a.y_coff + t.co_base;
  // side input of the epilogue: the residual, or x for the dgrad ReLU mask
  t.side = dgrad ? reinterpret_cast<const __half*>(a.gx) + vox * a.gx_ld + a.gx_coff + t.co_base
                 : (a.res ? reinterpret_cast<const __half*>(a.res) + vox * a.r_ld + a.r_coff + t.co_base : nullptr);
  return t;
}

// CPW > 0: NT <= 32*CPW... the warp owns CPW (1 or 2) alternate 16-column chunks and keeps the InstanceNorm sums of
// the stored values in REGISTERS across all tiles of the persistent CTA; the side input (residual / x) of tile t+1 is
// requested before tile t is processed, so its global-memory latency never sits on the per-tile critical path
// (measured: with the request issued at the top of its own tile, ~1.5k cycles per tile were exposed).
// CPW == 0: wide tiles (NT > 64), butterfly statistics; the MMA time of such tiles hides this epilogue.
template <int CPW>
__device__ __forceinline__ void epilogue_role(const TcParams& p, int warp, int lane, uint32_t tmem_base, uint32_t bar0,
                                              const float2* s_gnorm, float* s_stat) {
  const ConvArgs& a = p.a;
  auto T_FULL = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + 2 * p.SB + i); };
  auto T_EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + 2 * p.SB + 2 + i); };
  EpiCtx e;
  e.p = &p; e.q = warp & 3; e.half = warp >> 2; e.lane = lane; e.nchunks = p.NT >> 4;
  e.dgrad = a.gx != nullptr; e.want_stats = a.y_stats != nullptr; e.s_stat = s_stat; e.s_gnorm = s_gnorm;
  const int q = e.q, half = e.half;           // TMEM lane quadrant; which alternate 16-column chunks this warp owns
  const int row = q * 32 + lane;              // GEMM row = hl*8 + wl
  const int hl = row >> 3, wl = row & 7;
  const bool dgrad = e.dgrad, want_stats = e.want_stats;
  const int nchunks = e.nchunks;
  const bool has_side = dgrad || a.res != nullptr;
  int it = 0;
  TileWalk tw; tw.init(p); TileIter ti; ti.init(tw);
  if constexpr (CPW > 0) {
    float as0[16], aq0[16], as1[CPW > 1 ? 16 : 1], aq1[CPW > 1 ? 16 : 1];
#pragma unroll
    for (int j = 0; j < 16; ++j) { as0[j] = 0.f; aq0[j] = 0.f; }
    if (CPW > 1) {
#pragma unroll
      for (int j = 0; j < (CPW > 1 ? 16 : 1); ++j) { as1[j] = 0.f; aq1[j] = 0.f; }
    }
    int acc_b = -1;
    const bool own0 = half < nchunks, own1 = CPW > 1 && half + 2 < nchunks;
    uint4 n00 = make_uint4(0, 0, 0, 0), n01 = n00, n10 = n00, n11 = n00;       // side input of the NEXT tile
    EpiTile cur;
    if (ti.valid(tw)) {
      cur = epi_tile(p, ti.coord(), hl, wl, dgrad);
      if (has_side && cur.valid) {
        if (own0) { n00 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half)); n01 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half + 8)); }
        if (own1) { n10 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half + 32)); n11 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half + 40)); }
      }
    }
    for (; ti.valid(tw); ++it) {
      const int acc = (p.acc_stages == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
      const uint4 s00 = n00, s01 = n01, s10 = n10, s11 = n11;
      const EpiTile t = cur;
      ti.next(tw);
      TC_PROF(21);
      if (ti.valid(tw)) {                       // request the next tile's side input now
        cur = epi_tile(p, ti.coord(), hl, wl, dgrad);
        if (has_side && cur.valid) {
          if (own0) { n00 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half)); n01 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half + 8)); }
          if (own1) { n10 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half + 32)); n11 = __ldg(reinterpret_cast<const uint4*>(cur.side + 16 * half + 40)); }
        }
      }
      if (want_stats && t.b != acc_b) {
        if (acc_b >= 0) { flush_chunk(e, half, acc_b, as0, aq0); if constexpr (CPW > 1) flush_chunk(e, half + 2, acc_b, as1, aq1); }
        acc_b = t.b;
      }
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
      const float2* gn = s_gnorm + t.b * a.Cout + t.co_base;
      const float* bias = a.bias ? a.bias + t.co_base : nullptr;
      mbar_wait(T_FULL(acc), acc_phase, 6);
      tc_fence_after();
      {
        TC_PROF(13);
        reg_chunk(e, half, trow, T_EMPTY(acc), t.valid, s00, s01, has_side, gn, bias, t.yp, as0, aq0);
        if constexpr (CPW > 1) reg_chunk(e, half + 2, trow, T_EMPTY(acc), t.valid, s10, s11, has_side, gn, bias, t.yp, as1, aq1);
      }
      if (!own0) {                              // a warp that owns no chunk of this tile (NT == 16) still has to release it
        tc_fence_before();
        mbar_arrive(T_EMPTY(acc));
      }
    }
    if (want_stats && acc_b >= 0) { flush_chunk(e, half, acc_b, as0, aq0); if constexpr (CPW > 1) flush_chunk(e, half + 2, acc_b, as1, aq1); }
  } else {
    for (; ti.valid(tw); ti.next(tw), ++it) {
      const TileCoord tc = ti.coord();
      const int acc = (p.acc_stages == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
      const EpiTile t = epi_tile(p, tc, hl, wl, dgrad);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
      const float2* gn = s_gnorm + tc.b * a.Cout + t.co_base;
      const float* bias = a.bias ? a.bias + t.co_base : nullptr;
      float* wstat = s_stat + ((q * a.B + tc.b) * a.Cout + t.co_base) * 2;
      uint4 nx0 = make_uint4(0, 0, 0, 0), nx1 = make_uint4(0, 0, 0, 0);
      if (has_side && t.valid && half < nchunks) {
        nx0 = __ldg(reinterpret_cast<const uint4*>(t.side + 16 * half));
        nx1 = __ldg(reinterpret_cast<const uint4*>(t.side + 16 * half + 8));
      }
      mbar_wait(T_FULL(acc), acc_phase, 6);
      tc_fence_after();
      for (int c = half; c < nchunks; c += 2) {
        const int n0 = 16 * c;
        uint32_t v[16];
        tmem_ld16(trow + (uint32_t)n0, v);
        const uint4 cur0 = nx0, cur1 = nx1;
        if (has_side && t.valid && c + 2 < nchunks) {        // one chunk AHEAD
          nx0 = __ldg(reinterpret_cast<const uint4*>(t.side + n0 + 32));
          nx1 = __ldg(reinterpret_cast<const uint4*>(t.side + n0 + 40));
        }
        tmem_ld_wait();
        float r[16], s2[16];
        epi_chunk(a, t.valid, dgrad, v, cur0, cur1, has_side, !dgrad && has_side, gn + n0, bias ? bias + n0 : nullptr, t.yp + n0, r, s2);
        if (want_stats) {
          // column sums over the warp's 32 rows with a halving butterfly: 16 shuffles per quantity
          const float u = column_sum16(r, lane), q2 = column_sum16(s2, lane);
          if ((lane & 1) == 0) {
            const int col = column_of_lane16(lane);
            wstat[(n0 + col) * 2] += u;
            wstat[(n0 + col) * 2 + 1] += q2;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(T_EMPTY(acc));
    }
  }
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TC_PROF(31);
  if (p.row && (smem_u32(smem) & 1023u)) __trap();      // the row image's swizzle phase assumes 1024-byte aligned stages
  const ConvArgs& a = p.a;
  // canonical warp index: the shuffle makes it provably warp-uniform, so the role branches below are uniform
  // branches and the MMA warp's loop compiles to the uniform datapath (UIADD3 + UTCHMMA, no R2UR per operand)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int taps_hw = a.kh * a.kw;
  const int pd = a.kd / 2;

  // barrier block layout (uint64 each): a_full[SA] a_empty[SA] b_full[SB] b_empty[SB] t_full[2] t_empty[2]; then tmem ptr
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
  const uint32_t bar0 = smem_u32(bars);
  auto A_FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto A_EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.SA + i); };
  auto B_FULL = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + i); };
  auto B_EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + p.SB + i); };
  auto T_FULL = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + 2 * p.SB + i); };
  auto T_EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + 2 * p.SB + 2 + i); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(bars + 2 * p.SA + 2 * p.SB + 4);
  auto A_LAND = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.SA + 2 * p.SB + 4 + 1 + i); };   // TMA mode: the stage's box has landed

  float2* s_norm = reinterpret_cast<float2*>(smem + p.smem_norm_off);   // [B][Cin] {mean, rstd}
  float2* s_gnorm = reinterpret_cast<float2*>(smem + p.smem_gnorm_off); // [B][Cout] {mean, rstd} of dgrad_x
  float* s_stat = reinterpret_cast<float*>(smem + p.smem_stat_off);     // [4 quadrants][B][Cout][2]
  const uint32_t smem_a = smem_u32(smem + p.smem_a_off);
  const uint32_t smem_b = smem_u32(smem + p.smem_b_off);

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(A_FULL(i), kLoadThreads); mbar_init(A_EMPTY(i), 1); mbar_init(A_LAND(i), 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(B_FULL(i), 1); mbar_init(B_EMPTY(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(T_FULL(i), 1); mbar_init(T_EMPTY(i), kEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), (uint32_t)p.tmem_cols);
  {
    const double n = (double)a.D * a.H * a.W;
    for (int i = threadIdx.x; i < a.B * a.Cin; i += kThreads) {
      float m = 0.f, r = 1.f;
      if (a.x_stats) stats_to_mean_rstd(a.x_stats + (int64_t)i * 2, n, a.eps, m, r);
      s_norm[i] = make_float2(m, r);
    }
    if (a.gx) {
      for (int i = threadIdx.x; i < a.B * a.Cout; i += kThreads) {
        float m, r;
        stats_to_mean_rstd(a.g_stats + (int64_t)i * 2, n, a.g_eps, m, r);
        s_gnorm[i] = make_float2(m, r);
      }
    }
    const int nstat = a.y_stats ? kStatCopies * a.B * a.Cout * 2 : 0;
    for (int i = threadIdx.x; i < nstat; i += kThreads) s_stat[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int ksteps = p.KC / 16;

  if (warp >= kLoadWarp0 && warp < kWgtWarp) {
    // =========================== A LOADERS ===========================
    setmaxnreg_dec<kRegsLoad>();
    if (p.use_tma && !(a.x_stats || a.act)) loader_role_tma(p, smem, bar0);
    else if (p.use_tma) loader_role<3, true>(p, smem, s_norm, bar0);
    else if (p.prefetch >= 3) loader_role<3, false>(p, smem, s_norm, bar0);
    else loader_role<1, false>(p, smem, s_norm, bar0);
  } else if (warp >= kWgtWarp) {
    setmaxnreg_dec<kRegsMma>();
   if (warp == kWgtWarp) {
    // =========================== WEIGHT PRODUCER (bulk TMA) ===========================
    if (lane == 0) {
      const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.wimg);
      const int taps = a.kd * taps_hw;
      if (p.w_resident) {
        // small layers: the whole weight image is loaded once; the MMA warp indexes it directly
        const int nblobs = p.NTILES * taps * p.NKC;
        mbar_arrive_expect_tx(B_FULL(0), (uint32_t)(nblobs * p.b_stage_bytes));
        for (int i = 0; i < nblobs; ++i)
          bulk_g2s(smem_b + i * p.b_stage_bytes, wimg + (int64_t)i * p.b_stage_bytes, (uint32_t)p.b_stage_bytes, B_FULL(0));
      } else {
        Ring ring; ring.init(p.SB);
        TileWalk tw; tw.init(p); TileIter ti; ti.init(tw);
        for (; ti.valid(tw); ti.next(tw)) {
          const TileCoord tc = ti.coord();
          for (int kc = 0; kc < p.NKC; ++kc) {
            for (int zd = 0; zd < a.kd; ++zd) {
              const int din = tc.d + zd - pd;
              if ((unsigned)din >= (unsigned)a.D) continue;
              for (int thw = 0; thw < taps_hw; ++thw) {
                const int tap = zd * taps_hw + thw;
                mbar_wait(B_EMPTY(ring.idx), ring.phase ^ 1, 2);
                mbar_arrive_expect_tx(B_FULL(ring.idx), (uint32_t)p.b_stage_bytes);
                const uint8_t* src = wimg + ((int64_t)(tc.ntile * taps + tap) * p.NKC + kc) * p.b_stage_bytes;
                bulk_g2s(smem_b + ring.idx * p.b_stage_bytes, src, (uint32_t)p.b_stage_bytes, B_FULL(ring.idx));
                ring.advance();
              }
            }
          }
        }
      }
    }
   } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER ===========================
    // Production pattern: the WHOLE warp runs this loop with warp-uniform values (everything is derived from
    // kernel parameters, blockIdx and shuffled broadcasts), and one elected lane issues the tcgen05 instructions.
    // Uniform values live in uniform registers, so a tcgen05.mma costs a couple of UIADDs instead of an
    // elect / R2UR.BROADCAST sequence per operand, and the loop bounds are copied to locals so no constant-bank
    // load sits on the issue path.
    {
      const uint32_t elected = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      Ring ra, rb; ra.init(p.SA); rb.init(p.SB);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // A operand: plane image = no-swizzle K-major core matrices (LBO = plane, SBO = one halo row of 16-byte slots);
      // row image = K-major SWIZZLE_64B / 128B rows (SBO = one halo row of pitch-byte rows, 32 bytes per K step)
      const uint32_t vox16 = p.row ? (uint32_t)p.pitch >> 4 : 1u;                 // 16-byte units per voxel shift
      const uint64_t a_tmpl = p.row ? make_desc_sw(0, 16u, (uint32_t)(p.HALO_W * p.pitch), p.pitch == 128 ? 2u : 4u)
                                    : make_desc(0, (uint32_t)p.plane_stride, (uint32_t)p.HALO_W * 16u);
      const uint64_t b_tmpl = make_desc(0, (uint32_t)p.NT * 16u, 128u);
      const uint32_t a_kstep = p.row ? 2u : (2u * (uint32_t)p.plane_stride) >> 4, b_kstep = (2u * (uint32_t)p.NT * 16u) >> 4;
      const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4, b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
      const uint32_t smem_a16 = smem_a >> 4, smem_b16 = smem_b >> 4;
      const int kd = a.kd, kh = a.kh, kw = a.kw, NKC = p.NKC, D = a.D, NT = p.NT;
      const uint32_t a_rowstep = (uint32_t)p.HALO_W * vox16;                      // next tap row, in 16-byte units
      const int resident = p.w_resident, acc_stages = p.acc_stages;
      const int taps_all = kd * taps_hw;
      const uint32_t res_step = (uint32_t)NKC * b_stage16;
      // A stage ready: published by the loaders, or (raw input staged by TMA) the TMA's own transaction barrier
      const uint32_t a_ready0 = (p.use_tma && !(a.x_stats || a.act)) ? A_LAND(0) : A_FULL(0);
      int it = 0;
      if (resident) { mbar_wait_nocall(B_FULL(0), 0, 10); tc_fence_after(); }
      TileWalk tw; tw.init(p); TileIter ti; ti.init(tw);
      for (; ti.valid(tw); ti.next(tw), ++it) {
        const TileCoord tc = ti.coord();
        const int acc = (acc_stages == 2) ? (it & 1) : 0;
        const uint32_t acc_phase = (acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait_nocall(T_EMPTY(acc), acc_phase ^ 1, 8);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + (uint32_t)(acc * NT);
        uint32_t accumulate = 0;
        for (int kc = 0; kc < NKC; ++kc) {
          for (int zd = 0; zd < kd; ++zd) {
            const int din = tc.d + zd - pd;
            if ((unsigned)din >= (unsigned)D) continue;
            mbar_wait_nocall(a_ready0 + 8u * (uint32_t)ra.idx, ra.phase, 9);
            tc_fence_after();
            uint64_t da_row = a_tmpl + (uint64_t)(smem_a16 + (uint32_t)ra.idx * a_stage16);
            uint64_t db_res = b_tmpl + (uint64_t)(smem_b16 + (uint32_t)((tc.ntile * taps_all + zd * taps_hw) * NKC + kc) * b_stage16);
            bool done = false;
            if (resident && kh == 3 && kw == 3) {
              done = true;
              switch (ksteps) {
                case 1: issue_stage_resident<1, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, a_rowstep, vox16, res_step, elected); break;
                case 2: issue_stage_resident<2, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, a_rowstep, vox16, res_step, elected); break;
                case 3: issue_stage_resident<3, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, a_rowstep, vox16, res_step, elected); break;
                case 4: issue_stage_resident<4, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, a_rowstep, vox16, res_step, elected); break;
                default: done = false;
              }
            }
            if (!done) {
              for (int zh = 0; zh < kh; ++zh) {
                uint64_t da_tap = da_row;
                for (int zw = 0; zw < kw; ++zw) {
                  uint64_t db;
                  if (resident) {
                    db = db_res;
                    db_res += (uint64_t)res_step;
                  } else {
                    mbar_wait_nocall(B_FULL(rb.idx), rb.phase, 10);
                    tc_fence_after();
                    db = b_tmpl + (uint64_t)(smem_b16 + (uint32_t)rb.idx * b_stage16);
                  }
                  uint64_t da = da_tap;
#pragma unroll 4
                  for (int j = 0; j < ksteps; ++j) {
                    if (elected) umma_f16(tmem_d, da, db, idesc, accumulate);
                    accumulate = 1;
                    da += a_kstep; db += b_kstep;
                  }
                  if (!resident) {
                    if (elected) umma_commit(B_EMPTY(rb.idx));   // weights slot free once these MMAs retire
                    rb.advance();
                  }
                  da_tap += (uint64_t)vox16;        // next tap to the right: one voxel
                }
                da_row += (uint64_t)a_rowstep;      // next tap row: HALO_W voxels
              }
            }
            if (elected) umma_commit(A_EMPTY(ra.idx));         // halo tile free
            ra.advance();
          }
        }
        if (elected) umma_commit(T_FULL(acc));                 // accumulator complete -> epilogue
      }
    }
   }
  } else {
    // =========================== EPILOGUE (warps 0-7) ===========================
    setmaxnreg_inc<kRegsEpi>();
    if (p.NT <= 32) epilogue_role<1>(p, warp, lane, tmem_base, bar0, s_gnorm, s_stat);
    else if (p.NT <= 64) epilogue_role<2>(p, warp, lane, tmem_base, bar0, s_gnorm, s_stat);
    else epilogue_role<0>(p, warp, lane, tmem_base, bar0, s_gnorm, s_stat);
  }

  // ---- teardown: flush the per-CTA InstanceNorm partial sums, free TMEM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.y_stats) {
    const int n = a.B * a.Cout * 2;
    for (int i = threadIdx.x; i < n; i += kThreads) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < kStatCopies; ++q) s += (double)s_stat[q * n + i];
      if (s != 0.0) atomicAdd(&a.y_stats[i], s);
    }
  }
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace

TC_PROF_ENTRY(b200seg_conv_tc_prof)

bool conv3d_tc_shape_ok(int Cin, int Cout, int kd, int kh, int kw, int dtype) {
  if (dtype != B200SEG_F16) return false;
  if (tc_pick_nt(Cout) == 0 || tc_pick_kc(Cin) == 0) return false;
  if (kh > 3 || kw > 3 || kd > 3 || kd < 1 || kh < 1 || kw < 1) return false;
  return true;
}

bool conv3d_fwd_tc_supported(const ConvArgs& a, int dtype) {
  if (!conv3d_tc_shape_ok(a.Cin, a.Cout, a.kd, a.kh, a.kw, dtype)) return false;
  if ((a.x_ld % 8) || (a.x_coff % 8) || (a.y_ld % 8) || (a.y_coff % 8)) return false;
  if (a.res && ((a.r_ld % 8) || (a.r_coff % 8))) return false;
  if (a.gx && ((a.gx_ld % 8) || (a.gx_coff % 8))) return false;
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.y) | reinterpret_cast<uintptr_t>(a.w)) & 15) return false;
  if (a.B * a.Cin > 4096 || a.B * a.Cout > 8192) return false;
  // per-CTA InstanceNorm partial sums / dgrad constants live in shared memory: [B][Cout] tables
  if ((a.y_stats || a.gx) && a.B * a.Cout > 2048) return false;
  return true;
}

// `a.w` must be the TC weight IMAGE ([ntile][tap][kchunk][KC/8][NT][8], b200seg_pack_weight layout=TC).
static int conv3d_fwd_tc_plan(const ConvArgs& a, cudaStream_t st, bool allow_row);

int conv3d_fwd_tc(const ConvArgs& a, int dtype, cudaStream_t st) {
  if (!conv3d_fwd_tc_supported(a, dtype)) return B200SEG_EUNSUPPORTED;
  const int rc = conv3d_fwd_tc_plan(a, st, true);
  return rc == -1 ? conv3d_fwd_tc_plan(a, st, false) : rc;     // -1: the row image does not fit this shape, use planes
}

static int conv3d_fwd_tc_plan(const ConvArgs& a, cudaStream_t st, bool allow_row) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.a = a;
  p.wimg = a.w;
  p.KC = tc_pick_kc(a.Cin); p.NKC = a.Cin / p.KC;
  p.NT = tc_pick_nt(a.Cout); p.NTILES = a.Cout / p.NT;
  p.HALO_H = TH + a.kh - 1; p.HALO_W = TW + a.kw - 1;
  p.nvox_h = p.HALO_H * p.HALO_W;
  int slots = p.nvox_h; if ((slots & 1) == 0) slots += 1;     // odd number of 16-B slots -> conflict-free plane stride
  // Operand staging (measured on every layer shape of the benchmark, profiles/r2_layer_times*.txt):
  //  * RAW inputs (every data-gradient launch) are staged by ONE tensor-TMA box per stage and consumed by the MMA warp
  //    straight off the TMA barrier — as the ROW image (one 128-byte SWIZZLE_128B row per halo voxel, 128 contiguous
  //    bytes per TMA request) when Cin is a multiple of 128 (3-9 % faster than planes there), else as the PLANE image
  //    [KC/8][voxel][8 ch] (16 bytes per request; better while the tile is a single K chunk);
  //  * inputs that need InstanceNorm / activation: TMA planes + in-place transform by the loader warps while Cin <= 64
  //    (5-12 % faster than cp.async), per-thread cp.async copies + transform beyond that.  The row image is available
  //    for them too (B200SEG_CONV_ROW_ALL=1) but measured 25-60 % slower: with the copy off the SM the loader warps'
  //    transform itself is what the tile waits for.
  const bool raw_input = !a.x_stats && a.act == 0;
  const bool tma_ok = !getenv("B200SEG_CONV_NO_TMA") && p.nvox_h <= 192;
  const bool row_pays = (raw_input && p.KC == 64 && p.NKC >= 2) || getenv("B200SEG_CONV_ROW_ALL");
  p.row = (tma_ok && allow_row && row_pays && !getenv("B200SEG_CONV_NO_ROW") &&
           b200seg_make_row_tmap(&p.tm_x, a.x, a.x_ld, a.x_coff, a.Cin, p.KC, a.B * a.D, a.H, a.W, p.HALO_W, p.HALO_H)) ? 1 : 0;
  p.pitch = p.KC * 2;
  if (p.row) p.use_tma = 1;
  else
    p.use_tma = (tma_ok && (raw_input || a.Cin <= 64 || getenv("B200SEG_CONV_TMA_ALL")) &&
                 b200seg_make_act_tmap(&p.tm_x, a.x, a.x_ld, a.x_coff, a.Cin, a.B * a.D, a.H, a.W, p.HALO_W, p.HALO_H, p.KC / 8)) ? 1 : 0;
  p.plane_stride = p.use_tma ? p.nvox_h * 16 : slots * 16;    // a TMA box is written densely
  p.a_stage_bytes = p.row ? p.nvox_h * p.pitch : (p.KC / 8) * p.plane_stride;
  p.a_stage_bytes = (p.a_stage_bytes + 1023) / 1024 * 1024;   // swizzle atoms are 1024-byte aligned
  p.b_stage_bytes = p.KC * p.NT * 2;
  p.tiles_h = (a.H + TH - 1) / TH; p.tiles_w = (a.W + TW - 1) / TW;
  int64_t nt = (int64_t)a.B * a.D * p.tiles_h * p.tiles_w * p.NTILES;
  if (nt > 0x7fffffff) return B200SEG_EUNSUPPORTED;
  p.n_tiles = (int)nt;
  p.acc_stages = (2 * p.NT <= 512) ? 2 : 1;
  int cols = p.acc_stages * p.NT, pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  // shared memory carve-up
  const int norm_bytes = a.B * a.Cin * 8;
  const int stat_bytes = a.y_stats ? kStatCopies * a.B * a.Cout * 2 * 4 : 0;
  const int gnorm_bytes = a.gx ? a.B * a.Cout * 8 : 0;
  const int budget = 227 * 1024 - 1024 - norm_bytes - stat_bytes - gnorm_bytes - 512;
  const int64_t w_total = (int64_t)a.kd * a.kh * a.kw * a.Cin * a.Cout * 2;
  int b_region;
  if (w_total <= 112 * 1024 && w_total + 2 * p.a_stage_bytes <= budget) {
    p.w_resident = 1; p.SB = 1;
    b_region = (int)w_total;
    p.SA = (budget - b_region) / p.a_stage_bytes; if (p.SA > 6) p.SA = 6;
  } else {
    p.w_resident = 0;
    p.SA = 4;
    while (p.SA > 2 && p.SA * p.a_stage_bytes + 3 * p.b_stage_bytes > budget) --p.SA;
    p.SB = (budget - p.SA * p.a_stage_bytes) / p.b_stage_bytes; if (p.SB > 8) p.SB = 8;
    if (p.SB < 2) return p.row ? -1 : B200SEG_EUNSUPPORTED;
    b_region = p.SB * p.b_stage_bytes;
  }
  if (p.SA < 2) return p.row ? -1 : B200SEG_EUNSUPPORTED;
  p.prefetch = p.SA - 1 < 3 ? p.SA - 1 : 3;
  if (p.use_tma && !raw_input && p.SA < 4) {                   // the TMA + transform loader runs three stages ahead
    if (p.row) return -1;
    p.use_tma = 0;
  }
  int off = 0;
  p.smem_a_off = off; off += p.SA * p.a_stage_bytes;
  off = (off + 127) / 128 * 128;
  p.smem_b_off = off; off += b_region;
  off = (off + 15) / 16 * 16;
  p.smem_bar_off = off; off += (3 * p.SA + 2 * p.SB + 5) * 8 + 16;
  off = (off + 15) / 16 * 16;
  p.smem_norm_off = off; off += norm_bytes;
  off = (off + 15) / 16 * 16;
  p.smem_gnorm_off = off; off += gnorm_bytes;
  off = (off + 15) / 16 * 16;
  p.smem_stat_off = off; off += stat_bytes;
  const int smem_bytes = off + 1024;       // slack for the 1024-B alignment of the dynamic segment
  int grid = p.n_tiles < B200SEG_NUM_SMS ? p.n_tiles : B200SEG_NUM_SMS;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  tc_apply_env();
  conv_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  B200_CHECK_LAUNCH("conv_tc_kernel");
  return B200SEG_OK;
}
