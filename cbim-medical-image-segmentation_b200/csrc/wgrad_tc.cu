// wgrad_tc.cu — conv3d weight gradient as tcgen05 GEMMs (sm_100a), fp16 operands, fp32 accumulation in TMEM.
// Replaces cuDNN wgrad behind autograd of nn.Conv3d (train_ddp.py:193/208); what it writes is the fp32
// parameter-gradient tensor DDP all-reduces.
//
//   dW[co][ci][tap] = sum_voxels dy[v][co] * a[v + tap][ci],      a = act(IN(x))
//
// GEMM view per CTA: D_tap[128 co][N ci] += dy^T[128 co x 128 voxels] * a_tap[128 voxels x N ci] for a GROUP of
// in-plane taps of one depth offset zd; K = voxels, accumulated over every voxel tile the CTA owns, so the
// accumulators stay in TMEM for the CTA's whole life (G*N <= 512 columns) and there is ONE epilogue.
//   * both operands are staged as [channel/8][voxel][8 ch] — the same image conv_tc.cu uses — and read by
//     the tensor core as MN-major no-swizzle matrices (core matrix = 8 voxels x 8 channels, 128 B):
//     A = dy tile (16x8 voxels), B = halo tile of `a` ((16+kh-1)x(8+kw-1) voxels); a tap is a shifted
//     B descriptor, exactly like the forward kernel.
//   * round 2: `a` is no longer materialised by a separate pass.  The loader warps stage RAW x with cp.async
//     (up to three stages in flight) and apply InstanceNorm-normalise + ReLU in place in shared memory once a stage
//     has landed (zero-filled padding voxels stay zero), exactly like conv_tc.cu's forward loader.
//   * only the REAL output channels of the M tile are staged (Cout = 32 stages 4 of the 16 planes); the tensor core
//     still reads 16 planes, the rows it computes from whatever follows are never read back.
//   * split-K over voxel tiles fills the machine: grid = jobs x S.  Every CTA adds its partial D tiles straight into
//     dW with fp32 reductions (red.global.add.f32) — no partial workspace, no second kernel.  The sum order of the S
//     partials is not fixed, so dW is reproducible to fp32 rounding (~1e-7 relative), like cuDNN's default wgrad.
//   * narrow layers (Cin tile <= 64) run in TS mode: dy^T is transposed smem -> TMEM once per voxel tile by the four
//     (otherwise idle) epilogue warps and the MMAs take their A operand from tensor memory.
// Warp roles (416 threads, 1 CTA/SM): warps 0-3 TS transposers + epilogue, warps 4-11 loaders, warp 12 MMA issue + TMEM alloc.
#include "common.cuh"
#include "conv_args.h"
#include "tc_common.cuh"
#include "tmap.h"
#include <string.h>
#include <stdlib.h>

namespace {

using namespace tc;

constexpr int TH = 16, TW = 8;
constexpr int kEpiWarps = 4;
constexpr int kLoadWarp0 = 4;
constexpr int kMmaWarp = 12;                // highest warp id = highest issue priority in its SM sub-partition
constexpr int kLoadThreads = 256;           // all loader warps cooperate on every stage
constexpr int kThreads = 13 * 32;   // 416
constexpr int MT = 128;                    // output-channel tile (GEMM M)

struct WgParams {
  const __half* x; int x_ld, x_coff;       // raw input; normalised + activated on the fly when x_stats / act
  const double* x_stats; float eps; int act;
  const __half* dy; int dy_ld, dy_coff;
  float* dw;
  int B, D, H, W, Cin, Cout, kd, kh, kw;
  int NTC, ci_tiles, co_tiles, ngroups, gbase, grem, S;
  int HALO_H, HALO_W, nvox_h, a_plane, dy_plane, a_bytes, dy_bytes, stage_bytes, NS, prefetch;
  int tiles_h, tiles_w, nvt;
  int tmem_cols;
  int ts;                                  // A operand (dy^T) is fed from TENSOR MEMORY (narrow Cin: see file header)
  int use_tma;                             // operand tiles are staged by tensor-TMA boxes (else 16-byte cp.async)
  int dy_rows;                             // the dy tile arrives by tensor-TMA as 128-byte swizzled ROWS (see wg_loader)
  int a_rows;                              // NTC == 64: so does the halo tile of x (one {64, HALO_W, HALO_H} box), transformed in place
  int smem_bar_off, smem_norm_off;
  alignas(64) CUtensorMap tm_dy;           // dy  as {8 ch, w, h, plane, b*D+d}
  alignas(64) CUtensorMap tm_x;            // x   likewise
  alignas(64) CUtensorMap tm_dyrow;        // dy  as {channel, w, h, b*D+d}, box {64, 8, 16, 1}, SWIZZLE_128B
  alignas(64) CUtensorMap tm_arow;         // x   as {channel, w, h, b*D+d}, box {64, HALO_W, HALO_H, 1}, SWIZZLE_128B
};
constexpr int kDyRowBox = TH * TW * 128;   // bytes of one 64-channel row-image box (128 voxels x 128 B)

struct Job { int co_tile, ci_tile, zd, grp, tap0, ntaps, s; };
__device__ __forceinline__ Job decode_job(const WgParams& p, int bid) {
  Job j;
  j.s = bid % p.S; int q = bid / p.S;
  j.grp = q % p.ngroups; q /= p.ngroups;
  j.zd = q % p.kd; q /= p.kd;
  j.ci_tile = q % p.ci_tiles; j.co_tile = q / p.ci_tiles;
  if (j.grp < p.grem) { j.ntaps = p.gbase + 1; j.tap0 = j.grp * (p.gbase + 1); }
  else { j.ntaps = p.gbase; j.tap0 = p.grem * (p.gbase + 1) + (j.grp - p.grem) * p.gbase; }
  return j;
}

// voxel tiles this CTA owns: vt = s, s+S, ... ; those whose input depth slice lies outside the volume are skipped.
// The walk is a mixed-radix counter (w-tile, h-tile, d, b) advanced by the digits of S: no divisions per tile.
struct VtWalk {
  int s0, s1, s2, s3;               // digits of the stride S
  int r0, r1, r2;                   // radices: tiles_w, tiles_h, D
  __device__ __forceinline__ void init(const WgParams& p) {
    r0 = p.tiles_w; r1 = p.tiles_h; r2 = p.D;
    int x = p.S;
    s0 = x % r0; x /= r0; s1 = x % r1; x /= r1; s2 = x % r2; s3 = x / r2;
  }
};
struct VtCursor {
  int vt, wi, hi, d, b, din;
  __device__ __forceinline__ void step(const VtWalk& k, const WgParams& p) {
    vt += p.S;
    int c;
    wi += k.s0; c = wi >= k.r0; if (c) wi -= k.r0;
    hi += k.s1 + c; c = hi >= k.r1; if (c) hi -= k.r1;
    d += k.s2 + c; c = d >= k.r2; if (c) d -= k.r2;
    b += k.s3 + c;
  }
  __device__ __forceinline__ void seek(const VtWalk& k, const WgParams& p, int zoff) {      // first valid tile at or after vt
    while (vt < p.nvt) { din = d + zoff; if ((unsigned)din < (unsigned)p.D) return; step(k, p); }
  }
  __device__ __forceinline__ void init(const VtWalk& k, const WgParams& p, int s, int zoff) {
    vt = s;
    int t = s;
    wi = t % k.r0; t /= k.r0; hi = t % k.r1; t /= k.r1; d = t % k.r2; b = t / k.r2;
    seek(k, p, zoff);
  }
  __device__ __forceinline__ bool valid(const WgParams& p) const { return vt < p.nvt; }
  __device__ __forceinline__ void next(const VtWalk& k, const WgParams& p, int zoff) { step(k, p); seek(k, p, zoff); }
  __device__ __forceinline__ int h0() const { return hi * TH; }
  __device__ __forceinline__ int w0() const { return wi * TW; }
};

template <int P>
__device__ __forceinline__ void wg_loader(const WgParams& p, const Job& job, uint8_t* smem, const float2* s_norm, uint32_t bar0) {
  const int lt = threadIdx.x - kLoadWarp0 * 32;
  const int ph = p.kh / 2, pw = p.kw / 2, zoff = job.zd - p.kd / 2;
  const int co0 = job.co_tile * MT;
  const int co_real = min(MT, p.Cout - co0);
  const int ci0 = job.ci_tile * p.NTC;
  // dy tile: thread owns plane (lt % cpv) and walks voxels v0, v0+vstep, ...
  const int cpv_d = co_real / 8;
  const int vstep_d = kLoadThreads / cpv_d;
  const bool act_d = lt < vstep_d * cpv_d;
  const int c8_d = lt % cpv_d, v0_d = lt / cpv_d;
  const int cpv_a = p.NTC / 8;
  const int vstep_a = kLoadThreads / cpv_a;
  const bool act_a = lt < vstep_a * cpv_a;
  const int c8_a = lt % cpv_a, v0_a = lt / cpv_a;
  const int sh_a = vstep_a / p.HALO_W, sw_a = vstep_a % p.HALO_W;
  const int hh0 = v0_a / p.HALO_W, ww0 = v0_a % p.HALO_W;
  const bool xform = (p.x_stats != nullptr) || (p.act != 0);
  auto FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.NS + i); };
  auto LAND = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.NS + 5 + i); };
  VtWalk vw; vw.init(p);
  VtCursor ci, cd;
  ci.init(vw, p, job.s, zoff); cd.init(vw, p, job.s, zoff);
  Ring ri, rd; ri.init(p.NS); rd.init(p.NS);
  const int dy_rows = p.dy_rows, nbox = (co_real + 63) / 64;       // a 32- / 48-channel tile is one box whose upper channels are zero-filled

  auto issue = [&]() {
    mbar_wait(EMPTY(ri.idx), ri.phase ^ 1, 1);
    const uint32_t sdy = smem_u32(smem + ri.idx * p.stage_bytes);
    const uint32_t sa = sdy + (uint32_t)p.dy_bytes;
    if (dy_rows) {
      // dy is raw and (in TS mode) only ever read by the transposer warps: one thread hands it to the TMA unit as one or
      // two {64 ch, 8, 16} boxes of 128-byte swizzled rows (ragged tiles zero-filled) — 32 KB per stage that no longer
      // pass through 2 048 16-byte cp.async requests of the loader warps (role profile: the loaders' issue loop,
      // 3 700 cycles per stage, was what the MMA warp waited for).  Completion is signalled on the stage's LAND barrier.
      if (lt == 0) {
        mbar_arrive_expect_tx(LAND(ri.idx), (uint32_t)(nbox * kDyRowBox));
        for (int hb = 0; hb < nbox; ++hb)
          tma_load_4d(sdy + (uint32_t)(hb * kDyRowBox), &p.tm_dyrow, LAND(ri.idx), co0 + hb * 64, ci.w0(), ci.h0(), ci.b * p.D + ci.d);
      }
    } else if (act_d) {
      const __half* src = p.dy + ((int64_t)(ci.b * p.D + ci.d) * p.H * p.W) * p.dy_ld + p.dy_coff + co0 + c8_d * 8;
      const uint32_t dst = sdy + (uint32_t)(c8_d * p.dy_plane);
#pragma unroll 4
      for (int v = v0_d; v < TH * TW; v += vstep_d) {
        const int h = ci.h0() + (v >> 3), w = ci.w0() + (v & 7);
        const bool ok = h < p.H && w < p.W;
        cp_async16(dst + (uint32_t)v * 16u, ok ? (const void*)(src + ((int64_t)h * p.W + w) * p.dy_ld) : (const void*)p.dy, ok ? 16u : 0u);
      }
    }
    if (act_a) {
      const __half* src = p.x + ((int64_t)(ci.b * p.D + ci.din) * p.H * p.W) * p.x_ld + p.x_coff + ci0 + c8_a * 8;
      const uint32_t dst = sa + (uint32_t)(c8_a * p.a_plane);
      int hh = hh0, ww = ww0;
#pragma unroll 4
      for (int v = v0_a; v < p.nvox_h; v += vstep_a) {
        const int h = ci.h0() - ph + hh, w = ci.w0() - pw + ww;
        const bool ok = (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
        cp_async16(dst + (uint32_t)v * 16u, ok ? (const void*)(src + ((int64_t)h * p.W + w) * p.x_ld) : (const void*)p.x, ok ? 16u : 0u);
        hh += sh_a; ww += sw_a;
        if (ww >= p.HALO_W) { ww -= p.HALO_W; ++hh; }
      }
    }
    ri.advance(); ci.next(vw, p, zoff);
  };

#pragma unroll
  for (int i = 0; i < P; ++i) { if (ci.valid(p)) issue(); cp_async_commit(); }
  while (cd.valid(p)) {
    { TC_PROF(11); cp_async_wait<P - 1>(); }
    if (xform && act_a) {
      TC_PROF(12);
      uint8_t* sp = smem + rd.idx * p.stage_bytes + p.dy_bytes + c8_a * p.a_plane;
      float sc[8], sf[8];                          // x*sc + sf == (x - mean) * rstd
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 mr = s_norm[cd.b * p.NTC + c8_a * 8 + j];
        sc[j] = mr.y; sf[j] = -mr.x * mr.y;
      }
      const int act = p.act;
      int hh = hh0, ww = ww0;
#pragma unroll 2
      for (int v = v0_a; v < p.nvox_h; v += vstep_a) {
        const int h = cd.h0() - ph + hh, w = cd.w0() - pw + ww;
        if ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W) {      // padding voxels stay zero
          uint4 raw = *reinterpret_cast<const uint4*>(sp + v * 16);
          __half2* hv = reinterpret_cast<__half2*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(hv[j]);
            f.x = fmaf(f.x, sc[2 * j], sf[2 * j]); f.y = fmaf(f.y, sc[2 * j + 1], sf[2 * j + 1]);
            if (act) { f.x = act_apply(f.x, act); f.y = act_apply(f.y, act); }
            hv[j] = __floats2half2_rn(f.x, f.y);
          }
          *reinterpret_cast<uint4*>(sp + v * 16) = raw;
        }
        hh += sh_a; ww += sw_a;
        if (ww >= p.HALO_W) { ww -= p.HALO_W; ++hh; }
      }
    }
    fence_proxy_async();
    mbar_arrive(FULL(rd.idx));
    rd.advance(); cd.next(vw, p, zoff);
    if (ci.valid(p)) { TC_PROF(15); issue(); }
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ---- TMA staging: one elected loader thread issues two tensor-TMA boxes per stage (dy tile, x halo tile; out-of-volume
// voxels zero-filled by the TMA unit), as many stages ahead as the ring has free slots.  When the input needs
// InstanceNorm / activation, all loader threads then transform the landed halo tile in place and publish FULL; raw
// operands are consumed straight off the TMA's own barrier (LAND).
__device__ __forceinline__ void wg_loader_tma(const WgParams& p, const Job& job, uint8_t* smem, const float2* s_norm, uint32_t bar0) {
  const int lt = threadIdx.x - kLoadWarp0 * 32;
  const int ph = p.kh / 2, pw = p.kw / 2, zoff = job.zd - p.kd / 2;
  const int co0 = job.co_tile * MT, ci0 = job.ci_tile * p.NTC;
  const bool xform = (p.x_stats != nullptr) || (p.act != 0);
  auto FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.NS + i); };
  auto LAND = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.NS + 5 + i); };
  VtWalk vw; vw.init(p);
  const uint32_t stage_tx = (uint32_t)(p.dy_bytes + p.a_bytes);
  if (!xform) {
    if (lt == 0) {
      VtCursor c; c.init(vw, p, job.s, zoff);
      Ring r; r.init(p.NS);
      for (; c.valid(p); c.next(vw, p, zoff)) {
        mbar_wait(EMPTY(r.idx), r.phase ^ 1, 1);
        const uint32_t sdy = smem_u32(smem + r.idx * p.stage_bytes);
        mbar_arrive_expect_tx(LAND(r.idx), stage_tx);
        tma_load_5d(sdy, &p.tm_dy, LAND(r.idx), 0, c.w0(), c.h0(), co0 / 8, c.b * p.D + c.d);
        tma_load_5d(sdy + (uint32_t)p.dy_bytes, &p.tm_x, LAND(r.idx), 0, c.w0() - pw, c.h0() - ph, ci0 / 8, c.b * p.D + c.din);
        r.advance();
      }
    }
    return;
  }
  const int cpv_a = p.NTC / 8;
  const int vstep_a = kLoadThreads / cpv_a;
  const bool act_a = lt < vstep_a * cpv_a;
  const int c8_a = lt % cpv_a, v0_a = lt / cpv_a;
  const int sh_a = vstep_a / p.HALO_W, sw_a = vstep_a % p.HALO_W;
  const int hh0 = v0_a / p.HALO_W, ww0 = v0_a % p.HALO_W;
  const int act = p.act;
  VtCursor ci, cd;
  ci.init(vw, p, job.s, zoff); cd.init(vw, p, job.s, zoff);
  Ring ri, rd; ri.init(p.NS); rd.init(p.NS);
  int ahead = 0;
  float sc[8], sf[8];
  int norm_b = -1;
  while (cd.valid(p)) {
    if (lt == 0) {                 // run ahead as far as the ring has free slots; block only when nothing is in flight
      while (ci.valid(p) && ahead < p.NS) {
        if (!mbar_test_wait(EMPTY(ri.idx), ri.phase ^ 1)) { if (ahead > 0) break; mbar_wait(EMPTY(ri.idx), ri.phase ^ 1, 1); }
        const uint32_t sdy = smem_u32(smem + ri.idx * p.stage_bytes);
        mbar_arrive_expect_tx(LAND(ri.idx), stage_tx);
        tma_load_5d(sdy, &p.tm_dy, LAND(ri.idx), 0, ci.w0(), ci.h0(), co0 / 8, ci.b * p.D + ci.d);
        tma_load_5d(sdy + (uint32_t)p.dy_bytes, &p.tm_x, LAND(ri.idx), 0, ci.w0() - pw, ci.h0() - ph, ci0 / 8, ci.b * p.D + ci.din);
        ri.advance(); ci.next(vw, p, zoff); ++ahead;
      }
    }
    mbar_wait(LAND(rd.idx), rd.phase, 7);
    if (act_a) {
      if (cd.b != norm_b) {
        norm_b = cd.b;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 mr = s_norm[cd.b * p.NTC + c8_a * 8 + j];
          sc[j] = mr.y; sf[j] = -mr.x * mr.y;
        }
      }
      uint8_t* sp = smem + rd.idx * p.stage_bytes + p.dy_bytes + c8_a * p.a_plane;
      int hh = hh0, ww = ww0;
#pragma unroll 2
      for (int v = v0_a; v < p.nvox_h; v += vstep_a) {
        const int h = cd.h0() - ph + hh, w = cd.w0() - pw + ww;
        if ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W) {      // zero-filled padding voxels stay zero
          uint4 raw = *reinterpret_cast<const uint4*>(sp + v * 16);
          __half2* hv = reinterpret_cast<__half2*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(hv[j]);
            f.x = act_apply(fmaf(f.x, sc[2 * j], sf[2 * j]), act); f.y = act_apply(fmaf(f.y, sc[2 * j + 1], sf[2 * j + 1]), act);
            hv[j] = __floats2half2_rn(f.x, f.y);
          }
          *reinterpret_cast<uint4*>(sp + v * 16) = raw;
        }
        hh += sh_a; ww += sw_a;
        if (ww >= p.HALO_W) { ww -= p.HALO_W; ++hh; }
      }
    }
    fence_proxy_async();
    mbar_arrive(FULL(rd.idx));
    rd.advance(); cd.next(vw, p, zoff);
    if (lt == 0) --ahead;
  }
}

// ---- row-image staging of BOTH operands (NTC == 64): one loader thread hands the stage to the TMA unit — the dy boxes and one
// {64 ch, HALO_W, HALO_H} box of x, each voxel a 128-byte SWIZZLE_128B row, conv padding and ragged tiles zero-filled — and
// the loader warps only apply InstanceNorm / activation to the landed halo tile: thread = fixed 8-channel chunk c (its
// constants stay in registers), the chunk of voxel row v sits at chunk position c ^ (v & 7).  No cp.async at all: the
// 3 700-cycle issue loop of the role profile is gone, what remains per stage is the ~1 600-cycle transform.
__device__ __forceinline__ void wg_loader_rows(const WgParams& p, const Job& job, uint8_t* smem, const float2* s_norm, uint32_t bar0) {
  const int lt = threadIdx.x - kLoadWarp0 * 32;
  const int ph = p.kh / 2, pw = p.kw / 2, zoff = job.zd - p.kd / 2;
  const int co0 = job.co_tile * MT, ci0 = job.ci_tile * p.NTC;
  const int co_real = min(MT, p.Cout - co0), nbox = (co_real + 63) / 64;
  const bool xform = (p.x_stats != nullptr) || (p.act != 0);
  auto FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.NS + i); };
  auto LAND = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.NS + 5 + i); };
  VtWalk vw; vw.init(p);
  const uint32_t stage_tx = (uint32_t)(nbox * kDyRowBox + p.nvox_h * 128);
  constexpr int cpv = 8;                                   // 16-byte chunks per 128-byte row
  const int vstep = kLoadThreads / cpv;                    // 32 voxel rows per pass
  const int c8 = lt % cpv, v0 = lt / cpv;
  const int sh = vstep / p.HALO_W, sw = vstep % p.HALO_W;
  const int hh0 = v0 / p.HALO_W, ww0 = v0 % p.HALO_W;
  const int act = p.act;
  VtCursor ci, cd;
  ci.init(vw, p, job.s, zoff); cd.init(vw, p, job.s, zoff);
  Ring ri, rd; ri.init(p.NS); rd.init(p.NS);
  int ahead = 0;
  float sc[8], sf[8];
  int norm_b = -1;
  while (cd.valid(p)) {
    if (lt == 0) {                 // run ahead as far as the ring has free slots; block only when nothing is in flight
      while (ci.valid(p) && ahead < p.NS) {
        if (!mbar_test_wait(EMPTY(ri.idx), ri.phase ^ 1)) { if (ahead > 0) break; mbar_wait(EMPTY(ri.idx), ri.phase ^ 1, 1); }
        const uint32_t sdy = smem_u32(smem + ri.idx * p.stage_bytes);
        mbar_arrive_expect_tx(LAND(ri.idx), stage_tx);
        for (int hb = 0; hb < nbox; ++hb)
          tma_load_4d(sdy + (uint32_t)(hb * kDyRowBox), &p.tm_dyrow, LAND(ri.idx), co0 + hb * 64, ci.w0(), ci.h0(), ci.b * p.D + ci.d);
        tma_load_4d(sdy + (uint32_t)p.dy_bytes, &p.tm_arow, LAND(ri.idx), ci0, ci.w0() - pw, ci.h0() - ph, ci.b * p.D + ci.din);
        ri.advance(); ci.next(vw, p, zoff); ++ahead;
      }
    }
    mbar_wait(LAND(rd.idx), rd.phase, 7);
    if (xform) {
      TC_PROF(12);
      if (cd.b != norm_b) {
        norm_b = cd.b;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 mr = s_norm[cd.b * p.NTC + c8 * 8 + j];
          sc[j] = mr.y; sf[j] = -mr.x * mr.y;
        }
      }
      uint8_t* sp = smem + rd.idx * p.stage_bytes + p.dy_bytes;
      int hh = hh0, ww = ww0;
#pragma unroll 2
      for (int v = v0; v < p.nvox_h; v += vstep) {
        const int h = cd.h0() - ph + hh, w = cd.w0() - pw + ww;
        if ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W) {      // zero-filled padding voxels stay zero
          uint4* q = reinterpret_cast<uint4*>(sp + v * 128 + ((c8 ^ (v & 7)) << 4));
          uint4 raw = *q;
          __half2* hv = reinterpret_cast<__half2*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(hv[j]);
            f.x = act_apply(fmaf(f.x, sc[2 * j], sf[2 * j]), act); f.y = act_apply(fmaf(f.y, sc[2 * j + 1], sf[2 * j + 1]), act);
            hv[j] = __floats2half2_rn(f.x, f.y);
          }
          *q = raw;
        }
        hh += sh; ww += sw;
        if (ww >= p.HALO_W) { ww -= p.HALO_W; ++hh; }
      }
    }
    fence_proxy_async();
    mbar_arrive(FULL(rd.idx));
    rd.advance(); cd.next(vw, p, zoff);
    if (lt == 0) --ahead;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TC_PROF(31);
  // canonical warp index: the shuffle makes it provably warp-uniform, so the role branches below are uniform
  // branches and the MMA warp's loop compiles to the uniform datapath (UIADD3 + UTCHMMA, no R2UR per operand)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const Job job = decode_job(p, blockIdx.x);
  const int zoff = job.zd - p.kd / 2;
  const int co0 = job.co_tile * MT;
  const int co_real = min(MT, p.Cout - co0);
  const int ci0 = job.ci_tile * p.NTC;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.NS + i); };
  const uint32_t DONE = bar0 + 8u * (uint32_t)(2 * p.NS);
  auto A_READY = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.NS + 1 + i); };     // TS mode: dy^T tile i sits in TMEM
  auto A_FREE = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.NS + 3 + i); };      //          the MMAs reading it retired
  auto LAND = [&](int i) { return bar0 + 8u * (uint32_t)(2 * p.NS + 5 + i); };         // TMA mode: the stage's boxes have landed
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(bars + 3 * p.NS + 5);
  float2* s_norm = reinterpret_cast<float2*>(smem + p.smem_norm_off);     // [B][NTC] {mean, rstd} of this job's channels

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NS; ++i) { mbar_init(FULL(i), kLoadThreads); mbar_init(EMPTY(i), 1); }
    mbar_init(DONE, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(A_READY(i), kEpiWarps * 32); mbar_init(A_FREE(i), 1); }
    for (int i = 0; i < p.NS; ++i) mbar_init(LAND(i), 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), (uint32_t)p.tmem_cols);
  {
    const double n = (double)p.D * p.H * p.W;
    for (int i = threadIdx.x; i < p.B * p.NTC; i += kThreads) {
      float m = 0.f, r = 1.f;
      const int b = i / p.NTC, c = i % p.NTC;
      if (p.x_stats) stats_to_mean_rstd(p.x_stats + ((int64_t)b * p.Cin + ci0 + c) * 2, n, p.eps, m, r);
      s_norm[i] = make_float2(m, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= kLoadWarp0 && warp < kMmaWarp) {
    // =========================== LOADERS ===========================
    if (p.a_rows) wg_loader_rows(p, job, smem, s_norm, bar0);
    else if (p.use_tma) wg_loader_tma(p, job, smem, s_norm, bar0);
    else if (p.prefetch >= 3) wg_loader<3>(p, job, smem, s_norm, bar0);
    else if (p.prefetch == 2) wg_loader<2>(p, job, smem, s_norm, bar0);
    else wg_loader<1>(p, job, smem, s_norm, bar0);
  } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER ===========================
    {   // whole warp, warp-uniform values, one elected lane issues (see conv_tc.cu)
      const uint32_t elected = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.NTC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t dy_lbo = 128u, dy_sbo = (uint32_t)p.dy_plane;
      const uint32_t a_lbo = (uint32_t)p.HALO_W * 16u, a_sbo = (uint32_t)p.a_plane;
      // lean issue loop (this one thread feeds the tensor core): descriptor templates + constant adds
      // dy as the A operand read from shared memory (non-TS mode): plane image = MN-major no-swizzle core matrices; row
      // image (dy_rows) = the canonical MN-major SWIZZLE_128B operand — lbo = pitch of the 64-channel blocks (one TMA box
      // each), sbo = pitch of the 8-voxel row groups (tools/umma_probe.cu, "mn=1 A=SW128": validated on hardware)
      const int dy_rows = p.dy_rows;
      const uint64_t dy_tmpl = dy_rows ? make_desc_sw(0, (uint32_t)kDyRowBox, 1024u, 2u) : make_desc(0, dy_lbo, dy_sbo);
      const uint64_t dy_kstep = dy_rows ? (uint64_t)((16 * 128) >> 4) : 16;    // 16 voxels: 16 rows of 128 B / two 128-byte voxel groups per plane
      // x halo tile as the B operand: plane image (tap = start shifted by whole 16-byte voxel slots) or row image
      // (MN-major SWIZZLE_128B, tap = start shifted by whole 128-byte rows, 8-row groups HALO_W rows apart: the
      // "shift / gw = 10" cases of tools/umma_probe.cu)
      const int a_rows = p.a_rows;
      const uint32_t vunit = a_rows ? 8u : 1u;                  // 16-byte units per voxel step of the start address
      const uint64_t a_tmpl = a_rows ? make_desc_sw(0, 16384u, (uint32_t)p.HALO_W * 128u, 2u) : make_desc(0, a_lbo, a_sbo);
      const uint32_t a_kstep = (2u * (uint32_t)p.HALO_W * 16u * (p.a_rows ? 8u : 1u)) >> 4;    // two halo rows of voxels per K=16 step
      const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, dy16 = (uint32_t)p.dy_bytes >> 4;
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const int zh0 = job.tap0 / p.kw, zw0 = job.tap0 % p.kw;
      const int ntaps = job.ntaps, kw = p.kw, HALO_W = p.HALO_W, NTC = p.NTC, NS = p.NS;
      int idx = 0; uint32_t phase = 0; uint32_t accumulate = 0;
      VtWalk vw; vw.init(p);
      VtCursor c; c.init(vw, p, job.s, zoff);
      const int ts = p.ts;
      // operands ready: FULL (published by the loaders) or, for raw operands staged by TMA, the TMA's own barrier
      const uint32_t ready0 = (p.use_tma && !(p.x_stats || p.act)) ? LAND(0) : FULL(0);
      const uint32_t idesc_ts = (1u << 4) | (1u << 16) | ((uint32_t)(p.NTC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int gmax = p.gbase + (p.grem ? 1 : 0);
      const uint32_t tmem_a0 = tmem_u + (uint32_t)(gmax * NTC);       // two 64-column dy^T buffers behind the accumulators
      int it = 0;
      for (; c.valid(p); c.next(vw, p, zoff), ++it) {
        mbar_wait_nocall(ready0 + 8u * (uint32_t)idx, phase, 9);
        if (ts) mbar_wait_nocall(A_READY(it & 1), (uint32_t)((it >> 1) & 1), 10);
        else if (dy_rows) mbar_wait_nocall(LAND(idx), phase, 10);      // the MMAs read the TMA-landed dy rows themselves
        tc_fence_after();
        const uint64_t da0 = dy_tmpl + (uint64_t)(smem16 + (uint32_t)idx * stage16);
        uint64_t db_tap = a_tmpl + (uint64_t)(smem16 + (uint32_t)idx * stage16 + dy16 + (uint32_t)(zh0 * HALO_W + zw0) * vunit);
        const uint32_t ta0 = tmem_a0 + (uint32_t)((it & 1) * 64);
        int zw = zw0;
        uint32_t tmem_d = tmem_u;
        for (int tl = 0; tl < ntaps; ++tl) {
          uint64_t da = da0, db = db_tap;
          if (ts) {
#pragma unroll
            for (int j = 0; j < (TH * TW) / 16; ++j) {
              if (elected) umma_f16_ts(tmem_d, ta0 + (uint32_t)(j * 8), db, idesc_ts, (accumulate | (uint32_t)(j > 0)));
              db += a_kstep;
            }
          } else {
#pragma unroll
            for (int j = 0; j < (TH * TW) / 16; ++j) {
              if (elected) umma_f16(tmem_d, da, db, idesc, (accumulate | (uint32_t)(j > 0)));
              da += dy_kstep;      // plane image: 2 voxel rows of the dy tile = 256 B; row image: 16 rows = 2 KB
              db += a_kstep;
            }
          }
          tmem_d += (uint32_t)NTC;
          // next in-plane tap: one voxel to the right, or wrap to the next halo row
          if (++zw == kw) { zw = 0; db_tap += (uint64_t)((uint32_t)(HALO_W - (kw - 1)) * vunit); } else db_tap += vunit;
        }
        accumulate = 1;
        if (elected) {
          umma_commit(EMPTY(idx));
          if (ts) umma_commit(A_FREE(it & 1));
        }
        if (++idx == NS) { idx = 0; phase ^= 1; }
      }
      if (elected) umma_commit(DONE);
    }
  } else if (warp < kEpiWarps) {
    // =========================== (TS mode: dy^T -> TMEM), then EPILOGUE (once) ===========================
    VtWalk vw; vw.init(p);
    VtCursor c; c.init(vw, p, job.s, zoff);
    const bool any = c.valid(p);               // did this CTA process any stage at all?
    if (p.ts) {
      // Narrow layers (Cin <= 64) are bound by the tensor core's shared-memory read of the A operand (4 KB per
      // 128 x N x 16 MMA, ~52 cycles whatever N is).  dy^T does not change between the taps of a tile, so it is
      // moved into TENSOR MEMORY once per tile — lane = output channel, 32-bit column c = voxels 2c, 2c+1 — and the
      // tile's ntaps*8 MMAs read A from there (tcgen05.mma with a TMEM A operand) at N/2 cycles each.
      const int gmax = p.gbase + (p.grem ? 1 : 0);
      const uint32_t tmem_a0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(gmax * p.NTC);
      const int col = warp * 32 + lane;         // this thread's GEMM row = output channel within the M tile
      const bool live = warp * 32 < co_real;     // warp-uniform
      const int dy_rows = p.dy_rows;
      // plane image: element (co, voxel) at plane co/8, slot voxel, channel co%8; row image: 128-byte row per voxel, its
      // 16-byte chunk (co%64)/8 stored at chunk position ((co%64)/8) ^ (voxel & 7) (SWIZZLE_128B), one 16 KB box per 64 co
      const uint8_t* lane_base = dy_rows ? smem + (col >> 6) * kDyRowBox + (col & 7) * 2 : smem + (col >> 3) * p.dy_plane + (col & 7) * 2;
      const int rchunk = (col & 63) >> 3;
      int idx = 0; uint32_t phase = 0; int it = 0;
      for (; c.valid(p); c.next(vw, p, zoff), ++it) {
        mbar_wait((p.use_tma || dy_rows) ? LAND(idx) : FULL(idx), phase, 4);       // dy is never transformed: landed is enough
        mbar_wait(A_FREE(it & 1), (uint32_t)(((it >> 1) & 1) ^ 1), 5);
        if (live) {
          TC_PROF(16);
          tc_fence_after();
          const uint8_t* src = lane_base + idx * p.stage_bytes;
          const uint32_t dst = tmem_a0 + (uint32_t)((it & 1) * 64);
#pragma unroll 2
          for (int kb = 0; kb < 8; ++kb) {       // 16 voxels -> 8 columns
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t lo, hi;
              if (dy_rows) {
                // voxels 16kb+2i and +1: same 8-row swizzle group (2i and 2i+1 < 16 never straddle it), (v & 7) = (2i)&7, +1
                const int v = 16 * kb + 2 * i;
                lo = *reinterpret_cast<const uint16_t*>(src + v * 128 + ((rchunk ^ (v & 7)) << 4));
                hi = *reinterpret_cast<const uint16_t*>(src + (v + 1) * 128 + ((rchunk ^ ((v + 1) & 7)) << 4));
              } else {
                lo = *reinterpret_cast<const uint16_t*>(src + (16 * kb + 2 * i) * 16);
                hi = *reinterpret_cast<const uint16_t*>(src + (16 * kb + 2 * i + 1) * 16);
              }
              w[i] = lo | (hi << 16);
            }
            tmem_st8(dst + (uint32_t)(kb * 8), w);
          }
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(A_READY(it & 1));
        if (++idx == p.NS) { idx = 0; phase ^= 1; }
      }
    }
    mbar_wait(DONE, 0, 3);
    tc_fence_after();
    const int row = warp * 32 + lane;
    if (any && warp * 32 < co_real) {          // warp-uniform: TMEM loads are warp-collective
      const int taps = p.kd * p.kh * p.kw, taps_hw = p.kh * p.kw;
      const bool mine = row < co_real;
      float* drow = p.dw + ((int64_t)(co0 + (mine ? row : 0)) * p.Cin + ci0) * taps + job.zd * taps_hw + job.tap0;
      for (int tl = 0; tl < job.ntaps; ++tl) {
        for (int n0 = 0; n0 < p.NTC; n0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tl * p.NTC + n0), v);
          tmem_ld_wait();
          if (mine) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(drow + (int64_t)(n0 + j) * taps + tl, __uint_as_float(v[j]));   // RED.E.ADD.F32
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// Cin tile (GEMM N).  Measured on the benchmark's layers (profiles/r2_layer_times_wgrad_ntc.txt): when Cin is a multiple of
// 128 a 64-wide tile beats the 128-wide one by 10-30 % (TS mode applies, six taps share one staged tile instead of
// three or four, 72 jobs x S = 144 CTAs instead of 108, three or four stages fit instead of two); 96 stays best for
// Cin = 96 / 192 / 576.
int pick_ntc(int Cin) {
  if (Cin % 16) return 0;
  int cap = 128;
  if (const char* e = getenv("B200SEG_WGRAD_NTC_MAX")) { const int v = atoi(e); if (v >= 16) cap = v; }   // tuning knob
  if (Cin % 128 == 0 && cap >= 64) return 64;
  if (Cin <= cap) return Cin;
  const int c[] = {128, 96, 64, 48, 32, 16};
  for (int v : c) if (v <= cap && Cin % v == 0) return v;
  return 0;
}

bool fill_params(const WgradArgs& a, WgParams& p, bool tma = false, bool dyrows = false, bool arows = false) {
  memset(&p, 0, sizeof(p));
  p.B = a.B; p.D = a.D; p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout; p.kd = a.kd; p.kh = a.kh; p.kw = a.kw;
  p.NTC = pick_ntc(a.Cin);
  if (!p.NTC || a.Cout % 8) return false;
  if (a.B * p.NTC > 2048) return false;
  p.ci_tiles = a.Cin / p.NTC;
  p.co_tiles = (a.Cout + MT - 1) / MT;
  const int taps_hw = a.kh * a.kw;
  // TS mode (dy^T in tensor memory) pays when the MMA is bound by the smem read of A: N = Cin tile <= 64
  p.ts = (p.NTC <= 64 && !getenv("B200SEG_WGRAD_NO_TS")) ? 1 : 0;
  int G = (p.ts ? 384 : 512) / p.NTC; if (G > taps_hw) G = taps_hw;
  p.ngroups = (taps_hw + G - 1) / G;
  p.gbase = taps_hw / p.ngroups; p.grem = taps_hw % p.ngroups;
  p.HALO_H = TH + a.kh - 1; p.HALO_W = TW + a.kw - 1; p.nvox_h = p.HALO_H * p.HALO_W;
  int slots = p.nvox_h; if ((slots & 1) == 0) ++slots;
  p.use_tma = tma ? 1 : 0;                                     // (dense planes: a TMA box is written contiguously)
  p.a_plane = p.use_tma ? p.nvox_h * 16 : slots * 16;
  p.dy_plane = p.use_tma ? TH * TW * 16 : (TH * TW + 1) * 16;
  p.a_bytes = (p.NTC / 8) * p.a_plane; p.a_bytes = (p.a_bytes + 127) / 128 * 128;
  // only the real output-channel planes of the widest M tile are staged; the descriptor's 16-plane footprint beyond
  // them falls on the `a` tile / the next stage / the tail slack (allocated below), whose values feed rows never read
  const int co_max = a.Cout < MT ? a.Cout : MT;
  p.dy_bytes = (co_max / 8) * p.dy_plane; p.dy_bytes = (p.dy_bytes + 127) / 128 * 128;
  if (dyrows) {
    // row image of dy (TS mode, output-channel tiles that are multiples of 64): 16 KB boxes at 1024-byte aligned
    // addresses (SWIZZLE_128B is a function of the shared-memory address), so every stage is a multiple of 1 KB
    if (p.use_tma) return false;
    p.dy_rows = 1;
    p.dy_bytes = ((co_max + 63) / 64) * kDyRowBox;
    p.a_bytes = (p.a_bytes + 1023) / 1024 * 1024;
    if (arows) {
      if (p.NTC != 64) return false;
      p.a_rows = 1;
      p.a_bytes = (p.nvox_h * 128 + 1023) / 1024 * 1024;
    }
  }
  p.stage_bytes = p.a_bytes + p.dy_bytes;
  const int norm_bytes = a.B * p.NTC * 8;
  const int budget = 227 * 1024 - 2048 - 16 * p.dy_plane - norm_bytes;
  p.NS = budget / p.stage_bytes; if (p.NS > 6) p.NS = 6;
  if (p.NS < 2) return false;
  p.prefetch = p.NS - 1 < 3 ? p.NS - 1 : 3;
  p.tiles_h = (a.H + TH - 1) / TH; p.tiles_w = (a.W + TW - 1) / TW;
  const int64_t nvt = (int64_t)a.B * a.D * p.tiles_h * p.tiles_w;
  if (nvt > 0x7fffffff) return false;
  p.nvt = (int)nvt;
  const int64_t jobs = (int64_t)p.co_tiles * p.ci_tiles * a.kd * p.ngroups;
  int S = (int)(B200SEG_NUM_SMS / jobs); if (S < 1) S = 1;
  if (S > p.nvt) S = p.nvt;
  p.S = S;
  const int gmax = p.gbase + (p.grem ? 1 : 0);
  int cols = gmax * p.NTC + (p.ts ? 128 : 0), pow2 = 32; while (pow2 < cols) pow2 <<= 1;
  if (pow2 > 512) return false;
  p.tmem_cols = pow2;
  int off = p.NS * p.stage_bytes + 16 * p.dy_plane;       // + slack for the 16-plane descriptor footprint
  off = (off + 15) / 16 * 16;
  p.smem_bar_off = off; off += (3 * p.NS + 6) * 8 + 16;
  off = (off + 15) / 16 * 16;
  p.smem_norm_off = off;
  return true;
}

}  // namespace

TC_PROF_ENTRY(b200seg_wgrad_tc_prof)

bool conv3d_wgrad_tc_supported(const WgradArgs& a, int dtype) {
  if (dtype != B200SEG_F16) return false;
  if (a.kd > 3 || a.kh > 3 || a.kw > 3) return false;
  if ((a.x_ld % 8) || (a.x_coff % 8) || (a.dy_ld % 8) || (a.dy_coff % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.dy)) & 15) return false;
  if (a.dbias) return false;
  WgParams p;
  return fill_params(a, p);
}

// no workspace any more (kept in the ABI: b200seg_conv3d_wgrad_workspace reports 0 for the tensor-core path)
size_t conv3d_wgrad_tc_workspace(const WgradArgs&) { return 0; }

// dw must be zero-initialised (or hold a gradient to accumulate into): every CTA ADDS its partial sums.
int conv3d_wgrad_tc(const WgradArgs& a, int dtype, void* /*workspace*/, size_t /*ws_bytes*/, cudaStream_t st) {
  if (!conv3d_wgrad_tc_supported(a, dtype)) return B200SEG_EUNSUPPORTED;
  WgParams p;
  // tensor-TMA staging of both operands is opt-in (B200SEG_WGRAD_TMA=1): with 16-byte box rows the TMA unit is slower
  // than 256 threads of cp.async on all but two layer shapes (profiles/r2_layer_times_wgrad_tma.txt)
  const bool want_tma = getenv("B200SEG_WGRAD_TMA") != nullptr;
  fill_params(a, p, want_tma);
  p.x = reinterpret_cast<const __half*>(a.x); p.x_ld = a.x_ld; p.x_coff = a.x_coff;
  p.x_stats = a.x_stats; p.eps = a.eps; p.act = a.act;
  p.dy = reinterpret_cast<const __half*>(a.dy); p.dy_ld = a.dy_ld; p.dy_coff = a.dy_coff;
  p.dw = a.dw;
  if (p.use_tma) {
    const int co_planes = (a.Cout < MT ? a.Cout : MT) / 8;
    if (!b200seg_make_act_tmap(&p.tm_dy, a.dy, a.dy_ld, a.dy_coff, a.Cout, a.B * a.D, a.H, a.W, TW, TH, co_planes) ||
        !b200seg_make_act_tmap(&p.tm_x, a.x, a.x_ld, a.x_coff, a.Cin, a.B * a.D, a.H, a.W, p.HALO_W, p.HALO_H, p.NTC / 8)) {
      fill_params(a, p, false);               // tensor maps unavailable: fall back to the cp.async staging layout
      p.x = reinterpret_cast<const __half*>(a.x); p.x_ld = a.x_ld; p.x_coff = a.x_coff;
      p.x_stats = a.x_stats; p.eps = a.eps; p.act = a.act;
      p.dy = reinterpret_cast<const __half*>(a.dy); p.dy_ld = a.dy_ld; p.dy_coff = a.dy_coff;
      p.dw = a.dw;
    }
  }
  // dy by tensor-TMA as swizzled rows (B200SEG_WGRAD_DYROWS=0: A/B switch)
  if (!p.use_tma) {
    const char* e = getenv("B200SEG_WGRAD_DYROWS");
    WgParams q;
    const char* ea = getenv("B200SEG_WGRAD_AROWS");
    bool done = false;
    // both operands as rows when the Cin tile is 64 wide (A/B, profiles/r2g_layer_times_*: 128->128 k333 233 -> 181 us,
    // 384->256 1170 -> 963 us); B200SEG_WGRAD_AROWS=0 keeps the cp.async plane image for x
    if (!(e && e[0] == '0') && !(ea && ea[0] == '0') && fill_params(a, q, false, true, true) &&
        b200seg_make_row_tmap(&q.tm_dyrow, a.dy, a.dy_ld, a.dy_coff, a.Cout, 64, a.B * a.D, a.H, a.W, TW, TH) &&
        b200seg_make_row_tmap(&q.tm_arow, a.x, a.x_ld, a.x_coff, a.Cin, 64, a.B * a.D, a.H, a.W, q.HALO_W, q.HALO_H))
      done = true;
    if (!done && !(e && e[0] == '0') && fill_params(a, q, false, true) &&
        b200seg_make_row_tmap(&q.tm_dyrow, a.dy, a.dy_ld, a.dy_coff, a.Cout, 64, a.B * a.D, a.H, a.W, TW, TH))
      done = true;
    if (done) {
      q.x = p.x; q.x_ld = p.x_ld; q.x_coff = p.x_coff; q.x_stats = p.x_stats; q.eps = p.eps; q.act = p.act;
      q.dy = p.dy; q.dy_ld = p.dy_ld; q.dy_coff = p.dy_coff; q.dw = p.dw;
      p = q;
    }
  }
  const int64_t jobs = (int64_t)p.co_tiles * p.ci_tiles * a.kd * p.ngroups;
  const int smem_bytes = p.smem_norm_off + a.B * p.NTC * 8 + 64;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = (int)(jobs * p.S);
  tc_apply_env();
  wgrad_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  B200_CHECK_LAUNCH("wgrad_tc_kernel");
  return B200SEG_OK;
}
