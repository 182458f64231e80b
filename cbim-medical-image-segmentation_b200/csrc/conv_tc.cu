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
//     The loader warps apply InstanceNorm-normalise + ReLU while staging (the normalised activation
//     tensor never exists in HBM) and zero-fill padding / ragged-tile voxels.
//   * B operand = weights pre-packed on device into the exact shared-memory image per (ntile,tap,kchunk)
//     ([KC/8][NT][8] fp16), streamed by 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) through an mbarrier ring.
//   * D lives in TMEM (double-buffered when 2*NT <= 512 columns); 4 epilogue warps drain it with
//     tcgen05.ld while the MMA warp already works on the next tile.
// Warp roles (448 threads, 1 CTA/SM, persistent over tiles):
//   warps 0-3  epilogue (TMEM lane quadrant == warp id): bias / residual / dgrad ReLU-mask, fp16 store,
//              InstanceNorm sums (or IN-backward sums) of what was stored
//   warps 4-11 A loaders, two groups of 4 warps working on alternating stages
//   warp  12   weight producer (bulk TMA)
//   warp  13   TMEM alloc + tcgen05.mma issue (highest warp id = highest issue priority)
// Roofline: tensor pipe (dense fp16) for NT>=128; for NT<128 the MMA is bound by the shared-memory read of
// A (SS mode), see DESIGN.md.
#include "common.cuh"
#include "conv_args.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int TH = 16, TW = 8;            // output tile (h, w); M = 128
constexpr int kEpiWarps = 4;
// The SM arbitrates highest-warp-id-first inside a sub-partition (B300_MICROARCH.md): the single MMA-issuing warp
// must never queue behind ALU-heavy loader / epilogue warps, so it gets the highest id (measured: 4-5x faster issue).
constexpr int kLoadWarp0 = 4;
constexpr int kWgtWarp = 12;
constexpr int kMmaWarp = 13;
constexpr int kLoadGroups = 2;
constexpr int kLoadGroupThreads = 128;
constexpr int kThreads = 14 * 32;   // 448
constexpr uint32_t kSpinLimit = 1u << 24;

struct TcParams {
  ConvArgs a;
  const void* wimg;        // weight image [ntile][tap][kchunk][KC/8][NT][8]
  int KC, NKC, NT, NTILES;
  int HALO_H, HALO_W, nvox_h, plane_stride;   // plane_stride in bytes (odd multiple of 16)
  int a_stage_bytes, b_stage_bytes, SA, SB;
  int tiles_h, tiles_w, n_tiles;
  int tmem_cols, acc_stages;
  int w_resident;          // all weights of the layer live in shared memory for the CTA's lifetime (no B ring)
  int smem_a_off, smem_b_off, smem_bar_off, smem_norm_off, smem_gnorm_off, smem_stat_off;
  int* err_flag;
  int debug;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.  The diagnostics
// live out of line so the polling loop stays a handful of instructions (the single MMA-issuing thread runs it
// once per weight tile).
__device__ __noinline__ void mbar_timeout(int* err_flag, int code, uint32_t parity) {
  if (err_flag) atomicExch(err_flag, code);
  printf("b200seg conv_tc: mbarrier timeout code=%d block=%d thread=%d parity=%u\n", code, blockIdx.x, threadIdx.x, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) mbar_timeout(err_flag, code, parity);
  }
}
// The MMA warp's variant: NO function call on the slow path.  A call inside the issue loop makes the compiler keep
// the loop-carried descriptors in vector registers, and every tcgen05.mma then needs five R2UR.BROADCASTs
// (~250 cycles per MMA measured); with the inline trap the whole loop runs on the uniform datapath.
__device__ __forceinline__ void mbar_wait_nocall(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 x fp16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
  return d;                    // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// optional cycle accounting of block 0 (B200SEG_TC_DEBUG=1): where each warp role spends its time
__device__ long long g_tc_dbg[32];
__device__ long long g_tc_trace[4][256];   // [role][event] raw clock64 stamps of block 0 (roles: 0 mma, 1 loader g0, 2 loader g1, 3 epilogue)
#define TRACE(role, idx) do { const int _i = (idx); if (dbg && _i < 256) g_tc_trace[role][_i] = clock64(); } while (0)
#define DBG_ADD(slot, val) do { if (dbg) atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_dbg[slot]), (unsigned long long)(val)); } while (0)

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}

// Straight-line issue of one staged halo tile against RESIDENT weights: every tap and K step unrolled, descriptor low
// words are `base + compile-time-shaped offsets`, high words constant.  ~5 uniform instructions per tcgen05.mma
// and no loop / constant-bank traffic between them (the 32-channel layers are issue-bound otherwise).
template <int KS, int KH, int KW>
__device__ __forceinline__ void issue_stage_resident(uint32_t tmem_d, uint64_t da_stage, uint64_t db_stage, uint32_t idesc,
                                                     uint32_t& accumulate, uint32_t a_kstep, uint32_t b_kstep,
                                                     uint32_t halo_w, uint32_t b_tap_step, uint32_t elected) {
  const uint32_t a_lo = (uint32_t)da_stage, a_hi = (uint32_t)(da_stage >> 32);
  const uint32_t b_lo = (uint32_t)db_stage, b_hi = (uint32_t)(db_stage >> 32);
#pragma unroll
  for (int zh = 0; zh < KH; ++zh) {
#pragma unroll
    for (int zw = 0; zw < KW; ++zw) {
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)zh * halo_w + (uint32_t)zw + (uint32_t)j * a_kstep);
        const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(zh * KW + zw) * b_tap_step + (uint32_t)j * b_kstep);
        if (elected) umma_f16(tmem_d, da, db, idesc, accumulate);
        accumulate = 1;
      }
    }
  }
}

struct Ring {
  int idx; uint32_t phase; int n;
  __device__ __forceinline__ void init(int n_) { idx = 0; phase = 0; n = n_; }
  __device__ __forceinline__ void advance() { if (++idx == n) { idx = 0; phase ^= 1; } }
};

// sum over the 32 lanes of v[j] for every j in 0..15; lane L returns column
// ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)   (v is clobbered)
__device__ __forceinline__ float column_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = (lane & 16) ? v[i] : v[i + 8], keep = (lane & 16) ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (lane & 8) ? v[i] : v[i + 4], keep = (lane & 8) ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 4) ? v[i] : v[i + 2], keep = (lane & 4) ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = (lane & 2) ? v[0] : v[1], keep = (lane & 2) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct TileCoord { int b, d, h0, w0, ntile; };
// Persistent tile walk t = blockIdx.x, +gridDim.x, ... as a mixed-radix counter (ntile, w-tile, h-tile, d, b):
// one set of divisions per kernel instead of four per tile per warp role (~1000 cycles/tile measured).
struct TileIter {
  int ntile, wi, hi, d, b;          // current digits
  int s0, s1, s2, s3, s4;           // digits of the stride
  int r0, r1, r2, r3;               // radices
  int t, n_tiles, stride;
  __device__ __forceinline__ void init(const TcParams& p) {
    r0 = p.NTILES; r1 = p.tiles_w; r2 = p.tiles_h; r3 = p.a.D;
    n_tiles = p.n_tiles; stride = gridDim.x; t = blockIdx.x;
    int x = t;
    ntile = x % r0; x /= r0; wi = x % r1; x /= r1; hi = x % r2; x /= r2; d = x % r3; b = x / r3;
    x = stride;
    s0 = x % r0; x /= r0; s1 = x % r1; x /= r1; s2 = x % r2; x /= r2; s3 = x % r3; s4 = x / r3;
  }
  __device__ __forceinline__ bool valid() const { return t < n_tiles; }
  __device__ __forceinline__ TileCoord coord() const { TileCoord c; c.b = b; c.d = d; c.h0 = hi * TH; c.w0 = wi * TW; c.ntile = ntile; return c; }
  __device__ __forceinline__ void next() {
    t += stride;
    int c;
    ntile += s0; c = ntile >= r0; if (c) ntile -= r0;
    wi += s1 + c; c = wi >= r1; if (c) wi -= r1;
    hi += s2 + c; c = hi >= r2; if (c) hi -= r2;
    d += s3 + c; c = d >= r3; if (c) d -= r3;
    b += s4 + c;
  }
};

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvArgs& a = p.a;
  // canonical warp index: the shuffle makes it provably warp-uniform, so the role branches below are uniform
  // branches and the MMA warp's loop compiles to the uniform datapath (UIADD3 + UTCHMMA, no R2UR per operand)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int taps_hw = a.kh * a.kw;
  const int pd = a.kd / 2, ph = a.kh / 2, pw = a.kw / 2;

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

  float2* s_norm = reinterpret_cast<float2*>(smem + p.smem_norm_off);   // [B][Cin] {mean, rstd}
  float2* s_gnorm = reinterpret_cast<float2*>(smem + p.smem_gnorm_off); // [B][Cout] {mean, rstd} of dgrad_x
  float* s_stat = reinterpret_cast<float*>(smem + p.smem_stat_off);     // [4 warps][B][Cout][2]
  const uint32_t smem_a = smem_u32(smem + p.smem_a_off);
  const uint32_t smem_b = smem_u32(smem + p.smem_b_off);

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(A_FULL(i), kLoadGroupThreads); mbar_init(A_EMPTY(i), 1); }
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
    const int nstat = kEpiWarps * a.B * a.Cout * 2;
    for (int i = threadIdx.x; i < nstat; i += kThreads) s_stat[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int ksteps = p.KC / 16;
  const bool dbg = p.debug && blockIdx.x == 0;

  if (warp >= kLoadWarp0 && warp < kWgtWarp) {
    // =========================== A LOADERS ===========================
    // Compact code on purpose: 14 warps run 4 different programs on this SM, and a loader body unrolled over its
    // 12 chunks (~25 KB of SASS) evicted the MMA warp's tiny issue loop from the instruction cache on every
    // iteration (measured ~300 cycles per tcgen05.mma instead of ~50).  Each thread owns ONE 8-channel plane
    // (its InstanceNorm scale/shift live in registers for the whole stage) and walks the halo voxels in steps.
    const int grp = (warp - kLoadWarp0) >> 2;
    const int lt = threadIdx.x - (kLoadWarp0 * 32 + grp * kLoadGroupThreads);
    const int cpv = p.KC / 8;
    const int act_thr = (kLoadGroupThreads / cpv) * cpv;
    const int vstep = kLoadGroupThreads / cpv;
    const int c8 = lt % cpv, v0 = lt / cpv;
    const int sh = vstep / p.HALO_W, sw = vstep % p.HALO_W;
    const bool xform = (a.x_stats != nullptr) || (a.act != 0);
    const __half* xbase = reinterpret_cast<const __half*>(a.x);
    Ring ring; ring.init(p.SA);
    int stage_no = 0, tr_l = 0;
    TileIter ti; ti.init(p);
    for (; ti.valid(); ti.next()) {
      const TileCoord tc = ti.coord();
      for (int kc = 0; kc < p.NKC; ++kc) {
        for (int zd = 0; zd < a.kd; ++zd) {
          const int din = tc.d + zd - pd;
          if ((unsigned)din >= (unsigned)a.D) continue;
          if ((stage_no & 1) == grp) {
            const bool dl = dbg && grp == 0 && lt == 0;
            long long q0 = 0, q1 = 0;
            if (dl) q0 = clock64();
            if (lt == 0) TRACE(1 + grp, tr_l++);                   // stage start (before A_EMPTY wait)
            mbar_wait(A_EMPTY(ring.idx), ring.phase ^ 1, p.err_flag, 1);
            if (lt == 0) TRACE(1 + grp, tr_l++);                   // slot free
            if (dl) q1 = clock64();
            if (lt < act_thr && !(p.debug & 2)) {
              uint8_t* sdst = smem + p.smem_a_off + ring.idx * p.a_stage_bytes + c8 * p.plane_stride;
              const __half* xs = xbase + ((int64_t)(tc.b * a.D + din) * a.H * a.W) * a.x_ld + a.x_coff + kc * p.KC + c8 * 8;
              float sc[8], sf[8];          // x*sc + sf == (x - mean) * rstd
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 mr = s_norm[tc.b * a.Cin + kc * p.KC + c8 * 8 + j];
                sc[j] = mr.y; sf[j] = -mr.x * mr.y;
              }
              int hh = v0 / p.HALO_W, ww = v0 % p.HALO_W;
#pragma unroll 1
              for (int v = v0; v < p.nvox_h; v += 6 * vstep) {
                uint4 raw[6]; int vv[6]; bool ok[6];
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                  vv[u] = hh * p.HALO_W + ww;
                  const int h = tc.h0 - ph + hh, w = tc.w0 - pw + ww;
                  ok[u] = (vv[u] < p.nvox_h) && ((unsigned)h < (unsigned)a.H) && ((unsigned)w < (unsigned)a.W);
                  raw[u] = make_uint4(0, 0, 0, 0);
                  if (ok[u]) raw[u] = __ldg(reinterpret_cast<const uint4*>(xs + ((int64_t)h * a.W + w) * a.x_ld));
                  hh += sh; ww += sw;
                  if (ww >= p.HALO_W) { ww -= p.HALO_W; ++hh; }
                }
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                  if (vv[u] >= p.nvox_h) continue;
                  uint4 o = raw[u];
                  if (ok[u] && xform) {
                    const __half2* hv = reinterpret_cast<const __half2*>(&raw[u]);
                    __half2* ov = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      float2 f = __half22float2(hv[j]);
                      f.x = fmaf(f.x, sc[2 * j], sf[2 * j]); f.y = fmaf(f.y, sc[2 * j + 1], sf[2 * j + 1]);
                      if (a.act == B200SEG_ACT_RELU) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
                      ov[j] = __floats2half2_rn(f.x, f.y);
                    }
                  }
                  *reinterpret_cast<uint4*>(sdst + vv[u] * 16) = o;
                }
              }
            }
            long long q3 = 0;
            if (dl) q3 = clock64();
            if (lt == 0) TRACE(1 + grp, tr_l++);                   // stored
            fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(A_FULL(ring.idx));
            if (lt == 0) TRACE(1 + grp, tr_l++);                   // published
            if (dl) { DBG_ADD(0, q1 - q0); DBG_ADD(2, q3 - q1); DBG_ADD(3, clock64() - q3); DBG_ADD(4, 1); }
          }
          ring.advance();
          ++stage_no;
        }
      }
    }
  } else if (warp == kWgtWarp) {
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
        TileIter ti; ti.init(p);
        for (; ti.valid(); ti.next()) {
          const TileCoord tc = ti.coord();
          for (int kc = 0; kc < p.NKC; ++kc) {
            for (int zd = 0; zd < a.kd; ++zd) {
              const int din = tc.d + zd - pd;
              if ((unsigned)din >= (unsigned)a.D) continue;
              for (int thw = 0; thw < taps_hw; ++thw) {
                const int tap = zd * taps_hw + thw;
                mbar_wait(B_EMPTY(ring.idx), ring.phase ^ 1, p.err_flag, 2);
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
    // elect / R2UR.BROADCAST sequence per operand (measured: 134 -> see profiles/ cycles per MMA), and the loop
    // bounds are copied to locals so no constant-bank load sits on the issue path.
    {
      const uint32_t elected = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      Ring ra, rb; ra.init(p.SA); rb.init(p.SB);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_tmpl = make_desc(0, (uint32_t)p.plane_stride, (uint32_t)p.HALO_W * 16u);
      const uint64_t b_tmpl = make_desc(0, (uint32_t)p.NT * 16u, 128u);
      const uint32_t a_kstep = (2u * (uint32_t)p.plane_stride) >> 4, b_kstep = (2u * (uint32_t)p.NT * 16u) >> 4;
      const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4, b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
      const uint32_t smem_a16 = smem_a >> 4, smem_b16 = smem_b >> 4;
      const int kd = a.kd, kh = a.kh, kw = a.kw, NKC = p.NKC, D = a.D, NT = p.NT, HALO_W = p.HALO_W;
      const int resident = p.w_resident, n_tiles = p.n_tiles, acc_stages = p.acc_stages;
      const int taps_all = kd * taps_hw;
      const uint32_t res_step = (uint32_t)NKC * b_stage16;
      int it = 0, tr_m = 0;
      if (resident) { mbar_wait_nocall(B_FULL(0), 0); tc_fence_after(); }
      TileIter ti; ti.init(p);
      for (; ti.valid(); ti.next(), ++it) {
        const TileCoord tc = ti.coord();
        const int acc = (acc_stages == 2) ? (it & 1) : 0;
        const uint32_t acc_phase = (acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
        long long m0 = 0;
        if (dbg && lane == 0) m0 = clock64();
        mbar_wait_nocall(T_EMPTY(acc), acc_phase ^ 1);
        if (dbg && lane == 0) DBG_ADD(8, clock64() - m0);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + (uint32_t)(acc * NT);
        uint32_t accumulate = 0;
        for (int kc = 0; kc < NKC; ++kc) {
          for (int zd = 0; zd < kd; ++zd) {
            const int din = tc.d + zd - pd;
            if ((unsigned)din >= (unsigned)D) continue;
            if (dbg && lane == 0) m0 = clock64();
            if (lane == 0) TRACE(0, tr_m++);                       // stage start (before A_FULL wait)
            mbar_wait_nocall(A_FULL(ra.idx), ra.phase);
            if (lane == 0) TRACE(0, tr_m++);                       // A ready
            if (dbg && lane == 0) { DBG_ADD(9, clock64() - m0); DBG_ADD(12, 1); m0 = clock64(); }
            tc_fence_after();
            uint64_t da_row = a_tmpl + (uint64_t)(smem_a16 + (uint32_t)ra.idx * a_stage16);
            uint64_t db_res = b_tmpl + (uint64_t)(smem_b16 + (uint32_t)((tc.ntile * taps_all + zd * taps_hw) * NKC + kc) * b_stage16);
            bool done = false;
            if (resident && kh == 3 && kw == 3) {
              done = true;
              switch (ksteps) {
                case 1: issue_stage_resident<1, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, (uint32_t)HALO_W, res_step, elected); break;
                case 2: issue_stage_resident<2, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, (uint32_t)HALO_W, res_step, elected); break;
                case 3: issue_stage_resident<3, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, (uint32_t)HALO_W, res_step, elected); break;
                case 4: issue_stage_resident<4, 3, 3>(tmem_d, da_row, db_res, idesc, accumulate, a_kstep, b_kstep, (uint32_t)HALO_W, res_step, elected); break;
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
                  mbar_wait_nocall(B_FULL(rb.idx), rb.phase);
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
                da_tap += 1;                      // next tap to the right: 16 B
              }
              da_row += (uint64_t)HALO_W;         // next tap row: HALO_W voxels
            }
            }
            if (lane == 0) TRACE(0, tr_m++);                       // MMAs issued
            if (elected) umma_commit(A_EMPTY(ra.idx));         // halo tile free
            ra.advance();
            if (lane == 0) TRACE(0, tr_m++);                       // committed
            if (dbg && lane == 0) DBG_ADD(11, clock64() - m0);      // issue time of one stage (incl. weight waits)
          }
        }
        if (elected) umma_commit(T_FULL(acc));                 // accumulator complete -> epilogue
      }
    }
  } else {
    // =========================== EPILOGUE (warps 0-3) ===========================
    const int q = warp;                         // TMEM lane quadrant
    const int row = q * 32 + lane;              // GEMM row = hl*8 + wl
    const int hl = row >> 3, wl = row & 7;
    const bool dgrad = a.gx != nullptr;
    int it = 0;
    TileIter ti; ti.init(p);
    for (; ti.valid(); ti.next(), ++it) {
      const TileCoord tc = ti.coord();
      const int acc = (p.acc_stages == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
      const int h = tc.h0 + hl, w = tc.w0 + wl;
      const bool valid = (h < a.H) && (w < a.W);
      const int64_t vox = ((int64_t)(tc.b * a.D + tc.d) * a.H + h) * a.W + w;
      const int co_base = tc.ntile * p.NT;
      __half* yp = reinterpret_cast<__half*>(a.y) + vox * a.y_ld + a.y_coff + co_base;
      const __half* rp = a.res ? reinterpret_cast<const __half*>(a.res) + vox * a.r_ld + a.r_coff + co_base : nullptr;
      const __half* gp = dgrad ? reinterpret_cast<const __half*>(a.gx) + vox * a.gx_ld + a.gx_coff + co_base : nullptr;
      float* wstat = s_stat + ((q * a.B + tc.b) * a.Cout + co_base) * 2;
      // side input of the epilogue (residual, or x for the dgrad ReLU mask): fetched one 16-channel chunk AHEAD so
      // its global-memory latency overlaps the wait for the accumulator and the previous chunk's work
      const __half* side = dgrad ? gp : rp;
      uint4 nx0 = make_uint4(0, 0, 0, 0), nx1 = make_uint4(0, 0, 0, 0);
      if (side && valid) {
        nx0 = __ldg(reinterpret_cast<const uint4*>(side));
        nx1 = __ldg(reinterpret_cast<const uint4*>(side + 8));
      }
      long long e0 = 0;
      const bool de = dbg && threadIdx.x == 0;
      if (de) e0 = clock64();
      if (threadIdx.x == 0) TRACE(3, 3 * it);                      // tile start (before T_FULL wait)
      mbar_wait(T_FULL(acc), acc_phase, p.err_flag, 6);
      if (threadIdx.x == 0) TRACE(3, 3 * it + 1);                  // accumulator ready
      if (de) { DBG_ADD(16, clock64() - e0); e0 = clock64(); }
      tc_fence_after();
      for (int n0 = 0; n0 < ((p.debug & 4) ? 0 : p.NT); n0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT + n0), v);
        const uint4 cur0 = nx0, cur1 = nx1;
        if (side && valid && n0 + 16 < p.NT) {
          nx0 = __ldg(reinterpret_cast<const uint4*>(side + n0 + 16));
          nx1 = __ldg(reinterpret_cast<const uint4*>(side + n0 + 24));
        }
        tmem_ld_wait();
        float r[16], s2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(v[j]);
        if (valid) {
          if (a.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] += a.bias[co_base + n0 + j];
          }
          float sv[16];
          if (side) {
            const __half2* h0 = reinterpret_cast<const __half2*>(&cur0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&cur1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
              sv[2 * j] = f0.x; sv[2 * j + 1] = f0.y; sv[8 + 2 * j] = f1.x; sv[8 + 2 * j + 1] = f1.y;
            }
          }
          if (dgrad) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 mr = s_gnorm[tc.b * a.Cout + co_base + n0 + j];
              const float hx = (sv[j] - mr.x) * mr.y;
              float g = (a.g_act == B200SEG_ACT_RELU && !(hx > 0.f)) ? 0.f : r[j];
              g = __half2float(__float2half_rn(g));
              r[j] = g; s2[j] = g * hx;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __half2float(__float2half_rn(r[j]));
            if (rp) {
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = __half2float(__float2half_rn(r[j] + sv[j]));
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) s2[j] = r[j] * r[j];
          }
          st8<__half>(yp + n0, reinterpret_cast<const float(&)[8]>(r[0]));
          st8<__half>(yp + n0 + 8, reinterpret_cast<const float(&)[8]>(r[8]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { r[j] = 0.f; s2[j] = 0.f; }
        }
        if (a.y_stats) {
          // column sums over the warp's 32 rows with a halving butterfly: 16 shuffles per quantity instead of 80.
          // Afterwards lanes 2c and 2c+1 both hold the sum of column col(lane).
          const float u = column_sum16(r, lane), q2 = column_sum16(s2, lane);
          if ((lane & 1) == 0) {
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            wstat[(n0 + col) * 2] += u;
            wstat[(n0 + col) * 2 + 1] += q2;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(T_EMPTY(acc));
      if (threadIdx.x == 0) TRACE(3, 3 * it + 2);                  // drained
      if (de) { DBG_ADD(17, clock64() - e0); DBG_ADD(18, 1); }
    }
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
      for (int q = 0; q < kEpiWarps; ++q) s += (double)s_stat[q * n + i];
      if (s != 0.0) atomicAdd(&a.y_stats[i], s);
    }
  }
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace

bool conv3d_tc_shape_ok(int Cin, int Cout, int kd, int kh, int kw, int dtype) {
  if (dtype != B200SEG_F16) return false;
  if (tc_pick_nt(Cout) == 0 || tc_pick_kc(Cin) == 0) return false;
  if (kh > 3 || kw > 3 || kd > 3 || kd < 1 || kh < 1 || kw < 1) return false;
  return true;
}

bool conv3d_fwd_tc_supported(const ConvArgs& a, int dtype) {
  if (!conv3d_tc_shape_ok(a.Cin, a.Cout, a.kd, a.kh, a.kw, dtype)) return false;
  if (dtype != B200SEG_F16) return false;
  if (tc_pick_nt(a.Cout) == 0 || tc_pick_kc(a.Cin) == 0) return false;
  if (a.kh > 3 || a.kw > 3 || a.kd > 3) return false;
  if ((a.x_ld % 8) || (a.x_coff % 8) || (a.y_ld % 8) || (a.y_coff % 8)) return false;
  if (a.res && ((a.r_ld % 8) || (a.r_coff % 8))) return false;
  if (a.gx && ((a.gx_ld % 8) || (a.gx_coff % 8))) return false;
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.y) | reinterpret_cast<uintptr_t>(a.w)) & 15) return false;
  if (a.B * a.Cin > 4096 || a.B * a.Cout > 2048) return false;
  return true;
}

// `a.w` must be the TC weight IMAGE ([ntile][tap][kchunk][KC/8][NT][8], b200seg_pack_weight layout=TC).
int conv3d_fwd_tc(const ConvArgs& a, int dtype, cudaStream_t st) {
  if (!conv3d_fwd_tc_supported(a, dtype)) return B200SEG_EUNSUPPORTED;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.a = a;
  p.wimg = a.w;
  p.KC = tc_pick_kc(a.Cin); p.NKC = a.Cin / p.KC;
  p.NT = tc_pick_nt(a.Cout); p.NTILES = a.Cout / p.NT;
  p.HALO_H = TH + a.kh - 1; p.HALO_W = TW + a.kw - 1;
  p.nvox_h = p.HALO_H * p.HALO_W;
  int slots = p.nvox_h; if ((slots & 1) == 0) slots += 1;     // odd number of 16-B slots -> conflict-free plane stride
  p.plane_stride = slots * 16;
  p.a_stage_bytes = (p.KC / 8) * p.plane_stride;
  p.a_stage_bytes = (p.a_stage_bytes + 127) / 128 * 128;
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
  const int stat_bytes = kEpiWarps * a.B * a.Cout * 2 * 4;
  const int budget = 227 * 1024 - 1024 - norm_bytes - stat_bytes - a.B * a.Cout * 8 - 512;
  const int64_t w_total = (int64_t)a.kd * a.kh * a.kw * a.Cin * a.Cout * 2;
  int b_region;
  if (w_total <= 112 * 1024 && w_total + 2 * p.a_stage_bytes <= budget) {
    p.w_resident = 1; p.SB = 1;
    b_region = (int)w_total;
    p.SA = (budget - b_region) / p.a_stage_bytes; if (p.SA > 4) p.SA = 4;
  } else {
    p.w_resident = 0;
    p.SA = 4;
    while (p.SA > 2 && p.SA * p.a_stage_bytes + 3 * p.b_stage_bytes > budget) --p.SA;
    p.SB = (budget - p.SA * p.a_stage_bytes) / p.b_stage_bytes; if (p.SB > 8) p.SB = 8;
    if (p.SB < 2) return B200SEG_EUNSUPPORTED;
    b_region = p.SB * p.b_stage_bytes;
  }
  int off = 0;
  p.smem_a_off = off; off += p.SA * p.a_stage_bytes;
  off = (off + 127) / 128 * 128;
  p.smem_b_off = off; off += b_region;
  off = (off + 15) / 16 * 16;
  p.smem_bar_off = off; off += (2 * p.SA + 2 * p.SB + 4) * 8 + 16;
  off = (off + 15) / 16 * 16;
  p.smem_norm_off = off; off += norm_bytes;
  off = (off + 15) / 16 * 16;
  p.smem_gnorm_off = off; off += a.B * a.Cout * 8;
  off = (off + 15) / 16 * 16;
  p.smem_stat_off = off; off += stat_bytes;
  const int smem_bytes = off + 1024;       // slack for the 1024-B alignment of the dynamic segment
  p.err_flag = nullptr;
  { const char* dv = getenv("B200SEG_TC_DEBUG"); p.debug = dv ? atoi(dv) : 0; }
  int grid = p.n_tiles < B200SEG_NUM_SMS ? p.n_tiles : B200SEG_NUM_SMS;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  conv_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  B200_CHECK_LAUNCH("conv_tc_kernel");
  return B200SEG_OK;
}


// debug: read-and-clear the cycle counters written by block 0 when B200SEG_TC_DEBUG=1
extern "C" int b200seg_debug_tc_timers(long long* out32) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return b200seg_record_cuda(e, "sync");
  e = cudaMemcpyFromSymbol(out32, g_tc_dbg, sizeof(long long) * 32);
  if (e != cudaSuccess) return b200seg_record_cuda(e, "memcpyFromSymbol");
  long long z[32] = {0};
  e = cudaMemcpyToSymbol(g_tc_dbg, z, sizeof(z));
  if (e != cudaSuccess) return b200seg_record_cuda(e, "memcpyToSymbol");
  return B200SEG_OK;
}

extern "C" int b200seg_debug_tc_trace(long long* out1024) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return b200seg_record_cuda(e, "sync");
  e = cudaMemcpyFromSymbol(out1024, g_tc_trace, sizeof(long long) * 1024);
  if (e != cudaSuccess) return b200seg_record_cuda(e, "memcpyFromSymbol");
  return B200SEG_OK;
}
