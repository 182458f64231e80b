// wgrad_tc.cu — conv3d weight gradient as tcgen05 GEMMs (sm_100a), fp16 operands, fp32 accumulation in TMEM.
// Replaces cuDNN wgrad behind autograd of nn.Conv3d (train_ddp.py:193/208); what it writes is the fp32
// parameter-gradient tensor DDP all-reduces.
//
//   dW[co][ci][tap] = sum_voxels dy[v][co] * a[v + tap][ci],      a = act(IN(x))  (materialised once by
//   instnorm_apply into the workspace, so both operands are loaded raw).
//
// GEMM view per CTA: D_tap[128 co][N ci] += dy^T[128 co x 128 voxels] * a_tap[128 voxels x N ci] for a GROUP of
// in-plane taps of one depth offset zd; K = voxels, accumulated over every voxel tile the CTA owns, so the
// accumulators stay in TMEM for the CTA's whole life (G*N <= 512 columns) and there is ONE epilogue.
//   * both operands are staged as [channel/8][voxel][8 ch] — the same image conv_tc.cu uses — and read by
//     the tensor core as MN-major no-swizzle matrices (core matrix = 8 voxels x 8 channels, 128 B):
//     A = dy tile (16x8 voxels), B = halo tile of `a` ((16+kh-1)x(8+kw-1) voxels); a tap is a shifted
//     B descriptor, exactly like the forward kernel.
//   * split-K over voxel tiles fills the machine: grid = jobs x S; partial D tiles go to a workspace with
//     plain coalesced stores and a small second kernel reduces them into dW (+=) — no atomics.
// Warp roles (416 threads, 1 CTA/SM): warps 0-3 epilogue, warps 4-11 loaders, warp 12 MMA issue + TMEM alloc
// (was: loaders (cp.async, all 256 threads per stage, stage k published while stage k+1 is in flight).
#include "common.cuh"
#include "conv_args.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int TH = 16, TW = 8;
constexpr int kEpiWarps = 4;
constexpr int kLoadWarp0 = 4;
constexpr int kMmaWarp = 12;                // highest warp id = highest issue priority in its SM sub-partition
constexpr int kLoadGroups = 1;             // all loader warps cooperate on every stage (deferred publication needs
constexpr int kLoadGroupThreads = 256;     // consecutive stages from the same threads; works for a 2-slot ring)
constexpr int kThreads = 13 * 32;   // 416
constexpr uint32_t kSpinLimit = 1u << 24;
constexpr int MT = 128;                    // output-channel tile (GEMM M)

struct WgParams {
  const __half* a; int a_ld, a_coff;       // normalised+activated input (or raw x when no norm/act)
  const __half* dy; int dy_ld, dy_coff;
  float* partial;                          // [job][S][128][Gmax*NTC]
  int B, D, H, W, Cin, Cout, kd, kh, kw;
  int NTC, ci_tiles, co_tiles, G, ngroups, gbase, grem, S;
  int HALO_H, HALO_W, nvox_h, a_plane, dy_plane, a_bytes, dy_bytes, stage_bytes, NS;
  int tiles_h, tiles_w, nvt;
  int tmem_cols;
  int smem_bar_off;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __noinline__ void mbar_timeout(int code, uint32_t parity) {
  printf("b200seg wgrad_tc: mbarrier timeout code=%d block=%d thread=%d parity=%u\n", code, blockIdx.x, threadIdx.x, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) mbar_timeout(code, parity);
  }
}
__device__ __forceinline__ void mbar_wait_nocall(uint32_t bar, uint32_t parity) {   // see conv_tc.cu
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// SWIZZLE_NONE matrix descriptor.  MN-major operands: lbo = stride between core matrices along K (voxels),
// sbo = stride between core matrices along M/N (channel planes)  (cute::UMMA::make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}

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

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // canonical warp index: the shuffle makes it provably warp-uniform, so the role branches below are uniform
  // branches and the MMA warp's loop compiles to the uniform datapath (UIADD3 + UTCHMMA, no R2UR per operand)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const Job job = decode_job(p, blockIdx.x);
  const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;
  const int co0 = job.co_tile * MT;
  const int co_real = min(MT, p.Cout - co0);
  const int ci0 = job.ci_tile * p.NTC;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(p.NS + i); };
  const uint32_t DONE = bar0 + 8u * (uint32_t)(2 * p.NS);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(bars + 2 * p.NS + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NS; ++i) { mbar_init(FULL(i), kLoadGroupThreads); mbar_init(EMPTY(i), 1); }
    mbar_init(DONE, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), (uint32_t)p.tmem_cols);
  // zero the dy planes this job never writes (co tile narrower than 128): they are the M padding
  {
    const int planes_real = co_real / 8;
    if (planes_real < 16) {
      for (int s = 0; s < p.NS; ++s) {
        uint4* base = reinterpret_cast<uint4*>(smem + s * p.stage_bytes + planes_real * p.dy_plane);
        const int n16 = (16 - planes_real) * p.dy_plane / 16;
        for (int i = threadIdx.x; i < n16; i += kThreads) base[i] = make_uint4(0, 0, 0, 0);
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= kLoadWarp0 && warp < kMmaWarp) {
    // =========================== LOADERS ===========================
    const int lt = threadIdx.x - kLoadWarp0 * 32;
    // dy tile: cpv planes, thread owns plane (lt % cpv) and walks voxels v0, v0+vstep, ...
    const int cpv_d = co_real / 8;
    const int act_d = (kLoadGroupThreads / cpv_d) * cpv_d;
    const int vstep_d = kLoadGroupThreads / cpv_d;
    const int c8_d = lt % cpv_d, v0_d = lt / cpv_d;
    const int cpv_a = p.NTC / 8;
    const int act_a = (kLoadGroupThreads / cpv_a) * cpv_a;
    const int vstep_a = kLoadGroupThreads / cpv_a;
    const int c8_a = lt % cpv_a, v0_a = lt / cpv_a;
    const int sh_a = vstep_a / p.HALO_W, sw_a = vstep_a % p.HALO_W;
    // Both operands are raw fp16, so they are staged with cp.async (LDGSTS): no register round trip, every
    // 16-B chunk of a stage is in flight at once, out-of-volume voxels are zero-filled by src-size 0.  A stage is
    // published one stage late (wait_group 1 -> fence.proxy.async -> mbarrier arrive), so the copies of stage
    // k+1 overlap the completion of stage k.
    int idx = 0; uint32_t phase = 0; int stage_no = 0;
    int pending_slot = -1;
    for (int vt = job.s; vt < p.nvt; vt += p.S) {
      int t = vt;
      const int w0 = (t % p.tiles_w) * TW; t /= p.tiles_w;
      const int h0 = (t % p.tiles_h) * TH; t /= p.tiles_h;
      const int d = t % p.D; const int b = t / p.D;
      const int din = d + job.zd - pd;
      if ((unsigned)din >= (unsigned)p.D) continue;
      {
        mbar_wait(EMPTY(idx), phase ^ 1, 1);
        const uint32_t sdy = smem_u32(smem + idx * p.stage_bytes);
        const uint32_t sa = sdy + (uint32_t)p.dy_bytes;
        if (lt < act_d) {
          const __half* src = p.dy + ((int64_t)(b * p.D + d) * p.H * p.W) * p.dy_ld + p.dy_coff + co0 + c8_d * 8;
          const uint32_t dst = sdy + (uint32_t)(c8_d * p.dy_plane);
          for (int v = v0_d; v < TH * TW; v += vstep_d) {
            const int h = h0 + (v >> 3), w = w0 + (v & 7);
            const bool ok = h < p.H && w < p.W;
            cp_async16(dst + (uint32_t)v * 16u, ok ? src + ((int64_t)h * p.W + w) * p.dy_ld : src, ok ? 16u : 0u);
          }
        }
        if (lt < act_a) {
          const __half* src = p.a + ((int64_t)(b * p.D + din) * p.H * p.W) * p.a_ld + p.a_coff + ci0 + c8_a * 8;
          const uint32_t dst = sa + (uint32_t)(c8_a * p.a_plane);
          int hh = v0_a / p.HALO_W, ww = v0_a % p.HALO_W;
          for (int v = v0_a; v < p.nvox_h; v += vstep_a) {
            const int h = h0 - ph + hh, w = w0 - pw + ww;
            const bool ok = (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
            cp_async16(dst + (uint32_t)v * 16u, ok ? src + ((int64_t)h * p.W + w) * p.a_ld : src, ok ? 16u : 0u);
            hh += sh_a; ww += sw_a;
            if (ww >= p.HALO_W) { ww -= p.HALO_W; ++hh; }
          }
        }
        cp_async_commit();
        if (pending_slot >= 0) {
          cp_async_wait<1>();
          fence_proxy_async();
          mbar_arrive(FULL(pending_slot));
        }
        pending_slot = idx;
      }
      if (++idx == p.NS) { idx = 0; phase ^= 1; }
      ++stage_no;
    }
    if (pending_slot >= 0) {
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(FULL(pending_slot));
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER ===========================
    {   // whole warp, warp-uniform values, one elected lane issues (see conv_tc.cu)
      const uint32_t elected = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.NTC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t dy_lbo = 128u, dy_sbo = (uint32_t)p.dy_plane;
      uint32_t a_lbo = (uint32_t)p.HALO_W * 16u, a_sbo = (uint32_t)p.a_plane;
      // lean issue loop (this one thread feeds the tensor core): descriptor templates + constant adds
      const uint64_t dy_tmpl = make_desc(0, dy_lbo, dy_sbo), a_tmpl = make_desc(0, a_lbo, a_sbo);
      const uint32_t a_kstep = (2u * (uint32_t)p.HALO_W * 16u) >> 4;    // two voxel rows per K=16 step
      const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, dy16 = (uint32_t)p.dy_bytes >> 4;
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const int zh0 = job.tap0 / p.kw, zw0 = job.tap0 % p.kw;
      const int dslab = p.tiles_w * p.tiles_h;
      int idx = 0; uint32_t phase = 0; uint32_t accumulate = 0;
      for (int vt = job.s; vt < p.nvt; vt += p.S) {
        const int d = (vt / dslab) % p.D;
        const int din = d + job.zd - pd;
        if ((unsigned)din >= (unsigned)p.D) continue;
        mbar_wait_nocall(FULL(idx), phase);
        tc_fence_after();
        const uint64_t da0 = dy_tmpl + (uint64_t)(smem16 + (uint32_t)idx * stage16);
        uint64_t db_tap = a_tmpl + (uint64_t)(smem16 + (uint32_t)idx * stage16 + dy16 + (uint32_t)(zh0 * p.HALO_W + zw0));
        int zw = zw0;
        uint32_t tmem_d = tmem_u;
        for (int tl = 0; tl < job.ntaps; ++tl) {
          uint64_t da = da0, db = db_tap;
#pragma unroll
          for (int j = 0; j < (TH * TW) / 16; ++j) {
            if (elected) umma_f16(tmem_d, da, db, idesc, (accumulate | (uint32_t)(j > 0)));
            da += 16;            // 2 voxel rows of the dy tile = 256 B
            db += a_kstep;
          }
          tmem_d += (uint32_t)p.NTC;
          // next in-plane tap: one voxel to the right, or wrap to the next halo row
          if (++zw == p.kw) { zw = 0; db_tap += (uint64_t)(p.HALO_W - (p.kw - 1)); } else db_tap += 1;
        }
        accumulate = 1;
        if (elected) umma_commit(EMPTY(idx));
        if (++idx == p.NS) { idx = 0; phase ^= 1; }
      }
      if (elected) umma_commit(DONE);
    }
  } else if (warp < kEpiWarps) {
    // =========================== EPILOGUE (once) ===========================
    // did this CTA process any stage at all?
    bool any = false;
    for (int vt = job.s; vt < p.nvt; vt += p.S) {
      const int d = (vt / (p.tiles_w * p.tiles_h)) % p.D;
      if ((unsigned)(d + job.zd - pd) < (unsigned)p.D) { any = true; break; }
    }
    mbar_wait(DONE, 0, 3);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int gmax = p.gbase + (p.grem ? 1 : 0);
    float* dst = p.partial + ((int64_t)blockIdx.x * MT + row) * (gmax * p.NTC);
    for (int tl = 0; tl < job.ntaps; ++tl) {
      for (int n0 = 0; n0 < p.NTC; n0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tl * p.NTC + n0), v);
        tmem_ld_wait();
        float4* o = reinterpret_cast<float4*>(dst + tl * p.NTC + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 f = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                 __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          if (!any) f = make_float4(0.f, 0.f, 0.f, 0.f);
          o[q] = f;
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

// dw[co][ci][tap] += sum_s partial[job(co,ci,tap)][s][co%128][tl*NTC + ci%NTC]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, WgParams p) {
  const int taps_hw = p.kh * p.kw, taps = p.kd * taps_hw;
  const int64_t total = (int64_t)p.Cout * p.Cin * taps;
  const int gmax = p.gbase + (p.grem ? 1 : 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps); int64_t t = i / taps; const int ci = (int)(t % p.Cin); const int co = (int)(t / p.Cin);
    const int zd = tap / taps_hw, thw = tap % taps_hw;
    int g, tl;
    if (thw < p.grem * (p.gbase + 1)) { g = thw / (p.gbase + 1); tl = thw % (p.gbase + 1); }
    else { const int r = thw - p.grem * (p.gbase + 1); g = p.grem + r / p.gbase; tl = r % p.gbase; }
    const int co_tile = co / MT, row = co % MT, ci_tile = ci / p.NTC, n = ci % p.NTC;
    const int64_t jobid = (((int64_t)co_tile * p.ci_tiles + ci_tile) * p.kd + zd) * p.ngroups + g;
    const float* src = partial + ((jobid * p.S) * MT + row) * (int64_t)(gmax * p.NTC) + tl * p.NTC + n;
    float s = 0.f;
    for (int k = 0; k < p.S; ++k) s += src[(int64_t)k * MT * (gmax * p.NTC)];
    dw[i] += s;
  }
}

int pick_ntc(int Cin) {
  if (Cin % 16) return 0;
  if (Cin <= 128) return Cin;
  const int c[] = {128, 96, 64, 48, 32, 16};
  for (int v : c) if (Cin % v == 0) return v;
  return 0;
}

bool fill_params(const WgradArgs& a, WgParams& p) {
  memset(&p, 0, sizeof(p));
  p.B = a.B; p.D = a.D; p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout; p.kd = a.kd; p.kh = a.kh; p.kw = a.kw;
  p.NTC = pick_ntc(a.Cin);
  if (!p.NTC || a.Cout % 8) return false;
  p.ci_tiles = a.Cin / p.NTC;
  p.co_tiles = (a.Cout + MT - 1) / MT;
  const int taps_hw = a.kh * a.kw;
  int G = 512 / p.NTC; if (G > taps_hw) G = taps_hw;
  p.G = G;
  p.ngroups = (taps_hw + G - 1) / G;
  p.gbase = taps_hw / p.ngroups; p.grem = taps_hw % p.ngroups;
  p.HALO_H = TH + a.kh - 1; p.HALO_W = TW + a.kw - 1; p.nvox_h = p.HALO_H * p.HALO_W;
  int slots = p.nvox_h; if ((slots & 1) == 0) ++slots;
  p.a_plane = slots * 16;
  p.dy_plane = (TH * TW + 1) * 16;
  p.a_bytes = (p.NTC / 8) * p.a_plane; p.a_bytes = (p.a_bytes + 127) / 128 * 128;
  p.dy_bytes = 16 * p.dy_plane; p.dy_bytes = (p.dy_bytes + 127) / 128 * 128;
  p.stage_bytes = p.a_bytes + p.dy_bytes;
  const int budget = 227 * 1024 - 2048;
  p.NS = budget / p.stage_bytes; if (p.NS > 4) p.NS = 4;
  if (p.NS < 2) return false;
  p.tiles_h = (a.H + TH - 1) / TH; p.tiles_w = (a.W + TW - 1) / TW;
  const int64_t nvt = (int64_t)a.B * a.D * p.tiles_h * p.tiles_w;
  if (nvt > 0x7fffffff) return false;
  p.nvt = (int)nvt;
  const int64_t jobs = (int64_t)p.co_tiles * p.ci_tiles * a.kd * p.ngroups;
  int S = (int)(B200SEG_NUM_SMS / jobs); if (S < 1) S = 1;
  if (S > p.nvt) S = p.nvt;
  p.S = S;
  const int gmax = p.gbase + (p.grem ? 1 : 0);
  int cols = gmax * p.NTC, pow2 = 32; while (pow2 < cols) pow2 <<= 1;
  if (pow2 > 512) return false;
  p.tmem_cols = pow2;
  p.smem_bar_off = p.NS * p.stage_bytes;
  return true;
}

}  // namespace

// element-wise pre-pass: act(IN(x)) -> workspace (instnorm.cu)
extern "C" int b200seg_instnorm_apply(const void* x, int dtype, int x_ld, int x_coff, const double* stats, float eps,
                                      int act, void* y, int y_ld, int y_coff, int B, int64_t V, int C, void* stream);

bool conv3d_wgrad_tc_supported(const WgradArgs& a, int dtype) {
  if (dtype != B200SEG_F16) return false;
  if (a.kd > 3 || a.kh > 3 || a.kw > 3) return false;
  if ((a.x_ld % 8) || (a.x_coff % 8) || (a.dy_ld % 8) || (a.dy_coff % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.dy)) & 15) return false;
  if (a.dbias) return false;
  WgParams p;
  return fill_params(a, p);
}

size_t conv3d_wgrad_tc_workspace(const WgradArgs& a) {
  WgParams p;
  if (!fill_params(a, p)) return 0;
  const int gmax = p.gbase + (p.grem ? 1 : 0);
  const int64_t jobs = (int64_t)p.co_tiles * p.ci_tiles * a.kd * p.ngroups;
  size_t part = (size_t)jobs * p.S * MT * gmax * p.NTC * sizeof(float);
  size_t abuf = (a.x_stats || a.act) ? (size_t)a.B * a.D * a.H * a.W * a.Cin * sizeof(__half) : 0;
  return ((part + 255) / 256) * 256 + ((abuf + 255) / 256) * 256;
}

int conv3d_wgrad_tc(const WgradArgs& a, int dtype, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (!conv3d_wgrad_tc_supported(a, dtype)) return B200SEG_EUNSUPPORTED;
  WgParams p;
  fill_params(a, p);
  if (!workspace || ws_bytes < conv3d_wgrad_tc_workspace(a)) return B200SEG_EINVAL;
  const int gmax = p.gbase + (p.grem ? 1 : 0);
  const int64_t jobs = (int64_t)p.co_tiles * p.ci_tiles * a.kd * p.ngroups;
  const size_t part = ((size_t)jobs * p.S * MT * gmax * p.NTC * sizeof(float) + 255) / 256 * 256;
  p.partial = reinterpret_cast<float*>(workspace);
  p.dy = reinterpret_cast<const __half*>(a.dy); p.dy_ld = a.dy_ld; p.dy_coff = a.dy_coff;
  if (a.x_stats || a.act) {
    __half* abuf = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(workspace) + part);
    int rc = b200seg_instnorm_apply(a.x, B200SEG_F16, a.x_ld, a.x_coff, a.x_stats, a.eps, a.act, abuf, a.Cin, 0,
                                    a.B, (int64_t)a.D * a.H * a.W, a.Cin, (void*)st);
    if (rc) return rc;
    p.a = abuf; p.a_ld = a.Cin; p.a_coff = 0;
  } else {
    p.a = reinterpret_cast<const __half*>(a.x); p.a_ld = a.x_ld; p.a_coff = a.x_coff;
  }
  const int smem_bytes = p.smem_bar_off + (2 * p.NS + 2) * 8 + 64;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = (int)(jobs * p.S);
  wgrad_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  B200_CHECK_LAUNCH("wgrad_tc_kernel");
  const int64_t total = (int64_t)a.Cout * a.Cin * a.kd * a.kh * a.kw;
  int rgrid = ceil_div(total, 256); if (rgrid > B200SEG_NUM_SMS * 16) rgrid = B200SEG_NUM_SMS * 16;
  wgrad_reduce_kernel<<<rgrid, 256, 0, st>>>(p.partial, a.dw, p);
  B200_CHECK_LAUNCH("wgrad_reduce_kernel");
  return B200SEG_OK;
}
