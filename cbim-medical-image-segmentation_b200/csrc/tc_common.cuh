// tc_common.cuh — PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, wgrad_tc.cu, window_attn.cu): mbarrier,
// bulk-TMA, cp.async, TMEM allocation / loads, tcgen05.mma issue + commit, SWIZZLE_NONE matrix descriptors.
// sm_100a only; every wrapper is a thin inline-asm statement so the SASS shows UTCHMMA / LDTM / UBLKCP / SYNCS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace tc {

constexpr long long kWaitTimeoutCycles = 4000000000ll;   // ~2 s at 1.9 GHz: a protocol bug traps, it never hangs the GPU
constexpr uint32_t kSuspendNs = 20000;         // suspend-time hint of a blocked try_wait (the thread is woken on completion)
// per-TU copy of the hint the kernels actually pass (tuning knob: env B200SEG_SUSPEND_NS, read once by tc_apply_env())
__constant__ uint32_t c_suspend_ns = kSuspendNs;

// ---- role profiler (tools/tc_prof.py; compiled only into libb200seg_prof.so with -DB200SEG_TC_PROFILE) ----
// Lane 0 of every warp adds the cycles it spends inside a TC_PROF(code) scope (and inside every mbar_wait, keyed by
// the wait site's code) to a per-CTA row; slot 31 = kernel lifetime of thread 0, slots 32+code = scope entry counts.
#ifdef B200SEG_TC_PROFILE
constexpr int kProfSlots = 64, kProfRows = 160;
__device__ unsigned long long g_tc_prof[kProfRows * kProfSlots];
struct ProfScope {
  int code; long long t0;
  __device__ __forceinline__ explicit ProfScope(int c) { code = c; t0 = clock64(); }
  __device__ __forceinline__ ~ProfScope() {
    if ((threadIdx.x & 31) == 0 && blockIdx.x < kProfRows) {
      atomicAdd(&g_tc_prof[blockIdx.x * kProfSlots + (code & 31)], (unsigned long long)(clock64() - t0));
      atomicAdd(&g_tc_prof[blockIdx.x * kProfSlots + 32 + (code & 31)], 1ull);
    }
  }
};
#define TC_PROF(code) ProfScope prof_scope_##code(code)
#define TC_PROF_ENTRY(name) \
  extern "C" int name(unsigned long long* out, int reset) { \
    static unsigned long long h[kProfRows * kProfSlots]; \
    if (cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h)) != cudaSuccess) return -1; \
    for (int i = 0; i < kProfSlots; ++i) { out[i] = 0; for (int r = 0; r < kProfRows; ++r) out[i] += h[r * kProfSlots + i]; } \
    if (reset) { memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g_tc_prof, h, sizeof(h)); } \
    return 0; }
#else
#define TC_PROF(code)
#define TC_PROF_ENTRY(name)
#endif

inline void tc_apply_env() {
  static bool done = false;
  if (done) return;
  done = true;
  if (const char* e = getenv("B200SEG_SUSPEND_NS")) { const uint32_t v = (uint32_t)atoi(e); cudaMemcpyToSymbol(c_suspend_ns, &v, sizeof(v)); }
}

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(c_suspend_ns) : "memory");
  return ok != 0;
}
// non-blocking probe (no suspend): has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.  No out-of-line
// diagnostics: a CALL anywhere in a kernel that uses setmaxnreg makes ptxas keep every role inside the SMALLEST
// register allotment (measured: the epilogue stayed below R88 and spilled its accumulators), and a call inside the
// MMA issue loop pushes the loop off the uniform datapath.  `code` only documents the wait site.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code) {
  (void)code;
  TC_PROF(code);
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWaitTimeoutCycles) __trap();
  }
}
// (kept as a separate name for the MMA warp's waits: with a call on the slow path the compiler keeps the loop-carried
// descriptors in vector registers and every tcgen05.mma needs five R2UR.BROADCASTs, ~250 cycles per MMA measured)
__device__ __forceinline__ void mbar_wait_nocall(uint32_t bar, uint32_t parity, int code = 0) { mbar_wait(bar, parity, code); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk TMA global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 5-D tensor TMA load (SASS: UTMALDG): one box of the tensor map into shared memory, zero-filling out-of-bound
// coordinates (negative ones included), completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 16-byte cp.async (SASS: LDGSTS); src_bytes == 0 zero-fills the destination (out-of-volume voxels)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 x fp16 -> fp32  (SASS: UTCHMMA)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand is read from tensor memory (no shared-memory read of A)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// TMEM -> registers, 32 lanes x 16 consecutive 32-bit columns per warp (SASS: LDTM)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, 32 lanes x 8 consecutive 32-bit columns per warp (SASS: STTM)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, version 1 = sm_100).
//   K-major operand : lbo = stride between core matrices along K, sbo = stride between 8-row groups along M/N
//   MN-major operand: lbo = stride between core matrices along K (voxel groups), sbo = stride along M/N (channel planes)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
  return d;                    // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// swizzled operand rows (one 64- / 128-byte row per GEMM row resp. K index, as a tensor-TMA box with the matching swizzle
// writes them): layout_type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B; lbo = distance between swizzle-wide blocks along the
// leading (contiguous) dimension, sbo = distance between 8-row groups.  The swizzle is a function of the ABSOLUTE
// shared-memory address bits, so a start address shifted by whole rows (a convolution tap) and an 8-row-group pitch
// of HALO_W rows are legal (tools/umma_probe.cu, profiles/r1_umma_layout_probe.txt).
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return make_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)(layout_type & 7u) << 61);
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}

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
__device__ __forceinline__ int column_of_lane16(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// Register re-distribution between warpgroups (4 consecutive warps execute it together).  The kernel is launched with
// 65536 / blockDim registers per thread; role code then grows / shrinks its warpgroup's allotment.
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

struct Ring {
  int idx; uint32_t phase; int n;
  __device__ __forceinline__ void init(int n_) { idx = 0; phase = 0; n = n_; }
  __device__ __forceinline__ void advance() { if (++idx == n) { idx = 0; phase ^= 1; } }
};

}  // namespace tc
