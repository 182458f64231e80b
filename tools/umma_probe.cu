// umma_probe.cu — hardware experiment: tcgen05.mma operand-fetch behaviour for the shared-memory layouts the
// conv kernels can use.  Measures cycles per MMA (clock64 around a burst of MMAs + commit) and checks the result,
// for: no-swizzle "interleaved" K-major operands (what conv_tc.cu v1 uses) vs SWIZZLE_128B rows, including the
// case the implicit-GEMM needs: a SHIFTED start address (tap shift by s voxel rows) with 8-row groups that are
// HALO_W rows apart (SBO not a multiple of 1024 B), with and without the descriptor base_offset.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct Cfg {
  int a_layout;   // 0 none (planes of 16-B voxel slots), 1 SW128 (128-B voxel rows, address-bit swizzle)
  int b_layout;   // 0 none, 1 SW128
  int N;          // 32..256
  int shift;      // tap shift in voxel rows
  int gw;         // voxel rows between consecutive 8-row groups (8 = dense, 10 = halo width)
  int base_mode;  // 0: base_offset 0, 1: base_offset = shift & 7
  int reps;       // MMAs issued = reps * 4 (K = 64)
  int mn_major;   // 0: K-major A (conv fwd), 1: MN-major A and B (wgrad view): A[m=channel][k=voxel]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 5) - 2); }
__device__ __forceinline__ float bval(int n, int k) { return (float)(((n * 5 + k * 11) % 7) - 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

constexpr int A_BYTES = 96 * 1024, B_BYTES = 64 * 1024;

__global__ void __launch_bounds__(128, 1) probe_kernel(Cfg c, float* out, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                    // 1024-aligned
  uint8_t* sB = smem + A_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  volatile uint32_t* tptr = reinterpret_cast<volatile uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = 64;
  for (int i = tid; i < (A_BYTES + B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // ---- fill A: logical A[m][k], m = 0..127 (GEMM row), k = 0..63
  // K-major (conv): GEMM row m <-> voxel slot v = (m/8)*gw + (m%8) + shift ; k <-> channel
  // MN-major (wgrad): GEMM row m <-> channel m (two 64-channel blocks), k <-> voxel slot v = (k/8)*gw + (k%8) + shift
  for (int e = tid; e < 128 * K; e += 128) {
    int m = e / K, k = e % K;
    __half val = __float2half(aval(m, k));
    int v, ch;
    if (!c.mn_major) { v = (m / 8) * c.gw + (m % 8) + c.shift; ch = k; }
    else { v = (k / 8) * c.gw + (k % 8) + c.shift; ch = m; }
    if (c.a_layout == 0) {
      const int plane = 181 * 16;     // odd number of 16-B slots, as conv_tc.cu
      *reinterpret_cast<__half*>(sA + (ch / 8) * plane + v * 16 + (ch % 8) * 2) = val;
    } else {
      const int blk = ch / 64, cc = ch % 64;       // 64-channel blocks are separate buffers 32 KB apart
      uint32_t off = (uint32_t)blk * 32768u + (uint32_t)v * 128u + (uint32_t)(((cc / 8) ^ (v & 7)) * 16) + (cc % 8) * 2;
      *reinterpret_cast<__half*>(sA + off) = val;
    }
  }
  // ---- fill B: logical B[n][k]
  for (int e = tid; e < c.N * K; e += 128) {
    int n = e / K, k = e % K;
    __half val = __float2half(bval(n, k));
    if (!c.mn_major) {
      if (c.b_layout == 0) *reinterpret_cast<__half*>(sB + (k / 8) * (c.N * 16) + n * 16 + (k % 8) * 2) = val;
      else *reinterpret_cast<__half*>(sB + n * 128 + (((k / 8) ^ (n & 7)) * 16) + (k % 8) * 2) = val;
    } else {
      // MN-major B: n = channel (N <= 64 here per block of 128 B), k = voxel slot (dense 8-row groups, no shift)
      int v = k;
      if (c.b_layout == 0) *reinterpret_cast<__half*>(sB + (n / 8) * (129 * 16) + v * 16 + (n % 8) * 2) = val;
      else {
        const int blk = n / 64, cc = n % 64;
        *reinterpret_cast<__half*>(sB + blk * 16384 + v * 128 + (((cc / 8) ^ (v & 7)) * 16) + (cc % 8) * 2) = val;
      }
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tptr)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;

  if (tid == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (c.mn_major) idesc |= (1u << 15) | (1u << 16);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    long long t0 = clock64();
    for (int r = 0; r < c.reps; ++r) {
      for (int j = 0; j < 4; ++j) {
        uint64_t da, db;
        if (!c.mn_major) {
          if (c.a_layout == 0) da = make_desc(a0 + c.shift * 16 + (2 * j) * (181 * 16), 181 * 16, c.gw * 16, 0, 0);
          else da = make_desc(a0 + c.shift * 128 + j * 32, 16, c.gw * 128, 2, c.base_mode ? (c.shift & 7) : 0);
          if (c.b_layout == 0) db = make_desc(b0 + (2 * j) * (c.N * 16), c.N * 16, 128, 0, 0);
          else db = make_desc(b0 + j * 32, 16, 1024, 2, 0);
        } else {
          // MN-major: K step j = voxel rows 2j, 2j+1 (16 voxels); lbo = K-direction stride, sbo = MN-direction stride
          if (c.a_layout == 0) da = make_desc(a0 + c.shift * 16 + j * 2 * (c.gw * 16), c.gw * 16, 181 * 16, 0, 0);
          else da = make_desc(a0 + c.shift * 128 + j * 2 * (c.gw * 128), 32768, c.gw * 128, 2, c.base_mode ? (c.shift & 7) : 0);
          if (c.b_layout == 0) db = make_desc(b0 + j * 2 * 128, 128, 129 * 16, 0, 0);
          else db = make_desc(b0 + j * 2 * 1024, 16384, 1024, 2, 0);
        }
        uint32_t acc = (r | j) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0; long long guard = 0;
    while (!ok && guard++ < (1ll << 26)) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    long long t1 = clock64();
    cycles[blockIdx.x] = ok ? (t1 - t0) : -1;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D[m][n] -> out[m*N + n] (block 0 only)
  for (int n0 = 0; n0 < c.N; n0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0)
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * c.N + n0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

static float h_aval(int m, int k) { return (float)(((m * 7 + k * 3) % 5) - 2); }
static float h_bval(int n, int k) { return (float)(((n * 5 + k * 11) % 7) - 3); }

int main() {
  const int smem = A_BYTES + B_BYTES + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 128 * 256 * sizeof(float)));
  CK(cudaMalloc(&cyc, 256 * sizeof(long long)));
  std::vector<Cfg> cfgs;
  const int reps = 64;
  for (int N : {32, 64, 128, 256}) {
    cfgs.push_back({0, 0, N, 0, 8, 0, reps, 0});     // v1 layouts, dense
    cfgs.push_back({0, 0, N, 11, 10, 0, reps, 0});    // v1 layouts, tap-shifted halo view (what conv_tc does)
    cfgs.push_back({1, 1, N, 0, 8, 0, reps, 0});     // canonical SW128 both
    cfgs.push_back({1, 0, N, 0, 8, 0, reps, 0});     // SW128 A, plain B
    cfgs.push_back({0, 1, N, 0, 8, 0, reps, 0});     // plain A, SW128 B
  }
  // the question that matters: SW128 A read through a shifted start with 10-row group pitch
  for (int s : {0, 1, 3, 8, 11, 21}) {
    cfgs.push_back({1, 1, 128, s, 10, 0, reps, 0});
    cfgs.push_back({1, 1, 128, s, 10, 1, reps, 0});
    cfgs.push_back({1, 1, 128, s, 8, 0, reps, 0});
    cfgs.push_back({1, 1, 128, s, 8, 1, reps, 0});
  }
  // MN-major (wgrad view)
  for (int N : {32, 64}) {
    cfgs.push_back({0, 0, N, 0, 8, 0, reps, 1});
    cfgs.push_back({0, 0, N, 11, 10, 0, reps, 1});
    cfgs.push_back({1, 1, N, 0, 8, 0, reps, 1});
    cfgs.push_back({1, 1, N, 11, 10, 0, reps, 1});
    cfgs.push_back({1, 1, N, 11, 10, 1, reps, 1});
    cfgs.push_back({1, 1, N, 3, 10, 0, reps, 1});
    cfgs.push_back({1, 1, N, 3, 10, 1, reps, 1});
  }
  std::vector<float> h(128 * 256);
  for (auto& c : cfgs) {
    for (int grid : {1, 148}) {
      CK(cudaMemset(out, 0, 128 * 256 * sizeof(float)));
      probe_kernel<<<grid, 128, smem>>>(c, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cfg mn=%d a=%d b=%d N=%d s=%d gw=%d bm=%d: CUDA error %s\n", c.mn_major, c.a_layout, c.b_layout, c.N, c.shift, c.gw, c.base_mode, cudaGetErrorString(e)); return 1; }
      long long hc[256];
      CK(cudaMemcpy(hc, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(h.data(), out, 128 * c.N * sizeof(float), cudaMemcpyDeviceToHost));
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < c.N; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)h_aval(m, k) * h_bval(n, k);
          ref *= c.reps;
          double d = fabs(ref - h[m * c.N + n]);
          if (d > maxerr) maxerr = d;
          if (d > 0.5) ++bad;
        }
      long long mx = 0; for (int i = 0; i < grid; ++i) if (hc[i] > mx) mx = hc[i];
      printf("mn=%d A=%s B=%s N=%3d shift=%2d gw=%2d base_off=%d grid=%3d : %7.1f cyc/MMA  result %s (bad %d, maxerr %.1f)\n",
             c.mn_major, c.a_layout ? "SW128" : "none ", c.b_layout ? "SW128" : "none ", c.N, c.shift, c.gw, c.base_mode, grid,
             (double)mx / (c.reps * 4), bad ? "WRONG" : "ok", bad, maxerr);
    }
  }
  return 0;
}
