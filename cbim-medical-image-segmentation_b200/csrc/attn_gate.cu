// attn_gate.cu — the additive attention gate of Attention-UNet,
// model/dim3/attention_unet_utils.py:7-37 (AttentionBlock):
//     t   = relu(IN(W_g g) + IN(W_x x))                    (the two 1x1x1 convs run on the tcgen05 conv kernel,
//                                                           the sum is the resblock_out kernel with ReLU)
//     psi = sigmoid(IN(conv1x1(t; int_ch -> 1)))           <- gate_rowdot_* here: a per-voxel dot product, not a GEMM
//     out = x * psi                                        <- gate_apply_* here
// The reference runs conv(int_ch->1) + InstanceNorm3d(1) + Sigmoid + a broadcast multiply as four library kernels with
// three [B,1,D,H,W] intermediates; here the forward is two passes (dot + its two sums; gate + the IN sums the next conv's
// loader needs) and the backward two (d/dx, d/dpsi and the norm's two sums; d/dt and d/dw).  HBM-bound: 16-byte loads of
// 8 channels per thread, fp64 atomics for the per-(b,c) sums.  Layout NDHWC with (ld, coff) slices as everywhere else.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf(float h) { return 1.f / (1.f + __expf(-h)); }

// mean / rstd of the single psi channel of batch b from its {sum, sumsq}
__device__ __forceinline__ void psi_norm(const double* pstats, int b, int64_t V, float eps, float& mean, float& rstd) {
  stats_to_mean_rstd(pstats + (int64_t)b * 2, (double)V, eps, mean, rstd);
}

// ---- p[b][v] = round_T( sum_c w[c] * t[b][v][c] ), pstats[b] += {sum p, sum p^2}
// 8 lanes per voxel (each lane walks 16-byte chunks j = lane8, lane8+8, ...), 32 voxels per block pass.
template <typename T>
__global__ void __launch_bounds__(kThreads)
gate_rowdot_fwd_kernel(const T* __restrict__ t, int t_ld, const float* __restrict__ w, float* __restrict__ p, double* __restrict__ pstats,
                       int64_t V, int C, int64_t vpb) {
  const int b = blockIdx.y, l8 = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int64_t v0 = (int64_t)blockIdx.x * vpb, v1 = (v0 + vpb < V) ? v0 + vpb : V;
  const T* tb = t + (int64_t)b * V * t_ld;
  float s1 = 0.f, s2 = 0.f;
  // the trip count is uniform over the block (the four voxels of a warp may straddle v1: predicate, do not diverge —
  // the shuffles below need every lane)
  for (int64_t vb = v0; vb < v1; vb += kThreads / 8) {
    const int64_t v = vb + slot;
    const bool live = v < v1;
    float acc = 0.f;
    for (int j = l8; live && j < C / 8; j += 8) {
      float a[8];
      ld8<T>(tb + v * t_ld + j * 8, a);
      const float4 w0 = *reinterpret_cast<const float4*>(w + j * 8), w1 = *reinterpret_cast<const float4*>(w + j * 8 + 4);
      acc += a[0] * w0.x + a[1] * w0.y + a[2] * w0.z + a[3] * w0.w + a[4] * w1.x + a[5] * w1.y + a[6] * w1.z + a[7] * w1.w;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (live && l8 == 0) {
      const float r = Elem<T>::round(acc);          // the conv output is a storage-dtype tensor in the reference
      p[(int64_t)b * V + v] = r;
      s1 += r; s2 += r * r;
    }
  }
  __shared__ double sh[2][kThreads / 32];
  double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = d1; sh[1][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) s += sh[threadIdx.x][i];
    atomicAdd(&pstats[(int64_t)b * 2 + threadIdx.x], s);
  }
}

// Thread layout of the channel-stationary kernels: thread owns channel chunk (tid % cpv) of voxel slot (tid / cpv).
struct ChanIter { int cpv, vpp, cchunk, vloc; bool active; };
__device__ __forceinline__ ChanIter chan_iter(int C) {
  ChanIter it;
  it.cpv = C / 8; it.vpp = kThreads / it.cpv; it.cchunk = threadIdx.x % it.cpv; it.vloc = threadIdx.x / it.cpv;
  it.active = it.vloc < it.vpp;
  return it;
}
// sum NV*8 per-thread partials over the threads sharing a channel chunk; dst[c*NV + k] += (double / float atomics)
template <int NV, typename D>
__device__ __forceinline__ void chan_reduce(const float* acc, const ChanIter& it, float* smem, D* dst, int C) {
  constexpr int W = NV * 8;
#pragma unroll
  for (int i = 0; i < W; ++i) smem[threadIdx.x * W + i] = it.active ? acc[i] : 0.f;
  __syncthreads();
  for (int o = threadIdx.x; o < C * NV; o += kThreads) {
    const int c = o / NV, k = o % NV, chunk = c / 8, e = c % 8;
    double s = 0.0;
    for (int vl = 0; vl < it.vpp; ++vl) s += (double)smem[(vl * it.cpv + chunk) * W + k * 8 + e];
    atomicAdd(&dst[c * NV + k], (D)s);
  }
}

// ---- out[b][v][c] = round_T(x * sigmoid(IN(p))), ostats[b][c] += {sum, sumsq} of out
template <typename T>
__global__ void __launch_bounds__(kThreads)
gate_apply_fwd_kernel(const T* __restrict__ x, int x_ld, int x_coff, const float* __restrict__ p, const double* __restrict__ pstats,
                      float eps, T* __restrict__ out, int o_ld, int o_coff, double* __restrict__ ostats, int64_t V, int C, int64_t vpb) {
  extern __shared__ float smem[];
  const ChanIter it = chan_iter(C);
  const int b = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * vpb, v1 = (v0 + vpb < V) ? v0 + vpb : V;
  float mean, rstd;
  psi_norm(pstats, b, V, eps, mean, rstd);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (it.active) {
    const T* xb = x + (int64_t)b * V * x_ld + x_coff + it.cchunk * 8;
    T* ob = out + (int64_t)b * V * o_ld + o_coff + it.cchunk * 8;
    for (int64_t v = v0 + it.vloc; v < v1; v += it.vpp) {
      const float s = sigmoidf((p[(int64_t)b * V + v] - mean) * rstd);
      float a[8];
      ld8<T>(xb + v * x_ld, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = Elem<T>::round(a[i] * s); acc[i] += a[i]; acc[8 + i] += a[i] * a[i]; }
      st8<T>(ob + v * o_ld, a);
    }
  }
  if (ostats) chan_reduce<2, double>(acc, it, smem, ostats + (int64_t)b * C * 2, C);
}

// ---- dx[b][v][c] = dout * s;  dz[b][v] = (sum_c dout * x) * s (1 - s);  bsums[b] += {sum dz, sum dz * phat}
template <typename T>
__global__ void __launch_bounds__(kThreads)
gate_apply_bwd_kernel(const T* __restrict__ dout, int d_ld, int d_coff, const T* __restrict__ x, int x_ld, int x_coff,
                      const float* __restrict__ p, const double* __restrict__ pstats, float eps, T* __restrict__ dx, float* __restrict__ dz,
                      double* __restrict__ bsums, int64_t V, int C, int64_t vpb) {
  const int b = blockIdx.y, l8 = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int64_t v0 = (int64_t)blockIdx.x * vpb, v1 = (v0 + vpb < V) ? v0 + vpb : V;
  float mean, rstd;
  psi_norm(pstats, b, V, eps, mean, rstd);
  const T* db = dout + (int64_t)b * V * d_ld + d_coff;
  const T* xb = x + (int64_t)b * V * x_ld + x_coff;
  T* dxb = dx + (int64_t)b * V * C;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t vb = v0; vb < v1; vb += kThreads / 8) {          // block-uniform trip count, see gate_rowdot_fwd_kernel
    const int64_t v = vb + slot;
    const bool live = v < v1;
    const float ph = live ? (p[(int64_t)b * V + v] - mean) * rstd : 0.f;
    const float s = sigmoidf(ph);
    float dot = 0.f;
    for (int j = l8; live && j < C / 8; j += 8) {
      float g[8], a[8];
      ld8<T>(db + v * d_ld + j * 8, g);
      ld8<T>(xb + v * x_ld + j * 8, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) { dot += g[i] * a[i]; g[i] *= s; }
      st8<T>(dxb + v * C + j * 8, g);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    if (live && l8 == 0) {
      const float z = dot * s * (1.f - s);
      dz[(int64_t)b * V + v] = z;
      s1 += z; s2 += z * ph;
    }
  }
  __shared__ double sh[2][kThreads / 32];
  double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = d1; sh[1][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) s += sh[threadIdx.x][i];
    atomicAdd(&bsums[(int64_t)b * 2 + threadIdx.x], s);
  }
}

// ---- dp = rstd (dz - S1/V - phat S2/V)  (InstanceNorm backward of the single psi channel);
//      dt[b][v][c] = dp * w[c];  dw[c] += sum_{b,v} dp * t[b][v][c]
template <typename T>
__global__ void __launch_bounds__(kThreads)
gate_rowdot_bwd_kernel(const float* __restrict__ dz, const double* __restrict__ bsums, const float* __restrict__ p,
                       const double* __restrict__ pstats, float eps, const T* __restrict__ t, int t_ld, const float* __restrict__ w,
                       T* __restrict__ dt, float* __restrict__ dw, int64_t V, int C, int64_t vpb) {
  extern __shared__ float smem[];
  const ChanIter it = chan_iter(C);
  const int b = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * vpb, v1 = (v0 + vpb < V) ? v0 + vpb : V;
  float mean, rstd;
  psi_norm(pstats, b, V, eps, mean, rstd);
  const float m1 = (float)(bsums[(int64_t)b * 2] / (double)V), m2 = (float)(bsums[(int64_t)b * 2 + 1] / (double)V);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (it.active) {
    float wc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wc[i] = w[it.cchunk * 8 + i];
    const T* tb = t + (int64_t)b * V * t_ld + it.cchunk * 8;
    T* dtb = dt + (int64_t)b * V * C + it.cchunk * 8;
    for (int64_t v = v0 + it.vloc; v < v1; v += it.vpp) {
      const float ph = (p[(int64_t)b * V + v] - mean) * rstd;
      const float dp = Elem<T>::round(rstd * (dz[(int64_t)b * V + v] - m1 - ph * m2));
      float a[8], o[8];
      ld8<T>(tb + v * t_ld, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[i] += dp * a[i]; o[i] = dp * wc[i]; }
      st8<T>(dtb + v * C, o);
    }
  }
  chan_reduce<1, float>(acc, it, smem, dw, C);
}

inline int64_t pick_vpb(int64_t V, int B) {
  // ~4 blocks per SM in total, at least one pass of 32 voxels per block
  int64_t blocks = (int64_t)B200SEG_NUM_SMS * 4 / (B > 0 ? B : 1);
  if (blocks < 1) blocks = 1;
  int64_t vpb = (V + blocks - 1) / blocks;
  if (vpb < 256) vpb = 256;
  return vpb;
}
inline bool al16(const void* ptr, int ld, int coff, int esz) {
  return ((reinterpret_cast<uintptr_t>(ptr) + (size_t)coff * esz) % 16 == 0) && ((size_t)ld * esz % 16 == 0);
}

}  // namespace

#define GATE_DISPATCH(DT, ...)                                        \
  if ((DT) == B200SEG_F16) { using T = __half; __VA_ARGS__ }           \
  else if ((DT) == B200SEG_F32) { using T = float; __VA_ARGS__ }       \
  else return B200SEG_EINVAL;

static int gate_shape_ok(int B, int64_t V, int C) {
  if (B <= 0 || V <= 0 || C <= 0) return B200SEG_EINVAL;
  if (C % 8 != 0 || C > 8 * kThreads) return B200SEG_EUNSUPPORTED;     // int_ch = out_ch / 2 of attention_up_block: 16 .. 128
  return B200SEG_OK;
}

extern "C" int b200seg_attn_gate_fwd(const void* t, int t_ld, const float* w, const void* x, int x_ld, int x_coff, float eps,
                                     float* p, double* pstats, void* out, int o_ld, int o_coff, double* ostats, int B, int64_t V,
                                     int Ct, int Cx, int dtype, void* stream) {
  if (!t || !w || !x || !p || !pstats || !out) return B200SEG_EINVAL;
  int rc = gate_shape_ok(B, V, Ct);
  if (rc) return rc;
  rc = gate_shape_ok(B, V, Cx);
  if (rc) return rc;
  const int esz = dtype == B200SEG_F16 ? 2 : 4;
  if (!al16(t, t_ld, 0, esz) || !al16(x, x_ld, x_coff, esz) || !al16(out, o_ld, o_coff, esz) || !al16(w, 4, 0, 4)) return B200SEG_EUNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int64_t vpb = pick_vpb(V, B);
  dim3 grid(ceil_div(V, vpb), B);
  GATE_DISPATCH(dtype, {
    gate_rowdot_fwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)t, t_ld, w, p, pstats, V, Ct, vpb);
    B200_CHECK_LAUNCH("gate_rowdot_fwd_kernel");
    gate_apply_fwd_kernel<T><<<grid, kThreads, sizeof(float) * kThreads * 16, st>>>((const T*)x, x_ld, x_coff, p, pstats, eps, (T*)out, o_ld,
                                                                                  o_coff, ostats, V, Cx, vpb);
  })
  B200_CHECK_LAUNCH("gate_apply_fwd_kernel");
  return B200SEG_OK;
}

extern "C" int b200seg_attn_gate_bwd(const void* dout, int d_ld, int d_coff, const void* x, int x_ld, int x_coff, const void* t, int t_ld,
                                     const float* w, const float* p, const double* pstats, float eps, void* dx, void* dt, float* dw,
                                     float* dz, double* bsums, int B, int64_t V, int Ct, int Cx, int dtype, void* stream) {
  if (!dout || !x || !t || !w || !p || !pstats || !dx || !dt || !dw || !dz || !bsums) return B200SEG_EINVAL;
  int rc = gate_shape_ok(B, V, Ct);
  if (rc) return rc;
  rc = gate_shape_ok(B, V, Cx);
  if (rc) return rc;
  const int esz = dtype == B200SEG_F16 ? 2 : 4;
  if (!al16(dout, d_ld, d_coff, esz) || !al16(x, x_ld, x_coff, esz) || !al16(t, t_ld, 0, esz) || !al16(dx, Cx, 0, esz) ||
      !al16(dt, Ct, 0, esz) || !al16(w, 4, 0, 4))
    return B200SEG_EUNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int64_t vpb = pick_vpb(V, B);
  dim3 grid(ceil_div(V, vpb), B);
  GATE_DISPATCH(dtype, {
    gate_apply_bwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)dout, d_ld, d_coff, (const T*)x, x_ld, x_coff, p, pstats, eps, (T*)dx, dz,
                                                       bsums, V, Cx, vpb);
    B200_CHECK_LAUNCH("gate_apply_bwd_kernel");
    gate_rowdot_bwd_kernel<T><<<grid, kThreads, sizeof(float) * kThreads * 8, st>>>(dz, bsums, p, pstats, eps, (const T*)t, t_ld, w, (T*)dt, dw,
                                                                                   V, Ct, vpb);
  })
  B200_CHECK_LAUNCH("gate_rowdot_bwd_kernel");
  return B200SEG_OK;
}
