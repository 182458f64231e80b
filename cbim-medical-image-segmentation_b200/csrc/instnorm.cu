// instnorm.cu — InstanceNorm3d (affine=False, no running stats) statistics, apply and the
// two-stage backward, on NDHWC tensors with (ld, coff) channel slicing.  All HBM-bound:
// 128-bit loads (8 fp16 channels per thread), warp/block reductions, fp64 atomics for the
// per-(b,c) sums.  Reference: nn.InstanceNorm3d(eps=1e-4) at conv_layers.py:40,42 + nn.ReLU :43.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// Work decomposition shared by every kernel here: grid = (voxel chunks, B).  A block walks
// its voxel range in passes of `vpp` voxels; thread t owns channel-chunk (t % cpv) of voxel
// (t / cpv) in each pass, so a thread's channels are fixed for the whole kernel.
struct Iter {
  int cpv;      // VEC-wide chunks per voxel
  int vpp;      // voxels per pass
  int64_t v0, v1;
  int cchunk;   // this thread's chunk (channel = cchunk*VEC)
  int vloc;     // this thread's voxel slot in a pass
  bool active;
};
template <int VEC>
__device__ __forceinline__ Iter make_iter(int C, int64_t V, int64_t vox_per_block) {
  Iter it;
  it.cpv = C / VEC;
  it.vpp = kThreads / it.cpv;
  if (it.vpp < 1) it.vpp = 1;
  it.v0 = (int64_t)blockIdx.x * vox_per_block;
  it.v1 = it.v0 + vox_per_block;
  if (it.v1 > V) it.v1 = V;
  it.cchunk = threadIdx.x % it.cpv;
  it.vloc = threadIdx.x / it.cpv;
  it.active = it.vloc < it.vpp;
  return it;
}

template <int VEC, typename T> struct Vec;
template <typename T> struct Vec<8, T> {
  static __device__ __forceinline__ void ld(const T* p, float (&v)[8]) { ld8<T>(p, v); }
  static __device__ __forceinline__ void st(T* p, const float (&v)[8]) { st8<T>(p, v); }
};
template <typename T> struct Vec<1, T> {
  static __device__ __forceinline__ void ld(const T* p, float (&v)[1]) { v[0] = Elem<T>::ld(p); }
  static __device__ __forceinline__ void st(T* p, const float (&v)[1]) { Elem<T>::st(p, v[0]); }
};

// Reduce per-thread partials over all threads that share a channel chunk, then fp64 atomics.
// acc: NV*VEC values per thread.  smem: float[kThreads][NV*VEC].
template <int VEC, int NV>
__device__ __forceinline__ void block_reduce_to_global(const float* acc, const Iter& it, float* smem,
                                                       double* gdst /* [C][NV] for this b */, int C) {
  constexpr int W = NV * VEC;
#pragma unroll
  for (int i = 0; i < W; ++i) smem[threadIdx.x * W + i] = it.active ? acc[i] : 0.f;
  __syncthreads();
  // one thread per (channel, value)
  for (int o = threadIdx.x; o < C * NV; o += kThreads) {
    int c = o / NV, k = o % NV;
    int chunk = c / VEC, e = c % VEC;
    double s = 0.0;
    for (int vl = 0; vl < it.vpp; ++vl) s += (double)smem[(vl * it.cpv + chunk) * W + k * VEC + e];
    atomicAdd(&gdst[c * NV + k], s);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
stats_kernel(const T* __restrict__ x, int ld, int coff, int64_t V, int C, int64_t vpb, double* __restrict__ stats) {
  extern __shared__ float smem[];
  Iter it = make_iter<VEC>(C, V, vpb);
  const int b = blockIdx.y;
  float acc[2 * VEC];
#pragma unroll
  for (int i = 0; i < 2 * VEC; ++i) acc[i] = 0.f;
  if (it.active) {
    const T* base = x + (int64_t)b * V * ld + coff + it.cchunk * VEC;
    for (int64_t v = it.v0 + it.vloc; v < it.v1; v += it.vpp) {
      float a[VEC];
      Vec<VEC, T>::ld(base + v * ld, a);
#pragma unroll
      for (int i = 0; i < VEC; ++i) { acc[i] += a[i]; acc[VEC + i] += a[i] * a[i]; }
    }
  }
  block_reduce_to_global<VEC, 2>(acc, it, smem, stats + (int64_t)b * C * 2, C);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
apply_kernel(const T* __restrict__ x, int x_ld, int x_coff, const double* __restrict__ stats, float eps, int act,
             T* __restrict__ y, int y_ld, int y_coff, int64_t V, int C, int64_t vpb) {
  Iter it = make_iter<VEC>(C, V, vpb);
  const int b = blockIdx.y;
  if (!it.active) return;
  float mean[VEC], rstd[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i)
    stats_to_mean_rstd(stats + ((int64_t)b * C + it.cchunk * VEC + i) * 2, (double)V, eps, mean[i], rstd[i]);
  const T* xb = x + (int64_t)b * V * x_ld + x_coff + it.cchunk * VEC;
  T* yb = y + (int64_t)b * V * y_ld + y_coff + it.cchunk * VEC;
  for (int64_t v = it.v0 + it.vloc; v < it.v1; v += it.vpp) {
    float a[VEC];
    Vec<VEC, T>::ld(xb + v * x_ld, a);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float h = (a[i] - mean[i]) * rstd[i];
      a[i] = act_apply(h, act);
    }
    Vec<VEC, T>::st(yb + v * y_ld, a);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
bwd_reduce_kernel(const T* __restrict__ dy, int dy_ld, int dy_coff, const T* __restrict__ x, int x_ld, int x_coff,
                  const double* __restrict__ stats, float eps, int act, T* __restrict__ g, int g_ld, int g_coff,
                  double* __restrict__ bstats, int64_t V, int C, int64_t vpb) {
  extern __shared__ float smem[];
  Iter it = make_iter<VEC>(C, V, vpb);
  const int b = blockIdx.y;
  float acc[2 * VEC];
#pragma unroll
  for (int i = 0; i < 2 * VEC; ++i) acc[i] = 0.f;
  if (it.active) {
    float mean[VEC], rstd[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      stats_to_mean_rstd(stats + ((int64_t)b * C + it.cchunk * VEC + i) * 2, (double)V, eps, mean[i], rstd[i]);
    const T* dyb = dy + (int64_t)b * V * dy_ld + dy_coff + it.cchunk * VEC;
    const T* xb = x + (int64_t)b * V * x_ld + x_coff + it.cchunk * VEC;
    T* gb = g + (int64_t)b * V * g_ld + g_coff + it.cchunk * VEC;
    for (int64_t v = it.v0 + it.vloc; v < it.v1; v += it.vpp) {
      float a[VEC], d[VEC];
      Vec<VEC, T>::ld(xb + v * x_ld, a);
      Vec<VEC, T>::ld(dyb + v * dy_ld, d);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float h = (a[i] - mean[i]) * rstd[i];
        float gg = d[i] * act_grad(h, act);
        gg = Elem<T>::round(gg);
        d[i] = gg;
        acc[i] += gg;
        acc[VEC + i] += gg * h;
      }
      Vec<VEC, T>::st(gb + v * g_ld, d);
    }
  }
  block_reduce_to_global<VEC, 2>(acc, it, smem, bstats + (int64_t)b * C * 2, C);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
bwd_apply_kernel(const T* __restrict__ g, int g_ld, int g_coff, const T* __restrict__ x, int x_ld, int x_coff,
                 const double* __restrict__ stats, const double* __restrict__ bstats, float eps,
                 const T* add, int add_ld, int add_coff,
                 T* dx, int dx_ld, int dx_coff, int64_t V, int C, int64_t vpb) {
  Iter it = make_iter<VEC>(C, V, vpb);
  const int b = blockIdx.y;
  if (!it.active) return;
  float mean[VEC], rstd[VEC], m1[VEC], m2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    int64_t o = ((int64_t)b * C + it.cchunk * VEC + i) * 2;
    stats_to_mean_rstd(stats + o, (double)V, eps, mean[i], rstd[i]);
    m1[i] = (float)(bstats[o] / (double)V);
    m2[i] = (float)(bstats[o + 1] / (double)V);
  }
  const T* gb = g + (int64_t)b * V * g_ld + g_coff + it.cchunk * VEC;
  const T* xb = x + (int64_t)b * V * x_ld + x_coff + it.cchunk * VEC;
  T* dxb = dx + (int64_t)b * V * dx_ld + dx_coff + it.cchunk * VEC;
  const T* addb = add ? add + (int64_t)b * V * add_ld + add_coff + it.cchunk * VEC : nullptr;
  const bool accumulate = add != nullptr;
  for (int64_t v = it.v0 + it.vloc; v < it.v1; v += it.vpp) {
    float a[VEC], d[VEC], o[VEC];
    Vec<VEC, T>::ld(xb + v * x_ld, a);
    Vec<VEC, T>::ld(gb + v * g_ld, d);
    if (accumulate) Vec<VEC, T>::ld(addb + v * add_ld, o);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float h = (a[i] - mean[i]) * rstd[i];
      float r = rstd[i] * (d[i] - m1[i] - h * m2[i]);
      o[i] = accumulate ? o[i] + r : r;
    }
    Vec<VEC, T>::st(dxb + v * dx_ld, o);
  }
}

template <typename TX, typename TY, int VEC>
__global__ void __launch_bounds__(kThreads)
copy_kernel(const TX* __restrict__ x, int x_ld, int x_coff, TY* __restrict__ y, int y_ld, int y_coff,
            int accumulate, int64_t V, int C, int64_t vpb) {
  Iter it = make_iter<VEC>(C, V, vpb);
  if (!it.active) return;
  const TX* xb = x + x_coff + it.cchunk * VEC;
  TY* yb = y + y_coff + it.cchunk * VEC;
  for (int64_t v = it.v0 + it.vloc; v < it.v1; v += it.vpp) {
    float a[VEC];
    Vec<VEC, TX>::ld(xb + v * x_ld, a);
    if (accumulate) {
      float o[VEC];
      Vec<VEC, TY>::ld(yb + v * y_ld, o);
#pragma unroll
      for (int i = 0; i < VEC; ++i) a[i] += o[i];
    }
    Vec<VEC, TY>::st(yb + v * y_ld, a);
  }
}

inline bool vec_ok(const void* p, int ld, int coff, int C, int esz) {
  return (C % 8 == 0) && (ld % 8 == 0) && (coff % 8 == 0) && (C / 8 <= kThreads) &&
         ((reinterpret_cast<uintptr_t>(p) % 16) == 0) && esz > 0;
}
inline int64_t pick_vpb(int64_t V, int B, int C) {
  // aim for >= ~6 blocks per SM in total, but keep >= 8 passes per block: a pass covers kThreads/(C/8) voxels,
  // so wide-channel / few-voxel tensors (MedFormer's 1280 x 864) still spread over the whole GPU
  int64_t want = (int64_t)B200SEG_NUM_SMS * 6 / (B > 0 ? B : 1);
  if (want < 1) want = 1;
  int64_t vpb = (V + want - 1) / want;
  int cpv = C % 8 ? C : C / 8;
  int64_t vpp = kThreads / (cpv > 0 ? cpv : 1);
  if (vpp < 1) vpp = 1;
  if (vpb < 8 * vpp) vpb = 8 * vpp;
  return vpb;
}


// ---- output stage of monai's UnetResBlock (call sites swin_unetr.py:129-226):
//   y = lrelu( IN(r2) + res ),  res = IN(r3) when the block has its 1x1 projection (stats3 != NULL), else r3 as is.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
resout_fwd_kernel(const T* __restrict__ r2, int r2_ld, const double* __restrict__ stats2, const T* __restrict__ r3, int r3_ld,
                  int r3_coff, const double* __restrict__ stats3, float eps, int act, T* __restrict__ y, int y_ld,
                  int64_t V, int C, int64_t vpb) {
  Iter it = make_iter<VEC>(C, V, vpb);
  const int b = blockIdx.y;
  if (!it.active) return;
  float m2[VEC], s2[VEC], m3[VEC], s3[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int64_t o = ((int64_t)b * C + it.cchunk * VEC + i) * 2;
    stats_to_mean_rstd(stats2 + o, (double)V, eps, m2[i], s2[i]);
    m3[i] = 0.f; s3[i] = 1.f;
    if (stats3) stats_to_mean_rstd(stats3 + o, (double)V, eps, m3[i], s3[i]);
  }
  const T* a2 = r2 + (int64_t)b * V * r2_ld + it.cchunk * VEC;
  const T* a3 = r3 + (int64_t)b * V * r3_ld + r3_coff + it.cchunk * VEC;
  T* yb = y + (int64_t)b * V * y_ld + it.cchunk * VEC;
  for (int64_t v = it.v0 + it.vloc; v < it.v1; v += it.vpp) {
    float a[VEC], c[VEC];
    Vec<VEC, T>::ld(a2 + v * r2_ld, a);
    Vec<VEC, T>::ld(a3 + v * r3_ld, c);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      // both normalised branches are rounded to the storage dtype before the add, like the reference's tensors
      const float h2 = Elem<T>::round((a[i] - m2[i]) * s2[i]);
      const float h3 = stats3 ? Elem<T>::round((c[i] - m3[i]) * s3[i]) : c[i];
      a[i] = act_apply(Elem<T>::round(h2 + h3), act);
    }
    Vec<VEC, T>::st(yb + v * y_ld, a);
  }
}

// g = dy * act'(y) (the sign of y is the sign of the pre-activation) and the three sums {g, g*xhat2, g*xhat3} per (b,c)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
resout_bwd_reduce_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld, const T* __restrict__ r2, int r2_ld,
                         const double* __restrict__ stats2, const T* __restrict__ r3, int r3_ld, int r3_coff,
                         const double* __restrict__ stats3, float eps, int act, T* __restrict__ g, double* __restrict__ sums,
                         int64_t V, int C, int64_t vpb) {
  extern __shared__ float smem[];
  Iter it = make_iter<VEC>(C, V, vpb);
  const int b = blockIdx.y;
  float acc[3 * VEC];
#pragma unroll
  for (int i = 0; i < 3 * VEC; ++i) acc[i] = 0.f;
  if (it.active) {
    float m2[VEC], s2[VEC], m3[VEC], s3[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int64_t o = ((int64_t)b * C + it.cchunk * VEC + i) * 2;
      stats_to_mean_rstd(stats2 + o, (double)V, eps, m2[i], s2[i]);
      m3[i] = 0.f; s3[i] = 1.f;
      if (stats3) stats_to_mean_rstd(stats3 + o, (double)V, eps, m3[i], s3[i]);
    }
    const T* dyb = dy + (int64_t)b * V * dy_ld + it.cchunk * VEC;
    const T* yb = y + (int64_t)b * V * y_ld + it.cchunk * VEC;
    const T* a2 = r2 + (int64_t)b * V * r2_ld + it.cchunk * VEC;
    const T* a3 = r3 + (int64_t)b * V * r3_ld + r3_coff + it.cchunk * VEC;
    T* gb = g + (int64_t)b * V * C + it.cchunk * VEC;
    for (int64_t v = it.v0 + it.vloc; v < it.v1; v += it.vpp) {
      float d[VEC], o[VEC], a[VEC], c[VEC];
      Vec<VEC, T>::ld(dyb + v * dy_ld, d);
      Vec<VEC, T>::ld(yb + v * y_ld, o);
      Vec<VEC, T>::ld(a2 + v * r2_ld, a);
      if (stats3) Vec<VEC, T>::ld(a3 + v * r3_ld, c);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float gg = Elem<T>::round(d[i] * act_grad(o[i], act));
        d[i] = gg;
        acc[i] += gg;
        acc[VEC + i] += gg * ((a[i] - m2[i]) * s2[i]);
        if (stats3) acc[2 * VEC + i] += gg * ((c[i] - m3[i]) * s3[i]);
      }
      Vec<VEC, T>::st(gb + v * C, d);
    }
  }
  block_reduce_to_global<VEC, 3>(acc, it, smem, sums + (int64_t)b * C * 3, C);
}

}  // namespace

#define DISPATCH_TV(DT, VECOK, ...)                                         \
  if ((DT) == B200SEG_F16) {                                                 \
    using T = __half;                                                        \
    if (VECOK) { constexpr int VEC = 8; __VA_ARGS__ } else { constexpr int VEC = 1; __VA_ARGS__ } \
  } else if ((DT) == B200SEG_F32) {                                          \
    using T = float;                                                         \
    if (VECOK) { constexpr int VEC = 8; __VA_ARGS__ } else { constexpr int VEC = 1; __VA_ARGS__ } \
  } else return B200SEG_EINVAL;

extern "C" int b200seg_instnorm_stats(const void* x, int dtype, int ld, int coff, int B, int64_t V, int C,
                                      double* stats, void* stream) {
  if (!x || !stats || B <= 0 || V <= 0 || C <= 0 || C > 4096) return B200SEG_EINVAL;
  cudaStream_t st = as_stream(stream);
  int64_t vpb = pick_vpb(V, B, C);
  dim3 grid(ceil_div(V, vpb), B);
  bool vok = vec_ok(x, ld, coff, C, 1);
  if (!vok && C > kThreads) return B200SEG_EUNSUPPORTED;
  DISPATCH_TV(dtype, vok, {
    size_t sm = sizeof(float) * kThreads * 2 * VEC;
    stats_kernel<T, VEC><<<grid, kThreads, sm, st>>>((const T*)x, ld, coff, V, C, vpb, stats);
  })
  B200_CHECK_LAUNCH("stats_kernel");
  return B200SEG_OK;
}

extern "C" int b200seg_instnorm_apply(const void* x, int dtype, int x_ld, int x_coff, const double* stats,
                                      float eps, int act, void* y, int y_ld, int y_coff, int B, int64_t V,
                                      int C, void* stream) {
  if (!x || !y || !stats || B <= 0 || V <= 0 || C <= 0) return B200SEG_EINVAL;
  cudaStream_t st = as_stream(stream);
  int64_t vpb = pick_vpb(V, B, C);
  dim3 grid(ceil_div(V, vpb), B);
  bool vok = vec_ok(x, x_ld, x_coff, C, 1) && vec_ok(y, y_ld, y_coff, C, 1);
  if (!vok && C > kThreads) return B200SEG_EUNSUPPORTED;
  DISPATCH_TV(dtype, vok, {
    apply_kernel<T, VEC><<<grid, kThreads, 0, st>>>((const T*)x, x_ld, x_coff, stats, eps, act, (T*)y, y_ld, y_coff, V, C, vpb);
  })
  B200_CHECK_LAUNCH("apply_kernel");
  return B200SEG_OK;
}

extern "C" int b200seg_instnorm_bwd_reduce(const void* dy, int dy_ld, int dy_coff, const void* x, int x_ld,
                                           int x_coff, int dtype, const double* stats, float eps, int act,
                                           void* g, int g_ld, int g_coff, double* bstats, int B, int64_t V,
                                           int C, void* stream) {
  if (!dy || !x || !stats || !g || !bstats || B <= 0 || V <= 0 || C <= 0) return B200SEG_EINVAL;
  cudaStream_t st = as_stream(stream);
  int64_t vpb = pick_vpb(V, B, C);
  dim3 grid(ceil_div(V, vpb), B);
  bool vok = vec_ok(dy, dy_ld, dy_coff, C, 1) && vec_ok(x, x_ld, x_coff, C, 1) && vec_ok(g, g_ld, g_coff, C, 1);
  if (!vok && C > kThreads) return B200SEG_EUNSUPPORTED;
  DISPATCH_TV(dtype, vok, {
    size_t sm = sizeof(float) * kThreads * 2 * VEC;
    bwd_reduce_kernel<T, VEC><<<grid, kThreads, sm, st>>>((const T*)dy, dy_ld, dy_coff, (const T*)x, x_ld, x_coff, stats,
                                                         eps, act, (T*)g, g_ld, g_coff, bstats, V, C, vpb);
  })
  B200_CHECK_LAUNCH("bwd_reduce_kernel");
  return B200SEG_OK;
}

extern "C" int b200seg_instnorm_bwd_apply(const void* g, int g_ld, int g_coff, const void* x, int x_ld, int x_coff,
                                          int dtype, const double* stats, const double* bstats, float eps,
                                          const void* add, int add_ld, int add_coff,
                                          void* dx, int dx_ld, int dx_coff, int B, int64_t V,
                                          int C, void* stream) {
  if (!g || !x || !stats || !bstats || !dx || B <= 0 || V <= 0 || C <= 0) return B200SEG_EINVAL;
  cudaStream_t st = as_stream(stream);
  int64_t vpb = pick_vpb(V, B, C);
  dim3 grid(ceil_div(V, vpb), B);
  bool vok = vec_ok(g, g_ld, g_coff, C, 1) && vec_ok(x, x_ld, x_coff, C, 1) && vec_ok(dx, dx_ld, dx_coff, C, 1) &&
             (!add || vec_ok(add, add_ld, add_coff, C, 1));
  if (!vok && C > kThreads) return B200SEG_EUNSUPPORTED;
  DISPATCH_TV(dtype, vok, {
    bwd_apply_kernel<T, VEC><<<grid, kThreads, 0, st>>>((const T*)g, g_ld, g_coff, (const T*)x, x_ld, x_coff, stats, bstats,
                                                        eps, (const T*)add, add_ld, add_coff, (T*)dx, dx_ld, dx_coff, V, C, vpb);
  })
  B200_CHECK_LAUNCH("bwd_apply_kernel");
  return B200SEG_OK;
}

template <typename TX, typename TY>
static int copy_t(const void* x, int x_ld, int x_coff, void* y, int y_ld, int y_coff, int accumulate,
                  int64_t nvox, int C, cudaStream_t st) {
  int64_t vpb = pick_vpb(nvox, 1, C);
  dim3 grid(ceil_div(nvox, vpb), 1);
  bool vok = vec_ok(x, x_ld, x_coff, C, 1) && vec_ok(y, y_ld, y_coff, C, 1);
  if (!vok && C > kThreads) return B200SEG_EUNSUPPORTED;
  if (vok) copy_kernel<TX, TY, 8><<<grid, kThreads, 0, st>>>((const TX*)x, x_ld, x_coff, (TY*)y, y_ld, y_coff, accumulate, nvox, C, vpb);
  else copy_kernel<TX, TY, 1><<<grid, kThreads, 0, st>>>((const TX*)x, x_ld, x_coff, (TY*)y, y_ld, y_coff, accumulate, nvox, C, vpb);
  return 0;
}

extern "C" int b200seg_copy_channels(const void* x, int x_dtype, int x_ld, int x_coff, void* y, int y_dtype,
                                     int y_ld, int y_coff, int accumulate, int64_t nvox, int C, void* stream) {
  if (!x || !y || nvox <= 0 || C <= 0) return B200SEG_EINVAL;
  cudaStream_t st = as_stream(stream);
  int rc;
  if (x_dtype == B200SEG_F16 && y_dtype == B200SEG_F16) rc = copy_t<__half, __half>(x, x_ld, x_coff, y, y_ld, y_coff, accumulate, nvox, C, st);
  else if (x_dtype == B200SEG_F32 && y_dtype == B200SEG_F16) rc = copy_t<float, __half>(x, x_ld, x_coff, y, y_ld, y_coff, accumulate, nvox, C, st);
  else if (x_dtype == B200SEG_F16 && y_dtype == B200SEG_F32) rc = copy_t<__half, float>(x, x_ld, x_coff, y, y_ld, y_coff, accumulate, nvox, C, st);
  else if (x_dtype == B200SEG_F32 && y_dtype == B200SEG_F32) rc = copy_t<float, float>(x, x_ld, x_coff, y, y_ld, y_coff, accumulate, nvox, C, st);
  else return B200SEG_EINVAL;
  if (rc) return rc;
  B200_CHECK_LAUNCH("copy_kernel");
  return B200SEG_OK;
}

extern "C" int b200seg_resblock_out_fwd(const void* r2, int r2_ld, const double* stats2, const void* r3, int r3_ld, int r3_coff,
                                        const double* stats3, float eps, int act, void* y, int y_ld, int B, int64_t V, int C,
                                        int dtype, void* stream) {
  if (!r2 || !stats2 || !r3 || !y || B <= 0 || V <= 0 || C <= 0) return B200SEG_EINVAL;
  cudaStream_t st = as_stream(stream);
  int64_t vpb = pick_vpb(V, B, C);
  dim3 grid(ceil_div(V, vpb), B);
  bool vok = vec_ok(r2, r2_ld, 0, C, 1) && vec_ok(r3, r3_ld, r3_coff, C, 1) && vec_ok(y, y_ld, 0, C, 1);
  if (!vok && C > kThreads) return B200SEG_EUNSUPPORTED;
  DISPATCH_TV(dtype, vok, {
    resout_fwd_kernel<T, VEC><<<grid, kThreads, 0, st>>>((const T*)r2, r2_ld, stats2, (const T*)r3, r3_ld, r3_coff, stats3, eps, act,
                                                        (T*)y, y_ld, V, C, vpb);
  })
  B200_CHECK_LAUNCH("resout_fwd_kernel");
  return B200SEG_OK;
}

extern "C" int b200seg_resblock_out_bwd_reduce(const void* dy, int dy_ld, const void* y, int y_ld, const void* r2, int r2_ld,
                                               const double* stats2, const void* r3, int r3_ld, int r3_coff,
                                               const double* stats3, float eps, int act, void* g, double* sums, int B,
                                               int64_t V, int C, int dtype, void* stream) {
  if (!dy || !y || !r2 || !stats2 || !r3 || !g || !sums || B <= 0 || V <= 0 || C <= 0) return B200SEG_EINVAL;
  cudaStream_t st = as_stream(stream);
  int64_t vpb = pick_vpb(V, B, C);
  dim3 grid(ceil_div(V, vpb), B);
  bool vok = vec_ok(dy, dy_ld, 0, C, 1) && vec_ok(y, y_ld, 0, C, 1) && vec_ok(r2, r2_ld, 0, C, 1) &&
             vec_ok(r3, r3_ld, r3_coff, C, 1) && vec_ok(g, C, 0, C, 1);
  if (!vok && C > kThreads) return B200SEG_EUNSUPPORTED;
  DISPATCH_TV(dtype, vok, {
    size_t sm = sizeof(float) * kThreads * 3 * VEC;
    resout_bwd_reduce_kernel<T, VEC><<<grid, kThreads, sm, st>>>((const T*)dy, dy_ld, (const T*)y, y_ld, (const T*)r2, r2_ld, stats2,
                                                                (const T*)r3, r3_ld, r3_coff, stats3, eps, act, (T*)g, sums, V, C, vpb);
  })
  B200_CHECK_LAUNCH("resout_bwd_reduce_kernel");
  return B200SEG_OK;
}
