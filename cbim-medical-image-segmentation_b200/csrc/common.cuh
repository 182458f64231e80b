// common.cuh — shared device/host helpers for libb200seg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/b200seg.h"

#define B200SEG_NUM_SMS 148

// ---- host-side error plumbing -------------------------------------------------
extern thread_local char g_b200seg_cuda_err[256];
int b200seg_record_cuda(cudaError_t e, const char* what);

#define B200_CHECK_LAUNCH(what)                                   \
  do {                                                            \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return b200seg_record_cuda(_e, what);  \
  } while (0)

#define B200_CUDA(call)                                           \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return b200seg_record_cuda(_e, #call); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- dtype helpers ------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float round(float v) { return v; }
};
template <> struct Elem<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  static __device__ __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
};

// 8 consecutive channels (one 16-byte chunk of fp16 / two of fp32) as floats
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__half>(const __half* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<__half>(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- activations (B200SEG_ACT_*) ------------------------------------------------
#define B200SEG_LRELU_SLOPE 0.01f
__device__ __forceinline__ float act_apply(float h, int act) {
  return act == B200SEG_ACT_RELU ? fmaxf(h, 0.f) : (act == B200SEG_ACT_LRELU ? (h > 0.f ? h : B200SEG_LRELU_SLOPE * h) : h);
}
// Branch-free forms for unrolled per-element loops (a run-time `act` switch inside such a loop compiles to uniform
// branches per ELEMENT, which serialises the elements' dependency chains — measured 1800 vs ~600 cycles per 16-column
// epilogue chunk): every supported activation is  h > 0 ? h : slope * h  with slope = 0 (ReLU), 0.01 (LeakyReLU), 1 (none).
__host__ __device__ __forceinline__ float act_slope(int act) {
  return act == B200SEG_ACT_RELU ? 0.f : (act == B200SEG_ACT_LRELU ? B200SEG_LRELU_SLOPE : 1.f);
}
__device__ __forceinline__ float act_apply_s(float h, float slope) { return h > 0.f ? h : slope * h; }
__device__ __forceinline__ float act_grad_s(float h, float slope) { return h > 0.f ? 1.f : slope; }
// derivative of the activation at pre-activation h, as the factor the incoming gradient is multiplied with
__device__ __forceinline__ float act_grad(float h, int act) {
  return act == B200SEG_ACT_RELU ? (h > 0.f ? 1.f : 0.f) : (act == B200SEG_ACT_LRELU ? (h > 0.f ? 1.f : B200SEG_LRELU_SLOPE) : 1.f);
}

// ---- reductions ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// mean / rstd of one (b,c) from the accumulated {sum, sumsq} (biased variance,
// as torch.nn.functional.instance_norm -> native_batch_norm in training mode).
__device__ __forceinline__ void stats_to_mean_rstd(const double* st, double n, float eps,
                                                   float& mean, float& rstd) {
  double m = st[0] / n;
  double var = st[1] / n - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
