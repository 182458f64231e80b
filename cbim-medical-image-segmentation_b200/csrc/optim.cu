// optim.cu — the optimiser tail of the training step as two multi-tensor kernels:
//   GradScaler unscale + inf/nan check  (train_ddp.py:193-195: scaler.step / scaler.update)
//   AdamW(lr, betas, eps=1e-5, weight_decay) (training/utils.py:8-14)  +  EMA update (training/utils.py:98-105)
// Stock PyTorch runs this as: one non-finite check pass over every gradient, the fused AdamW pass, and two foreach
// passes for the EMA (mul_, add_) — 52 bytes / parameter in four kernel families, plus a per-step loss.item() sync.
// Here: kernel 1 reads every gradient once (4 B/param) and raises found_inf; kernel 2 reads g, p, m, v, ema and
// writes p, m, v, ema (36 B/param), skipping everything on device when found_inf is set (GradScaler semantics).
// HBM-bound: algorithmic bytes = 40 B/param.  All tensors fp32, contiguous.
#include "common.cuh"

namespace {

constexpr int kOptChunk = 8192;      // elements per block

// table row: {grad, param, exp_avg, exp_avg_sq, ema (0 = none), numel}; chunk row: {tensor, first element}
__global__ void __launch_bounds__(256) nonfinite_check_kernel(const int64_t* __restrict__ tab, const int64_t* __restrict__ chunks,
                                                              float* __restrict__ found_inf) {
  const int64_t* c = chunks + 2 * (int64_t)blockIdx.x;
  const int64_t* t = tab + 6 * c[0];
  const float* g = reinterpret_cast<const float*>(t[0]);
  const int64_t n = t[5], i0 = c[1], i1 = i0 + kOptChunk < n ? i0 + kOptChunk : n;
  bool bad = false;
  if (((i0 | n) & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    for (int64_t i = i0 + 4 * threadIdx.x; i < i1; i += 4 * 256) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      bad |= !isfinite(v.x) | !isfinite(v.y) | !isfinite(v.z) | !isfinite(v.w);
    }
  } else {
    for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) bad |= !isfinite(g[i]);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) *found_inf = 1.f;
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, ema_alpha;
  const float* scale;          // device scalar: the loss scale the gradients carry (nullable = 1)
  const float* found_inf;      // device scalar (nullable = 0): non-zero skips the whole update
  const float* step;           // device scalar: number of optimiser steps ALREADY applied (skipped steps do not count)
};

struct AdamStep { float bc1, bc2s, alpha; };
__device__ __forceinline__ void adamw_one(float& p, float& m, float& v, float g, const AdamArgs& a, const AdamStep& s) {
  // torch.optim.AdamW (decoupled weight decay, no amsgrad, no maximize)
  p *= 1.f - a.lr * a.weight_decay;
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float denom = sqrtf(v) / s.bc2s + a.eps;
  p -= (a.lr / s.bc1) * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(const int64_t* __restrict__ tab, const int64_t* __restrict__ chunks, AdamArgs a) {
  const bool skip = a.found_inf && *a.found_inf != 0.f;     // GradScaler: no optimiser step; the EMA still runs
  const int64_t* c = chunks + 2 * (int64_t)blockIdx.x;
  const int64_t* t = tab + 6 * c[0];
  const float* g = reinterpret_cast<const float*>(t[0]);
  float* p = reinterpret_cast<float*>(t[1]);
  float* m = reinterpret_cast<float*>(t[2]);
  float* v = reinterpret_cast<float*>(t[3]);
  float* e = reinterpret_cast<float*>(t[4]);
  const int64_t n = t[5], i0 = c[1], i1 = i0 + kOptChunk < n ? i0 + kOptChunk : n;
  const float is = a.scale ? 1.f / *a.scale : 1.f;
  // bias corrections from the DEVICE step counter, so a step skipped for a non-finite gradient does not advance them
  // (torch's fused AdamW keeps its step tensors on the device for the same reason).  The EMA coefficient
  // alpha = min(1 - 1/(iteration + 1), ema_alpha) (training/utils.py:100) counts ITERATIONS: the host passes it.
  const float t1 = (a.step ? *a.step : 0.f) + 1.f;
  AdamStep s;
  s.bc1 = 1.f - powf(a.beta1, t1);
  s.bc2s = sqrtf(1.f - powf(a.beta2, t1));
  s.alpha = a.ema_alpha;
  const bool vec = ((i0 | n) & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) |
                                            reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(e)) & 15) == 0;
  if (skip && !e) return;
  if (vec) {
    for (int64_t i = i0 + 4 * threadIdx.x; i < i1; i += 4 * 256) {
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 p4 = *reinterpret_cast<float4*>(p + i), m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
      if (!skip) {
        adamw_one(p4.x, m4.x, v4.x, g4.x * is, a, s); adamw_one(p4.y, m4.y, v4.y, g4.y * is, a, s);
        adamw_one(p4.z, m4.z, v4.z, g4.z * is, a, s); adamw_one(p4.w, m4.w, v4.w, g4.w * is, a, s);
        *reinterpret_cast<float4*>(p + i) = p4; *reinterpret_cast<float4*>(m + i) = m4; *reinterpret_cast<float4*>(v + i) = v4;
      }
      if (e) {
        float4 e4 = *reinterpret_cast<float4*>(e + i);
        e4.x = s.alpha * e4.x + (1.f - s.alpha) * p4.x; e4.y = s.alpha * e4.y + (1.f - s.alpha) * p4.y;
        e4.z = s.alpha * e4.z + (1.f - s.alpha) * p4.z; e4.w = s.alpha * e4.w + (1.f - s.alpha) * p4.w;
        *reinterpret_cast<float4*>(e + i) = e4;
      }
    }
  } else {
    for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
      float pp = p[i], mm = m[i], vv = v[i];
      if (!skip) {
        adamw_one(pp, mm, vv, g[i] * is, a, s);
        p[i] = pp; m[i] = mm; v[i] = vv;
      }
      if (e) e[i] = s.alpha * e[i] + (1.f - s.alpha) * pp;
    }
  }
}

}  // namespace

extern "C" int b200seg_optim_chunk_elems(void) { return kOptChunk; }

// found_inf (device float, caller zeroes it) is raised when any gradient element is inf / nan
extern "C" int b200seg_grads_nonfinite(const int64_t* table_dev, const int64_t* chunks_dev, int nchunks, float* found_inf, void* stream) {
  if (nchunks == 0) return B200SEG_OK;
  if (!table_dev || !chunks_dev || !found_inf || nchunks < 0) return B200SEG_EINVAL;
  nonfinite_check_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(table_dev, chunks_dev, found_inf);
  B200_CHECK_LAUNCH("nonfinite_check_kernel");
  return B200SEG_OK;
}

__global__ void step_advance_kernel(float* step, const float* found_inf) {
  if (!(found_inf && *found_inf != 0.f)) *step += 1.f;
}

// One AdamW step (+ EMA of the updated parameters with the coefficient `ema_alpha` the host computed for this
// iteration) over every tensor of the table.  `step_dev` (device float) counts the APPLIED steps and is advanced here; `scale` /
// `found_inf` are the GradScaler's device scalars (nullable for a plain fp32 step).
extern "C" int b200seg_adamw_ema_step(const int64_t* table_dev, const int64_t* chunks_dev, int nchunks, float lr, float beta1,
                                      float beta2, float eps, float weight_decay, float ema_alpha, float* step_dev,
                                      const float* scale, const float* found_inf, void* stream) {
  if (nchunks == 0) return B200SEG_OK;
  if (!table_dev || !chunks_dev || nchunks < 0 || !step_dev) return B200SEG_EINVAL;
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.ema_alpha = ema_alpha;
  a.scale = scale; a.found_inf = found_inf; a.step = step_dev;
  adamw_ema_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(table_dev, chunks_dev, a);
  B200_CHECK_LAUNCH("adamw_ema_kernel");
  step_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev, found_inf);
  B200_CHECK_LAUNCH("step_advance_kernel");
  return B200SEG_OK;
}
