// inference.cu — device side of the evaluation consumers of net(x):
//   * sliding-window inference, inference/inference3d.py:28-92: per window  prob[region] += softmax(logits),
//     counter[region] += 1  (one fused kernel instead of softmax + two slice-adds), then  prob /= counter
//     (+ optional argmax label map in the same pass);
//   * one-hot Dice metric, metric/utils.py:62-82 (calculate_dice): per class intersection = |pred==c & target==c| and
//     summ = |pred==c| + |target==c| in one pass over the two label maps (the reference scatters two [N,C] masks).
// All HBM-bound elementwise / reduction kernels; logits may be fp16 or fp32 with arbitrary (voxel, channel) strides
// (the models hand out channels-last views), probabilities are fp32 NCDHW as the reference returns them.
#include "common.cuh"

namespace {

constexpr int kMaxClasses = 64;

// prob[b][c][d0+d][h0+h][w0+w] += softmax_c(logits[b][v][c]); counter[b][d0+d][h0+h][w0+w] += 1
template <typename T>
__global__ void softmax_accumulate_kernel(const T* __restrict__ logits, int64_t sb, int64_t sv, int64_t sc, float* __restrict__ prob,
                                          float* __restrict__ counter, int B, int C, int wd, int wh, int ww, int D, int H, int W,
                                          int d0, int h0, int w0) {
  const int64_t Vw = (int64_t)wd * wh * ww, total = (int64_t)B * Vw, V = (int64_t)D * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / Vw, v = i - b * Vw;
    const int w = (int)(v % ww); const int64_t t = v / ww; const int h = (int)(t % wh); const int d = (int)(t / wh);
    const T* lp = logits + b * sb + v * sv;
    float x[kMaxClasses];
    float m = -INFINITY;
#pragma unroll 4
    for (int c = 0; c < C; ++c) { x[c] = Elem<T>::ld(lp + c * sc); m = fmaxf(m, x[c]); }
    float s = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) { x[c] = __expf(x[c] - m); s += x[c]; }
    const float inv = 1.f / s;
    const int64_t o = ((int64_t)(d0 + d) * H + (h0 + h)) * W + (w0 + w);
    float* pp = prob + b * C * V + o;
#pragma unroll 4
    for (int c = 0; c < C; ++c) pp[(int64_t)c * V] += x[c] * inv;
    counter[b * V + o] += 1.f;
  }
}

// prob /= counter; optionally label[b][v] = argmax_c prob
__global__ void normalize_argmax_kernel(float* __restrict__ prob, const float* __restrict__ counter, uint8_t* __restrict__ label, int B,
                                        int C, int64_t V) {
  const int64_t total = (int64_t)B * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / V, v = i - b * V;
    const float inv = 1.f / counter[i];
    float* pp = prob + b * C * V + v;
    float best = -INFINITY; int arg = 0;
    for (int c = 0; c < C; ++c) {
      const float p = pp[(int64_t)c * V] * inv;
      pp[(int64_t)c * V] = p;
      if (p > best) { best = p; arg = c; }        // first maximum, like torch.argmax
    }
    if (label) label[i] = (uint8_t)arg;
  }
}

// sums[c] = {intersection, |pred == c| + |target == c|}; labels as uint8 / int64
template <typename TP, typename TT>
__global__ void __launch_bounds__(256) dice_metric_kernel(const TP* __restrict__ pred, const TT* __restrict__ target, int64_t N, int C,
                                                          unsigned long long* __restrict__ out) {
  __shared__ unsigned int s_i[kMaxClasses], s_s[kMaxClasses];
  for (int c = threadIdx.x; c < C; c += 256) { s_i[c] = 0; s_s[c] = 0; }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += (int64_t)gridDim.x * 256) {
    const int p = (int)pred[i], t = (int)target[i];
    if ((unsigned)p < (unsigned)C) atomicAdd(&s_s[p], 1u);
    if ((unsigned)t < (unsigned)C) atomicAdd(&s_s[t], 1u);
    if (p == t && (unsigned)p < (unsigned)C) atomicAdd(&s_i[p], 1u);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    if (s_i[c]) atomicAdd(&out[2 * c], (unsigned long long)s_i[c]);
    if (s_s[c]) atomicAdd(&out[2 * c + 1], (unsigned long long)s_s[c]);
  }
}

inline int grid_for(int64_t n, int th) {
  int64_t g = (n + th - 1) / th;
  const int64_t cap = (int64_t)B200SEG_NUM_SMS * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

// logits element (b, window voxel v, class c) at logits[b*sb + v*sv + c*sc]; prob fp32 [B][C][D][H][W], counter fp32 [B][D][H][W]
extern "C" int b200seg_softmax_accumulate(const void* logits, int dtype, int64_t sb, int64_t sv, int64_t sc, float* prob,
                                          float* counter, int B, int C, int wd, int wh, int ww, int D, int H, int W, int d0,
                                          int h0, int w0, void* stream) {
  if (!logits || !prob || !counter || B <= 0 || C <= 0 || wd <= 0 || wh <= 0 || ww <= 0) return B200SEG_EINVAL;
  if (C > kMaxClasses) return B200SEG_EUNSUPPORTED;
  if (d0 < 0 || h0 < 0 || w0 < 0 || d0 + wd > D || h0 + wh > H || w0 + ww > W) return B200SEG_EINVAL;
  const int64_t total = (int64_t)B * wd * wh * ww;
  cudaStream_t st = as_stream(stream);
  if (dtype == B200SEG_F16)
    softmax_accumulate_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>((const __half*)logits, sb, sv, sc, prob, counter, B, C, wd, wh, ww, D, H, W, d0, h0, w0);
  else if (dtype == B200SEG_F32)
    softmax_accumulate_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)logits, sb, sv, sc, prob, counter, B, C, wd, wh, ww, D, H, W, d0, h0, w0);
  else return B200SEG_EINVAL;
  B200_CHECK_LAUNCH("softmax_accumulate_kernel");
  return B200SEG_OK;
}

extern "C" int b200seg_normalize_argmax(float* prob, const float* counter, uint8_t* label, int B, int C, int64_t V, void* stream) {
  if (!prob || !counter || B <= 0 || C <= 0 || V <= 0) return B200SEG_EINVAL;
  if (label && C > 256) return B200SEG_EUNSUPPORTED;
  normalize_argmax_kernel<<<grid_for((int64_t)B * V, 256), 256, 0, as_stream(stream)>>>(prob, counter, label, B, C, V);
  B200_CHECK_LAUNCH("normalize_argmax_kernel");
  return B200SEG_OK;
}

// out: uint64 [C][2] = {intersection, summ}, zeroed by the caller; label_bytes: 1 (uint8) or 8 (int64) per map
extern "C" int b200seg_dice_metric(const void* pred, int pred_bytes, const void* target, int target_bytes, int64_t N, int C,
                                   unsigned long long* out, void* stream) {
  if (!pred || !target || !out || N <= 0 || C <= 0) return B200SEG_EINVAL;
  if (C > kMaxClasses) return B200SEG_EUNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int grid = grid_for(N, 256);
#define DM(TP, TT) dice_metric_kernel<TP, TT><<<grid, 256, 0, st>>>((const TP*)pred, (const TT*)target, N, C, out)
  if (pred_bytes == 1 && target_bytes == 1) DM(uint8_t, uint8_t);
  else if (pred_bytes == 1 && target_bytes == 8) DM(uint8_t, int64_t);
  else if (pred_bytes == 8 && target_bytes == 1) DM(int64_t, uint8_t);
  else if (pred_bytes == 8 && target_bytes == 8) DM(int64_t, int64_t);
  else return B200SEG_EINVAL;
#undef DM
  B200_CHECK_LAUNCH("dice_metric_kernel");
  return B200SEG_OK;
}
