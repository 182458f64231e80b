// small_conv.cu — HBM-bound special cases of the weight gradient that are not dense contractions
// (SURVEY.md §8d "HBM bandwidth" rows): the Cin=1 stem (unet_utils.py:14) and the 1x1x1 classifier head
// with a handful of output channels (`outc`, unet.py:47).  One pass over dy / x, warp-per-voxel-run with a
// lane per channel, fp32 register accumulators, one atomicAdd per (warp, output) at the end.
#include "common.cuh"
#include "conv_args.h"

namespace {

constexpr int kWarpsPerBlock = 8;

// ---- stem: dW[co][tap] = sum_v dy[v][co] * x[v + tap],  Cin == 1, Cout <= 64, taps <= 27
template <typename T, int KD, int KH, int KW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
wgrad_cin1_kernel(WgradArgs a, int64_t vox_per_warp) {
  constexpr int MAXT = KD * KH * KW;
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t V = (int64_t)a.D * a.H * a.W, total = (int64_t)a.B * V;
  constexpr int pd = KD / 2, ph = KH / 2, pw = KW / 2;
  const T* x = (const T*)a.x; const T* dy = (const T*)a.dy;
  float acc0[MAXT], acc1[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t) { acc0[t] = 0.f; acc1[t] = 0.f; }
  const bool two = a.Cout > 32;
  int64_t v0 = warp_id * vox_per_warp, v1 = v0 + vox_per_warp; if (v1 > total) v1 = total;
  for (int64_t gv = v0; gv < v1; ++gv) {
    const int64_t b = gv / V, v = gv - b * V;
    const int w = (int)(v % a.W); const int64_t t2 = v / a.W; const int h = (int)(t2 % a.H); const int d = (int)(t2 / a.H);
    const float g0 = lane < a.Cout ? Elem<T>::ld(dy + gv * a.dy_ld + a.dy_coff + lane) : 0.f;
    const float g1 = (two && lane + 32 < a.Cout) ? Elem<T>::ld(dy + gv * a.dy_ld + a.dy_coff + lane + 32) : 0.f;
    const T* xb = x + b * V * a.x_ld + a.x_coff;
#pragma unroll
    for (int zd = 0; zd < KD; ++zd) {
      const int dd = d + zd - pd;
#pragma unroll
      for (int zh = 0; zh < KH; ++zh) {
        const int hh = h + zh - ph;
#pragma unroll
        for (int zw = 0; zw < KW; ++zw) {
          const int ww = w + zw - pw;
          float xv = 0.f;          // warp-uniform address: one broadcast transaction
          if ((unsigned)dd < (unsigned)a.D && (unsigned)hh < (unsigned)a.H && (unsigned)ww < (unsigned)a.W)
            xv = Elem<T>::ld(xb + (((int64_t)dd * a.H + hh) * a.W + ww) * a.x_ld);
          constexpr int dummy = 0; (void)dummy;
          const int t = (zd * KH + zh) * KW + zw;
          acc0[t] = fmaf(g0, xv, acc0[t]);
          acc1[t] = fmaf(g1, xv, acc1[t]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MAXT; ++q) {
    if (lane < a.Cout) atomicAdd(&a.dw[(int64_t)lane * MAXT + q], acc0[q]);
    if (two && lane + 32 < a.Cout) atomicAdd(&a.dw[(int64_t)(lane + 32) * MAXT + q], acc1[q]);
  }
}

// ---- head: dW[co][ci] = sum_v dy[v][co] * a[v][ci], 1x1x1, Cout <= 16, Cin <= 128; dbias[co] = sum_v dy[v][co]
template <typename T, int MAXCO>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
wgrad_head_kernel(WgradArgs a, int64_t vox_per_warp) {
  __shared__ float s_mean_all[kWarpsPerBlock][128], s_rstd_all[kWarpsPerBlock][128];   // per warp: runs may sit in different samples
  float* s_mean = s_mean_all[threadIdx.x >> 5];
  float* s_rstd = s_rstd_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t V = (int64_t)a.D * a.H * a.W, total = (int64_t)a.B * V;
  const T* x = (const T*)a.x; const T* dy = (const T*)a.dy;
  const int nci = (a.Cin + 31) / 32;        // channels per lane: lane, lane+32, ...
  float acc[4][MAXCO];
  float bacc[MAXCO];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < MAXCO; ++c) acc[k][c] = 0.f;
#pragma unroll
  for (int c = 0; c < MAXCO; ++c) bacc[c] = 0.f;
  int64_t v0 = warp_id * vox_per_warp, v1 = v0 + vox_per_warp; if (v1 > total) v1 = total;
  int64_t cur_b = -1;
  for (int64_t gv = v0; gv < v1; ++gv) {
    const int64_t b = gv / V;
    if (a.x_stats && b != cur_b) {          // (block-uniform in practice: a warp's run rarely crosses a sample)
      __syncwarp();
      for (int c = lane; c < a.Cin; c += 32) stats_to_mean_rstd(a.x_stats + (b * a.Cin + c) * 2, (double)V, a.eps, s_mean[c], s_rstd[c]);
      __syncwarp();
      cur_b = b;
    }
    float g[MAXCO];
#pragma unroll
    for (int c = 0; c < MAXCO; ++c) g[c] = c < a.Cout ? Elem<T>::ld(dy + gv * a.dy_ld + a.dy_coff + c) : 0.f;
#pragma unroll
    for (int c = 0; c < MAXCO; ++c) bacc[c] += g[c];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= nci) break;
      const int ci = lane + 32 * k;
      float xv = 0.f;
      if (ci < a.Cin) {
        xv = Elem<T>::ld(x + gv * a.x_ld + a.x_coff + ci);
        if (a.x_stats) xv = (xv - s_mean[ci]) * s_rstd[ci];
        if (a.act == B200SEG_ACT_RELU) xv = fmaxf(xv, 0.f);
        xv = Elem<T>::round(xv);
      }
#pragma unroll
      for (int c = 0; c < MAXCO; ++c) acc[k][c] = fmaf(g[c], xv, acc[k][c]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ci = lane + 32 * k;
    if (k < nci && ci < a.Cin) {
#pragma unroll
      for (int c = 0; c < MAXCO; ++c) if (c < a.Cout) atomicAdd(&a.dw[(int64_t)c * a.Cin + ci], acc[k][c]);
    }
  }
  if (a.dbias && lane == 0) {
#pragma unroll
    for (int c = 0; c < MAXCO; ++c) if (c < a.Cout) atomicAdd(&a.dbias[c], bacc[c]);
  }
}

}  // namespace

// returns B200SEG_EUNSUPPORTED when the shape is not one of the special cases
int conv3d_wgrad_small(const WgradArgs& a, int dtype, cudaStream_t st) {
  const int taps = a.kd * a.kh * a.kw;
  const int64_t total = (int64_t)a.B * a.D * a.H * a.W;
  const int nwarps_target = B200SEG_NUM_SMS * kWarpsPerBlock * 4;
  int64_t vpw = (total + nwarps_target - 1) / nwarps_target; if (vpw < 64) vpw = 64;
  const int64_t nwarps = (total + vpw - 1) / vpw;
  const int grid = (int)((nwarps + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const bool k133 = (a.kd == 1 && a.kh == 3 && a.kw == 3), k333 = (a.kd == 3 && a.kh == 3 && a.kw == 3);
  if (a.Cin == 1 && a.Cout <= 64 && (k133 || k333) && !a.x_stats && !a.act && !a.dbias) {
    const int th = kWarpsPerBlock * 32;
    if (dtype == B200SEG_F16) {
      if (k133) wgrad_cin1_kernel<__half, 1, 3, 3><<<grid, th, 0, st>>>(a, vpw);
      else wgrad_cin1_kernel<__half, 3, 3, 3><<<grid, th, 0, st>>>(a, vpw);
    } else {
      if (k133) wgrad_cin1_kernel<float, 1, 3, 3><<<grid, th, 0, st>>>(a, vpw);
      else wgrad_cin1_kernel<float, 3, 3, 3><<<grid, th, 0, st>>>(a, vpw);
    }
    B200_CHECK_LAUNCH("wgrad_cin1_kernel");
    return B200SEG_OK;
  }
  if (taps == 1 && a.Cout <= 16 && a.Cin <= 128) {
    if (dtype == B200SEG_F16) {
      if (a.Cout <= 4) wgrad_head_kernel<__half, 4><<<grid, kWarpsPerBlock * 32, 0, st>>>(a, vpw);
      else wgrad_head_kernel<__half, 16><<<grid, kWarpsPerBlock * 32, 0, st>>>(a, vpw);
    } else {
      if (a.Cout <= 4) wgrad_head_kernel<float, 4><<<grid, kWarpsPerBlock * 32, 0, st>>>(a, vpw);
      else wgrad_head_kernel<float, 16><<<grid, kWarpsPerBlock * 32, 0, st>>>(a, vpw);
    }
    B200_CHECK_LAUNCH("wgrad_head_kernel");
    return B200SEG_OK;
  }
  return B200SEG_EUNSUPPORTED;
}
