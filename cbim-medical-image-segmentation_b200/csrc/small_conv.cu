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
  // block-level reduction first: thousands of warps hammering the same few hundred addresses with atomics
  // serialise in L2 (measured ~1 ms); one atomic per (block, output) instead
  __shared__ float s_acc[kWarpsPerBlock][32][MAXT];
  const int wid = threadIdx.x >> 5;
  for (int half = 0; half < (two ? 2 : 1); ++half) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < MAXT; ++q) s_acc[wid][lane][q] = half ? acc1[q] : acc0[q];
    __syncthreads();
    const int nco = min(32, a.Cout - 32 * half);
    for (int o = threadIdx.x; o < nco * MAXT; o += kWarpsPerBlock * 32) {
      const int co = o / MAXT, q = o % MAXT;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) sum += s_acc[w][co][q];
      atomicAdd(&a.dw[(int64_t)(co + 32 * half) * MAXT + q], sum);
    }
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
  constexpr int U = 4;                         // voxels in flight per warp iteration (memory-level parallelism)
  for (int64_t gv0 = v0; gv0 < v1; gv0 += U) {
    const int64_t b = gv0 / V;
    if (a.x_stats && (b != cur_b || (gv0 + U - 1) / V != b)) {
      // (re)load the sample's mean / rstd; a batch that straddles two samples is processed voxel by voxel below
      __syncwarp();
      for (int c = lane; c < a.Cin; c += 32) stats_to_mean_rstd(a.x_stats + (b * a.Cin + c) * 2, (double)V, a.eps, s_mean[c], s_rstd[c]);
      __syncwarp();
      cur_b = b;
    }
    const bool straddle = a.x_stats && ((gv0 + U - 1) / V != b) && (gv0 + U - 1 < v1);
    float g[U][MAXCO], xr[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t gv = gv0 + u;
      const bool in = gv < v1;
#pragma unroll
      for (int c = 0; c < MAXCO; ++c) g[u][c] = (in && c < a.Cout) ? Elem<T>::ld(dy + gv * a.dy_ld + a.dy_coff + c) : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ci = lane + 32 * k;
        xr[u][k] = (in && k < nci && ci < a.Cin) ? Elem<T>::ld(x + gv * a.x_ld + a.x_coff + ci) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t gv = gv0 + u;
      if (gv >= v1) break;
      if (straddle && gv / V != cur_b) {
        const int64_t b2 = gv / V;
        __syncwarp();
        for (int c = lane; c < a.Cin; c += 32) stats_to_mean_rstd(a.x_stats + (b2 * a.Cin + c) * 2, (double)V, a.eps, s_mean[c], s_rstd[c]);
        __syncwarp();
        cur_b = b2;
      }
#pragma unroll
      for (int c = 0; c < MAXCO; ++c) bacc[c] += g[u][c];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= nci) break;
        const int ci = lane + 32 * k;
        float xv = 0.f;
        if (ci < a.Cin) {
          xv = xr[u][k];
          if (a.x_stats) xv = (xv - s_mean[ci]) * s_rstd[ci];
          if (a.act == B200SEG_ACT_RELU) xv = fmaxf(xv, 0.f);
          xv = Elem<T>::round(xv);
        }
#pragma unroll
        for (int c = 0; c < MAXCO; ++c) acc[k][c] = fmaf(g[u][c], xv, acc[k][c]);
      }
    }
  }
  __shared__ float s_acc[kWarpsPerBlock][32][MAXCO];
  __shared__ float s_bacc[kWarpsPerBlock][MAXCO];
  const int wid = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < MAXCO; ++c) s_bacc[wid][c] = bacc[c];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= nci) break;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < MAXCO; ++c) s_acc[wid][lane][c] = acc[k][c];
    __syncthreads();
    const int nch = min(32, a.Cin - 32 * k);
    for (int o = threadIdx.x; o < nch * a.Cout; o += kWarpsPerBlock * 32) {
      const int c = o / nch, cl = o % nch;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) sum += s_acc[w][cl][c];
      atomicAdd(&a.dw[(int64_t)c * a.Cin + 32 * k + cl], sum);
    }
  }
  __syncthreads();
  if (a.dbias && threadIdx.x < a.Cout) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) sum += s_bacc[w][threadIdx.x];
    atomicAdd(&a.dbias[threadIdx.x], sum);
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

// ===================================================================== forward special cases
namespace {

__device__ __forceinline__ float col_sum16(float (&v)[16], int lane) {   // see conv_tc.cu::column_sum16
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float s = (lane & 16) ? v[i] : v[i + 8], k = (lane & 16) ? v[i + 8] : v[i]; v[i] = k + __shfl_xor_sync(0xffffffffu, s, 16); }
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float s = (lane & 8) ? v[i] : v[i + 4], k = (lane & 8) ? v[i + 4] : v[i]; v[i] = k + __shfl_xor_sync(0xffffffffu, s, 8); }
#pragma unroll
  for (int i = 0; i < 2; ++i) { const float s = (lane & 4) ? v[i] : v[i + 2], k = (lane & 4) ? v[i + 2] : v[i]; v[i] = k + __shfl_xor_sync(0xffffffffu, s, 4); }
  { const float s = (lane & 2) ? v[0] : v[1], k = (lane & 2) ? v[1] : v[0]; v[0] = k + __shfl_xor_sync(0xffffffffu, s, 2); }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- stem forward: y[v][0..31] = sum_tap w[tap][co] * x[v+tap]  (Cin == 1, Cout == 32), + IN sums of the stored y.
// One thread per voxel (64 B of output), weights in shared memory.  HBM-bound: writes 32 channels per voxel.
template <typename T, int KD, int KH, int KW>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(ConvArgs a) {
  constexpr int TAPS = KD * KH * KW;
  __shared__ float s_w[TAPS][32];
  __shared__ float s_red[8][32][2];
  const int64_t V = (int64_t)a.D * a.H * a.W;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < TAPS * 32; i += 256) s_w[i / 32][i % 32] = Elem<T>::ld((const T*)a.w + i);   // packed [tap][co][1]
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool valid = v < V;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  if (valid) {
    const int w0 = (int)(v % a.W); const int64_t t = v / a.W; const int h0 = (int)(t % a.H); const int d0 = (int)(t / a.H);
    const T* xb = (const T*)a.x + (int64_t)b * V * a.x_ld + a.x_coff;
#pragma unroll
    for (int zd = 0; zd < KD; ++zd)
#pragma unroll
      for (int zh = 0; zh < KH; ++zh)
#pragma unroll
        for (int zw = 0; zw < KW; ++zw) {
          const int d = d0 + zd - KD / 2, h = h0 + zh - KH / 2, w = w0 + zw - KW / 2;
          float xv = 0.f;
          if ((unsigned)d < (unsigned)a.D && (unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W)
            xv = Elem<T>::ld(xb + (((int64_t)d * a.H + h) * a.W + w) * a.x_ld);
          const float* wt = s_w[(zd * KH + zh) * KW + zw];
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[c] = fmaf(xv, wt[c], acc[c]);
        }
  }
  float sq[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { acc[c] = valid ? Elem<T>::round(acc[c]) : 0.f; sq[c] = acc[c] * acc[c]; }
  if (valid) {
    T* yp = (T*)a.y + ((int64_t)b * V + v) * a.y_ld + a.y_coff;
#pragma unroll
    for (int c = 0; c < 32; c += 8) st8<T>(yp + c, reinterpret_cast<const float(&)[8]>(acc[c]));
  }
  if (a.y_stats) {
    const float u0 = col_sum16(reinterpret_cast<float(&)[16]>(acc[0]), lane), u1 = col_sum16(reinterpret_cast<float(&)[16]>(acc[16]), lane);
    const float q0 = col_sum16(reinterpret_cast<float(&)[16]>(sq[0]), lane), q1 = col_sum16(reinterpret_cast<float(&)[16]>(sq[16]), lane);
    if ((lane & 1) == 0) {
      const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      s_red[wid][col][0] = u0; s_red[wid][col][1] = q0; s_red[wid][16 + col][0] = u1; s_red[wid][16 + col][1] = q1;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int c = threadIdx.x >> 1, k = threadIdx.x & 1;
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += (double)s_red[w][c][k];
      atomicAdd(&a.y_stats[((int64_t)b * 32 + c) * 2 + k], s);
    }
  }
}

// ---- 1x1x1 conv with few channels on one side (the classifier head forward 32->4 and its data-gradient 4->32):
// one thread per voxel, weights + bias in shared memory, vector loads/stores.  Cin, Cout multiples of 4, <= 64.
template <typename T, int CIN>
__global__ void __launch_bounds__(256)
pointwise_small_kernel(ConvArgs a) {
  extern __shared__ float s_wb[];                 // [Cout][Cin] then [Cout] bias
  const int n = a.Cout * a.Cin;
  for (int i = threadIdx.x; i < n; i += 256) s_wb[i] = Elem<T>::ld((const T*)a.w + i);     // packed [1][Cout][Cin]
  for (int i = threadIdx.x; i < a.Cout; i += 256) s_wb[n + i] = a.bias ? a.bias[i] : 0.f;
  __syncthreads();
  const int64_t total = (int64_t)a.B * a.D * a.H * a.W;
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= total) return;
  const T* xp = (const T*)a.x + v * a.x_ld + a.x_coff;
  T* yp = (T*)a.y + v * a.y_ld + a.y_coff;
  float xin[CIN];
#pragma unroll
  for (int c = 0; c < CIN; c += 4) {
    {
      if constexpr (sizeof(T) == 2) {
        const uint2 u = *reinterpret_cast<const uint2*>(xp + c);
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        xin[c] = f0.x; xin[c + 1] = f0.y; xin[c + 2] = f1.x; xin[c + 3] = f1.y;
      } else {
        const float4 f = *reinterpret_cast<const float4*>(xp + c);
        xin[c] = f.x; xin[c + 1] = f.y; xin[c + 2] = f.z; xin[c + 3] = f.w;
      }
    }
  }
  for (int co = 0; co < a.Cout; co += 4) {
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float s = s_wb[n + co + q];
      const float4* wr = reinterpret_cast<const float4*>(s_wb + (co + q) * CIN);
#pragma unroll
      for (int c = 0; c < CIN / 4; ++c) {
        const float4 w4 = wr[c];
        s = fmaf(xin[4 * c], w4.x, s); s = fmaf(xin[4 * c + 1], w4.y, s); s = fmaf(xin[4 * c + 2], w4.z, s); s = fmaf(xin[4 * c + 3], w4.w, s);
      }
      o[q] = s;
    }
    if constexpr (sizeof(T) == 2) {
      uint2 u;
      *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(o[0], o[1]);
      *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(o[2], o[3]);
      *reinterpret_cast<uint2*>(yp + co) = u;
    } else {
      *reinterpret_cast<float4*>(yp + co) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace

int conv3d_fwd_small(const ConvArgs& a, int dtype, cudaStream_t st) {
  const int64_t V = (int64_t)a.D * a.H * a.W;
  const bool plain = !a.x_stats && !a.act && !a.res && !a.gx;
  const bool k133 = (a.kd == 1 && a.kh == 3 && a.kw == 3), k333 = (a.kd == 3 && a.kh == 3 && a.kw == 3);
  if (plain && !a.bias && a.Cin == 1 && a.Cout == 32 && (k133 || k333) && (a.y_ld % 8 == 0) && (a.y_coff % 8 == 0)) {
    dim3 grid(ceil_div(V, 256), a.B);
    if (dtype == B200SEG_F16) { if (k133) stem_fwd_kernel<__half, 1, 3, 3><<<grid, 256, 0, st>>>(a); else stem_fwd_kernel<__half, 3, 3, 3><<<grid, 256, 0, st>>>(a); }
    else { if (k133) stem_fwd_kernel<float, 1, 3, 3><<<grid, 256, 0, st>>>(a); else stem_fwd_kernel<float, 3, 3, 3><<<grid, 256, 0, st>>>(a); }
    B200_CHECK_LAUNCH("stem_fwd_kernel");
    return B200SEG_OK;
  }
  const bool k111 = (a.kd == 1 && a.kh == 1 && a.kw == 1);
  if (plain && !a.y_stats && k111 && (a.Cin == 4 || a.Cin == 8 || a.Cin == 16 || a.Cin == 32 || a.Cin == 64) &&
      a.Cout % 4 == 0 && a.Cout <= 64 && a.x_ld % 4 == 0 && a.x_coff % 4 == 0 && a.y_ld % 4 == 0 && a.y_coff % 4 == 0) {
    const int64_t total = (int64_t)a.B * V;
    const size_t sm = sizeof(float) * ((size_t)a.Cout * a.Cin + a.Cout);
    const int grid = ceil_div(total, 256);
#define PW(TT, CI) pointwise_small_kernel<TT, CI><<<grid, 256, sm, st>>>(a)
#define PWT(TT) switch (a.Cin) { case 4: PW(TT, 4); break; case 8: PW(TT, 8); break; case 16: PW(TT, 16); break; case 32: PW(TT, 32); break; default: PW(TT, 64); }
    if (dtype == B200SEG_F16) { PWT(__half) } else { PWT(float) }
#undef PWT
#undef PW
    B200_CHECK_LAUNCH("pointwise_small_kernel");
    return B200SEG_OK;
  }
  return B200SEG_EUNSUPPORTED;
}
