// api.cu — C-ABI entry points that dispatch between the CUDA-core and tcgen05 conv paths, plus
// version / error plumbing.  No CPU fallback exists anywhere in this library: an unsupported
// shape or a non-sm_100 device is a hard error (SURVEY.md §8b "Errors").
#include "common.cuh"
#include "conv_args.h"
#include <stdio.h>
#include <string.h>

thread_local char g_b200seg_cuda_err[256] = "";

int b200seg_record_cuda(cudaError_t e, const char* what) {
  snprintf(g_b200seg_cuda_err, sizeof(g_b200seg_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return B200SEG_ECUDA;
}

extern "C" int b200seg_version(void) { return B200SEG_VERSION; }

extern "C" const char* b200seg_strerror(int code) {
  switch (code) {
    case B200SEG_OK: return "ok";
    case B200SEG_EINVAL: return "invalid argument";
    case B200SEG_EUNSUPPORTED: return "shape/dtype not supported by the requested algorithm";
    case B200SEG_ECUDA: return "CUDA runtime error";
    case B200SEG_ENODEVICE: return "device is not sm_100 (B200); this library has no fallback path";
    default: return "unknown error";
  }
}

extern "C" const char* b200seg_last_cuda_error(void) { return g_b200seg_cuda_err; }

extern "C" int b200seg_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { b200seg_record_cuda(e, "cudaGetDevice"); return B200SEG_ENODEVICE; }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { b200seg_record_cuda(e, "cudaDeviceGetAttribute"); return B200SEG_ENODEVICE; }
  return major == 10 ? B200SEG_OK : B200SEG_ENODEVICE;
}

extern "C" int b200seg_conv3d_algo(int Cin, int Cout, int kd, int kh, int kw, int dtype, int B) {
  // (outputs wider than 2048 channels per batch run on the tensor cores only without fused statistics, conv_tc.cu)
  if (conv3d_tc_shape_ok(Cin, Cout, kd, kh, kw, dtype) && B * Cin <= 4096 && B * Cout <= 8192) return B200SEG_ALGO_TC;
  return B200SEG_ALGO_DIRECT;
}

extern "C" int b200seg_conv3d_fwd(const void* x, int x_ld, int x_coff, const double* x_stats, float eps, int act,
                                  const void* w_packed, const float* bias, const void* residual, int r_ld,
                                  int r_coff, void* y, int y_ld, int y_coff, double* y_stats,
                                  const void* dgrad_x, int dx_ld, int dx_coff, const double* dgrad_stats,
                                  float dgrad_eps, int dgrad_act, int B, int D, int H, int W, int Cin, int Cout,
                                  int kd, int kh, int kw, int dtype, int algo, void* stream) {
  if (!x || !w_packed || !y || B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return B200SEG_EINVAL;
  if (kd <= 0 || kh <= 0 || kw <= 0 || !(kd & 1) || !(kh & 1) || !(kw & 1)) return B200SEG_EUNSUPPORTED;
  if (dtype != B200SEG_F16 && dtype != B200SEG_F32) return B200SEG_EINVAL;
  if (dgrad_x && !dgrad_stats) return B200SEG_EINVAL;
  if (dgrad_x && (residual || bias)) return B200SEG_EINVAL;
  ConvArgs a{x, x_ld, x_coff, x_stats, eps, act, w_packed, bias, residual, r_ld, r_coff, y, y_ld, y_coff, y_stats,
             dgrad_x, dx_ld, dx_coff, dgrad_stats, dgrad_eps, dgrad_act, B, D, H, W, Cin, Cout, kd, kh, kw};
  cudaStream_t st = as_stream(stream);
  // the packed-weight layout differs per algorithm, so the caller must name one (b200seg_conv3d_algo)
  if (algo == B200SEG_ALGO_TC) return conv3d_fwd_tc(a, dtype, st);
  if (algo != B200SEG_ALGO_DIRECT) return B200SEG_EINVAL;
  {
    const int rc = conv3d_fwd_small(a, dtype, st);      // HBM-bound special cases (stem, classifier head)
    if (rc != B200SEG_EUNSUPPORTED) return rc;
  }
  return conv3d_fwd_direct(a, dtype, st);
}


// ---- bias gradient db[c] += sum_v dy[v][c] as its own column-sum pass, so that convolutions / Linears WITH a bias
// (every nn.Linear of SwinUNETR: qkv, proj, fc1, fc2 — 32 per step) can take the tcgen05 weight-gradient kernel, which has
// no bias path.  HBM-bound: dy is read once (16-byte loads), block partials in shared memory, fp32 atomics.
namespace {
constexpr int kBgThreads = 256, kBgChan = 512;      // channels per block column (grid.y)
template <typename T>
__global__ void __launch_bounds__(kBgThreads) bias_grad_kernel(const T* __restrict__ dy, int ld, int coff, float* __restrict__ db,
                                                               int64_t nvox, int C, int64_t vpb) {
  __shared__ float sm[kBgThreads * 8];
  const int c0 = blockIdx.y * kBgChan, cn = (C - c0 < kBgChan) ? C - c0 : kBgChan;
  const int cpv = cn / 8, vpp = kBgThreads / cpv, cchunk = threadIdx.x % cpv, vloc = threadIdx.x / cpv;
  const bool active = vloc < vpp;
  const int64_t v0 = (int64_t)blockIdx.x * vpb, v1 = (v0 + vpb < nvox) ? v0 + vpb : nvox;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (active) {
    const T* p = dy + coff + c0 + cchunk * 8;
    for (int64_t v = v0 + vloc; v < v1; v += vpp) {
      float a[8];
      ld8<T>(p + v * ld, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += a[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[threadIdx.x * 8 + i] = active ? acc[i] : 0.f;
  __syncthreads();
  for (int o = threadIdx.x; o < cn; o += kBgThreads) {
    float s = 0.f;
    for (int vl = 0; vl < vpp; ++vl) s += sm[(vl * cpv + o / 8) * 8 + (o & 7)];
    atomicAdd(&db[c0 + o], s);
  }
}

int launch_bias_grad(const WgradArgs& a, int dtype, cudaStream_t st) {
  const int64_t nvox = (int64_t)a.B * a.D * a.H * a.W;
  int64_t blocks = (B200SEG_NUM_SMS * 4) / ((a.Cout + kBgChan - 1) / kBgChan);
  if (blocks < 1) blocks = 1;
  int64_t vpb = (nvox + blocks - 1) / blocks;
  if (vpb < 64) vpb = 64;
  dim3 grid(ceil_div(nvox, vpb), (a.Cout + kBgChan - 1) / kBgChan);
  if (dtype == B200SEG_F16) bias_grad_kernel<__half><<<grid, kBgThreads, 0, st>>>((const __half*)a.dy, a.dy_ld, a.dy_coff, a.dbias, nvox, a.Cout, vpb);
  else bias_grad_kernel<float><<<grid, kBgThreads, 0, st>>>((const float*)a.dy, a.dy_ld, a.dy_coff, a.dbias, nvox, a.Cout, vpb);
  B200_CHECK_LAUNCH("bias_grad_kernel");
  return B200SEG_OK;
}
}  // namespace

static int wgrad_args(WgradArgs& a, const void* x, int x_ld, int x_coff, const double* x_stats, float eps, int act,
                      const void* dy, int dy_ld, int dy_coff, float* dw, float* dbias, int B, int D, int H, int W,
                      int Cin, int Cout, int kd, int kh, int kw, int dtype) {
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return B200SEG_EINVAL;
  if (kd <= 0 || kh <= 0 || kw <= 0 || !(kd & 1) || !(kh & 1) || !(kw & 1)) return B200SEG_EUNSUPPORTED;
  if (dtype != B200SEG_F16 && dtype != B200SEG_F32) return B200SEG_EINVAL;
  a = WgradArgs{x, x_ld, x_coff, x_stats, eps, act, dy, dy_ld, dy_coff, dw, dbias, B, D, H, W, Cin, Cout, kd, kh, kw, 0};
  return B200SEG_OK;
}

extern "C" size_t b200seg_conv3d_wgrad_workspace(int x_ld, int x_coff, int normalised, int dy_ld, int dy_coff,
                                                 int want_bias, int B, int D, int H, int W, int Cin, int Cout,
                                                 int kd, int kh, int kw, int dtype, int algo) {
  WgradArgs a;
  static const double dummy_stats = 0.0;
  static float dummy_bias = 0.f;
  // 16-byte aligned placeholders stand in for the tensors (only shapes/strides decide)
  if (wgrad_args(a, (const void*)16, x_ld, x_coff, normalised ? &dummy_stats : nullptr, 1e-4f, normalised ? 1 : 0,
                 (const void*)16, dy_ld, dy_coff, (float*)16, want_bias ? &dummy_bias : nullptr, B, D, H, W, Cin, Cout,
                 kd, kh, kw, dtype)) return 0;
  if (algo == B200SEG_ALGO_DIRECT) return 0;
  if (!conv3d_wgrad_tc_supported(a, dtype)) return 0;
  return conv3d_wgrad_tc_workspace(a);
}

extern "C" int b200seg_conv3d_wgrad(const void* x, int x_ld, int x_coff, const double* x_stats, float eps, int act,
                                    const void* dy, int dy_ld, int dy_coff, float* dw, float* dbias, int B, int D,
                                    int H, int W, int Cin, int Cout, int kd, int kh, int kw, int dtype, int algo,
                                    void* workspace, size_t ws_bytes, void* stream) {
  if (!x || !dy || !dw) return B200SEG_EINVAL;
  WgradArgs a;
  int rc = wgrad_args(a, x, x_ld, x_coff, x_stats, eps, act, dy, dy_ld, dy_coff, dw, dbias, B, D, H, W, Cin, Cout, kd, kh, kw, dtype);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (algo == B200SEG_ALGO_TC) return conv3d_wgrad_tc(a, dtype, workspace, ws_bytes, st);
  if (algo == B200SEG_ALGO_AUTO && conv3d_wgrad_tc_supported(a, dtype))
    return conv3d_wgrad_tc(a, dtype, workspace, ws_bytes, st);
  if (algo != B200SEG_ALGO_AUTO && algo != B200SEG_ALGO_DIRECT) return B200SEG_EINVAL;
  if (algo == B200SEG_ALGO_AUTO && a.dbias && a.Cout % 8 == 0) {
    // the tcgen05 kernel has no bias path: take the bias gradient in its own pass and let it do the weight gradient.
    // Ahead of the special kernels on purpose: for the 8- / 16-class 1x1x1 heads this pair is 2-4.8x faster than
    // wgrad_head_kernel (tools/head_wgrad_ab.py: 48->16 @128^3 134 vs 646 us); Cout = 4 and the Cin = 1 stems do not qualify
    WgradArgs nb = a;
    nb.dbias = nullptr;
    if (conv3d_wgrad_tc_supported(nb, dtype)) {
      rc = launch_bias_grad(a, dtype, st);
      if (rc) return rc;
      return conv3d_wgrad_tc(nb, dtype, workspace, ws_bytes, st);
    }
  }
  if (algo == B200SEG_ALGO_AUTO) {
    rc = conv3d_wgrad_small(a, dtype, st);
    if (rc != B200SEG_EUNSUPPORTED) return rc;
  }
  return conv3d_wgrad_direct(a, dtype, st);
}
