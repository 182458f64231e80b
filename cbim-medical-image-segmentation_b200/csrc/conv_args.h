// conv_args.h — argument blocks shared by the direct and tcgen05 conv implementations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct ConvArgs {
  const void* x; int x_ld, x_coff;
  const double* x_stats; float eps; int act;
  const void* w; const float* bias;
  const void* res; int r_ld, r_coff;
  void* y; int y_ld, y_coff; double* y_stats;
  const void* gx; int gx_ld, gx_coff; const double* g_stats; float g_eps; int g_act;
  int B, D, H, W, Cin, Cout, kd, kh, kw;
};

struct WgradArgs {
  const void* x; int x_ld, x_coff; const double* x_stats; float eps; int act;
  const void* dy; int dy_ld, dy_coff;
  float* dw; float* dbias;
  int B, D, H, W, Cin, Cout, kd, kh, kw;
  int vox_per_block;
};

int conv3d_fwd_direct(const ConvArgs& a, int dtype, cudaStream_t st);
int conv3d_wgrad_direct(const WgradArgs& a, int dtype, cudaStream_t st);
// tcgen05 paths (conv_tc.cu / wgrad_tc.cu); return B200SEG_EUNSUPPORTED when the shape does not qualify
int conv3d_fwd_tc(const ConvArgs& a, int dtype, cudaStream_t st);
int conv3d_wgrad_tc(const WgradArgs& a, int dtype, cudaStream_t st);
bool conv3d_fwd_tc_supported(const ConvArgs& a, int dtype);
bool conv3d_wgrad_tc_supported(const WgradArgs& a, int dtype);
