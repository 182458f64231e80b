// conv_args.h — argument blocks shared by the direct and tcgen05 conv implementations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

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
// HBM-bound special cases (Cin=1 stem, 1x1x1 head); EUNSUPPORTED when the shape is not one of them
int conv3d_wgrad_small(const WgradArgs& a, int dtype, cudaStream_t st);
int conv3d_fwd_small(const ConvArgs& a, int dtype, cudaStream_t st);
// tcgen05 paths (conv_tc.cu / wgrad_tc.cu); return B200SEG_EUNSUPPORTED when the shape does not qualify
int conv3d_fwd_tc(const ConvArgs& a, int dtype, cudaStream_t st);
int conv3d_wgrad_tc(const WgradArgs& a, int dtype, void* workspace, size_t ws_bytes, cudaStream_t st);
size_t conv3d_wgrad_tc_workspace(const WgradArgs& a);
bool conv3d_fwd_tc_supported(const ConvArgs& a, int dtype);
bool conv3d_wgrad_tc_supported(const WgradArgs& a, int dtype);

// ---- tcgen05 tiling choices shared by the packer and the kernels
// N tile (output channels per CTA tile): whole Cout when <=256, else the largest even split.
__host__ __device__ static inline int tc_pick_nt(int Cout) {
  if (Cout % 16) return 0;
  if (Cout <= 256) return Cout;
  for (int t = (Cout + 255) / 256; t <= 16; ++t)
    if (Cout % t == 0 && (Cout / t) % 16 == 0 && Cout / t <= 256) return Cout / t;
  return 0;
}
// K chunk (input channels per staged halo tile): largest multiple of 16 dividing Cin, <= 64.
__host__ __device__ static inline int tc_pick_kc(int Cin) {
  if (Cin % 16) return 0;
  for (int kc = 64; kc >= 16; kc -= 16)
    if (Cin % kc == 0) return kc;
  return 0;
}
bool conv3d_tc_shape_ok(int Cin, int Cout, int kd, int kh, int kw, int dtype);
