// placeholder until the tcgen05 kernels land (next commit)
#include "common.cuh"
#include "conv_args.h"
bool conv3d_fwd_tc_supported(const ConvArgs&, int) { return false; }
bool conv3d_wgrad_tc_supported(const WgradArgs&, int) { return false; }
int conv3d_fwd_tc(const ConvArgs&, int, cudaStream_t) { return B200SEG_EUNSUPPORTED; }
int conv3d_wgrad_tc(const WgradArgs&, int, cudaStream_t) { return B200SEG_EUNSUPPORTED; }
