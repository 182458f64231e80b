// tmap.h — host-side construction of TMA tensor maps (CUtensorMap) for channels-last activations, without linking
// libcuda: cuTensorMapEncodeTiled is fetched through the runtime's driver-entry-point query.
//
// An activation slice x[b][d][h][w][coff : coff + C] (fp16, row pitch ld) is described as the 5-D tensor
//   dim0 = 8 channels (16 B, contiguous)      dim1 = w (stride ld*2 B)      dim2 = h (stride W*ld*2)
//   dim3 = channel plane c/8 (stride 16 B)    dim4 = b*D + d (stride H*W*ld*2)
// so that ONE cp.async.bulk.tensor box {8, bw, bh, planes, 1} lands in shared memory as [plane][bh x bw voxels][8 ch] —
// exactly the no-swizzle UMMA operand image conv_tc.cu / wgrad_tc.cu consume (K-major for the forward GEMM, MN-major
// for the weight gradient) — with out-of-volume voxels (conv padding, ragged tiles) zero-filled by the TMA unit.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef CUresult (*b200seg_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                            const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                            CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline b200seg_encode_tiled_fn b200seg_encode_tiled() {
  static b200seg_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<b200seg_encode_tiled_fn>(p);
  }
  return fn;
}

// returns false when the map cannot be built (driver entry point missing / shape rejected): callers use their
// cp.async staging path then
static inline bool b200seg_make_act_tmap(CUtensorMap* m, const void* base_fp16, int ld, int coff, int C, int BD, int H, int W,
                                         int box_w, int box_h, int box_planes) {
  b200seg_encode_tiled_fn enc = b200seg_encode_tiled();
  if (!enc || (C % 8) || (ld % 8) || (coff % 8)) return false;
  const char* base = reinterpret_cast<const char*>(base_fp16) + (size_t)coff * 2;
  if (reinterpret_cast<uintptr_t>(base) & 15) return false;
  if (box_w > 256 || box_h > 256 || box_planes > 256) return false;
  cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)BD};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, 16, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[5] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_planes, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Row image (round 2): the same slice as the 4-D tensor  dim0 = channel (contiguous)  dim1 = w  dim2 = h  dim3 = b*D + d,
// box {KC, bw, bh, 1} with KC*2 = 64 or 128 bytes and the matching hardware swizzle: shared memory receives one 64- /
// 128-byte ROW per halo voxel — the canonical K-major (forward / data gradient) resp. MN-major (weight gradient)
// SWIZZLE_64B / SWIZZLE_128B UMMA operand — and the TMA unit moves 64-128 contiguous bytes per request instead of 16.
static inline bool b200seg_make_row_tmap(CUtensorMap* m, const void* base_fp16, int ld, int coff, int C, int KC, int BD, int H, int W,
                                         int box_w, int box_h) {
  b200seg_encode_tiled_fn enc = b200seg_encode_tiled();
  // C need not be a multiple of KC: channels of a box beyond the tensor's extent are zero-filled like any out-of-range element
  if (!enc || (KC != 32 && KC != 64) || (C % 8) || (ld % 8) || (coff % 8)) return false;
  const char* base = reinterpret_cast<const char*>(base_fp16) + (size_t)coff * 2;
  if (reinterpret_cast<uintptr_t>(base) & 15) return false;
  if (box_w > 256 || box_h > 256) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BD};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
