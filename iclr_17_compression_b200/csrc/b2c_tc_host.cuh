// Host-side helpers shared by the tensor-core kernels: cuTensorMapEncodeTiled lookup and the tensor maps of the
// NHWC bf16 activations, output lattices, weight panels and the fp32 image.
#pragma once
#include <mutex>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"

namespace b2c {
namespace tc {

constexpr int TH = 8, TW = 16;            // output tile (pixels)
constexpr int TILE_M = TH * TW;           // 128
constexpr int KBLK = 64;                  // bf16 per k-block row (128 B)
constexpr int SLOT_BYTES = TILE_M * KBLK * 2;  // 16 KB operand slot

// Programmatic dependent launch between consecutive layers of the inference pipeline (B2C_PDL=0 disables it)
static inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("B2C_PDL");
    return !e || atoi(e) != 0;
  }();
  return on;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  });
  return fn;
}

// NHWC bf16 activation [B][H][W][128]: box = 64 channels x TW x TH pixels with traversal stride S
static inline int make_act_map(CUtensorMap* m, const void* base, int B, int H, int W, int S, int C) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)KBLK, (cuuint32_t)(TW * S), (cuuint32_t)(TH * S), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)S, (cuuint32_t)S, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

// NHWC bf16 output lattice: pixel (oy*ps + ry, ox*ps + rx) of [B][Hfull][Wfull][128] seen as a dense
// [B][Hg][Wg][128] tensor with strides; box = 64 channels x TW x TH pixels (TMA clips at the image border)
static inline int make_out_map(CUtensorMap* m, const void* base, int B, int Hg, int Wg, int Hfull, int Wfull, int ps, int ry,
                        int rx, int C, int box_rows = TH, int box_ch = KBLK) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  const char* origin = reinterpret_cast<const char*>(base) + ((size_t)ry * Wfull + rx) * C * 2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wg, (cuuint64_t)Hg, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ps * C * 2, (cuuint64_t)ps * Wfull * C * 2,
                           (cuuint64_t)Hfull * Wfull * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_ch, (cuuint32_t)TW, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(origin), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_ch == KBLK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(output) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

// bf16 matrix [rows][cols] (cols contiguous): box = 64 cols x box_rows
static inline int make_mat_map(CUtensorMap* m, const void* base, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(matrix) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

// fp32 NCHW image [B][3][H][W]: box = 72 x 37 x 3 (the im2col support of one 8x16 CONV1 output tile)
static inline int make_image_map(CUtensorMap* m, const void* base, int B, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  cuuint32_t box[4] = {(cuuint32_t)P1_W, (cuuint32_t)P1_H, 3, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(image) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

}  // namespace tc
}  // namespace b2c
