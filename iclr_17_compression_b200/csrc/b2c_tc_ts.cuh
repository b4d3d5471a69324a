// Pieces shared by the kernels that keep tcgen05.mma A operands in TMEM (b2c_conv1_ts.cu, b2c_conv1_rs.cu): TMEM
// stores, the TS form of the pair MMA, the CONV1 im2col builder into registers and the per-warp output tensor map.
#pragma once
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"
#include "b2c_tc_host.cuh"
#include "b2c_tcp_util.cuh"

namespace b2c {
namespace tc {

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_stN(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st32(taddr, r); }
__device__ __forceinline__ void tmem_stN(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T over the CTA pair (M = 256: each CTA supplies the 128 rows held in its own TMEM
// at column a_tmem, and half of B's rows); leader CTA only
__device__ __forceinline__ void umma_bf16_pair_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// CONV1 im2col: the CHP values k = KB*64 + PART*CHP .. (k = c*81 + ky*9 + kx, zero for k >= 243) of the pixel whose
// patch origin is `pp` (16-byte aligned), as CHP/2 packed bf16x2 hi words (pair i = k0+2i, k0+2i+1) followed by the
// CHP/2 lo words.  Same fetch pattern as tc::conv1_build_row (two conflict-free LDS.128 + one LDS.32 per (c, ky) patch
// row, only the pieces this k range uses); every index is a compile-time constant.
template <int KB, int PART, int CHP>
__device__ __forceinline__ void conv1_build_regs(const float* __restrict__ pp, uint32_t (&w)[CHP]) {
  constexpr int K0 = KB * 64 + PART * CHP;
  constexpr int G0 = K0 / 9;
  constexpr int G1 = (K0 + CHP - 1) / 9;
  constexpr int NG = G1 - G0 + 1;
  float vals[NG * 12];
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    const int g = G0 + gi;
    if (g < 27) {
      const float* rowf = pp + (g / 9) * P1_PLANE + (g % 9) * P1_W;
      const float4* row = reinterpret_cast<const float4*>(rowf);
      const int klo = (K0 > 9 * g ? K0 : 9 * g) - 9 * g;
      const int kend = K0 + CHP - 1 < 242 ? K0 + CHP - 1 : 242;
      const int khi = (kend < 9 * g + 8 ? kend : 9 * g + 8) - 9 * g;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      float c8 = 0.f;
      if (klo <= 3) a = row[0];
      if (klo <= 7 && khi >= 4) b = row[1];
      if (khi >= 8) c8 = rowf[8];
      vals[gi * 12 + 0] = a.x; vals[gi * 12 + 1] = a.y; vals[gi * 12 + 2] = a.z; vals[gi * 12 + 3] = a.w;
      vals[gi * 12 + 4] = b.x; vals[gi * 12 + 5] = b.y; vals[gi * 12 + 6] = b.z; vals[gi * 12 + 7] = b.w;
      vals[gi * 12 + 8] = c8; vals[gi * 12 + 9] = 0.f; vals[gi * 12 + 10] = 0.f; vals[gi * 12 + 11] = 0.f;
    } else {
#pragma unroll
      for (int t = 0; t < 12; ++t) vals[gi * 12 + t] = 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < CHP / 2; ++i) {
    const int k = K0 + 2 * i;
    const float v0 = (k < 243) ? vals[(k / 9 - G0) * 12 + (k % 9)] : 0.f;
    const float v1 = (k + 1 < 243) ? vals[((k + 1) / 9 - G0) * 12 + ((k + 1) % 9)] : 0.f;
    w[i] = cvt_bf16x2(v0, v1);
    w[CHP / 2 + i] = lo_of(v0, v1, w[i]);
  }
}

// NHWC bf16 output [B][Hg][Wg][128] for the per-warp stores: box = chp channels x 16 x 2 pixels, SWIZZLE_64B / _32B
static inline int make_warp_out_map(CUtensorMap* m, const void* base, int B, int Hg, int Wg, int chp) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)128, (cuuint64_t)Wg, (cuuint64_t)Hg, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)128 * 2, (cuuint64_t)Wg * 128 * 2, (cuuint64_t)Hg * Wg * 128 * 2};
  cuuint32_t box[4] = {(cuuint32_t)chp, (cuuint32_t)TW, 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, chp == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(warp output) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

// Returns 1 when the layer is not the one this kernel covers (the caller then tries the shared-memory form),
// otherwise the launch status.  B2C_TC_PERSIST_CONV1: 2 (default) = this kernel, 1 = shared-memory operands
// (b2c_conv1_tcp.cu), 0 = per-tile kernel.

}  // namespace tc
}  // namespace b2c
