// Shared-memory operand builders used by the tensor-core kernels: swizzled K-major row chunks and the
// 9x9 / stride-4 im2col rows (CONV1 forward, deconv3 dgrad, and the 9x9 weight gradients).
#pragma once
#include "b2c_tc.cuh"

namespace b2c {
namespace tc {

constexpr int IM_SLOT_BYTES = 128 * 64 * 2;  // one 128-row x 64-element bf16 tile

// CONV1 input patch of one 8x16 output tile: 3 channels x 37 rows x 69 (->72) columns of fp32
constexpr int P1_W = 72, P1_H = 37;
constexpr int P1_PLANE = P1_W * P1_H;
constexpr int P1_BYTES = 3 * P1_PLANE * 4;          // 31968
constexpr int P1_SLOT = 32768;

// fp32 pair -> packed bf16x2 (one F2FP instruction; the scalar F2F conversions run at quarter rate and were
// the bottleneck of every operand builder / epilogue), low half = a
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool SPLIT3>
__device__ __forceinline__ void store_row_chunk(uint8_t* slot_hi, uint8_t* slot_lo, int r, int cidx,
                                                const float (&v)[8]) {
  // K-major SWIZZLE_128B: row r at r*128, 16-byte chunk index XORed with (r & 7)
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cidx ^ (r & 7)) << 4);
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = cvt_bf16x2(v[2 * i], v[2 * i + 1]);
    if (SPLIT3) {   // lo = bf16(x - float(hi)): the residual is exact in fp32
      const float r0 = v[2 * i] - __uint_as_float(h[i] << 16);
      const float r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
      l[i] = cvt_bf16x2(r0, r1);
    }
  }
  *reinterpret_cast<uint4*>(slot_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (SPLIT3) *reinterpret_cast<uint4*>(slot_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// CONV1: the 32 im2col values k = KB*64 + HALF*32 .. +31 (k = c*81 + ky*9 + kx, zero for k >= 243) of the
// pixel whose patch origin is `pp` (16-byte aligned), written as four swizzled 16-byte chunks (hi, lo) of the
// pixel's 128-byte K-major row.  Each (c, ky) patch row is fetched with up to two conflict-free LDS.128 and one
// LDS.32 (9 floats); every index below is a compile-time constant.
template <int KB, int HALF, bool SPLIT3>
__device__ __forceinline__ void conv1_build_row(const float* __restrict__ pp, uint8_t* slot_hi, uint8_t* slot_lo,
                                                int r) {
  constexpr int K0 = KB * 64 + HALF * 32;
  constexpr int G0 = K0 / 9;
  constexpr int G1 = (K0 + 31) / 9;
  constexpr int NG = G1 - G0 + 1;
  float vals[NG * 12];
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    const int g = G0 + gi;
    if (g < 27) {
      // only the 16-byte pieces (kx 0-3, 4-7) and the scalar kx = 8 that this 32-wide k range uses are fetched:
      // shared-memory wavefronts, not instructions, bound the builders
      const float* rowf = pp + (g / 9) * P1_PLANE + (g % 9) * P1_W;
      const float4* row = reinterpret_cast<const float4*>(rowf);
      const int klo = (K0 > 9 * g ? K0 : 9 * g) - 9 * g;
      const int kend = K0 + 31 < 242 ? K0 + 31 : 242;
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
  for (int ci = 0; ci < 4; ++ci) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = K0 + ci * 8 + j;
      v[j] = (k < 243) ? vals[(k / 9 - G0) * 12 + (k % 9)] : 0.f;
    }
    store_row_chunk<SPLIT3>(slot_hi, slot_lo, r, HALF * 4 + ci, v);
  }
}


}  // namespace tc
}  // namespace b2c
