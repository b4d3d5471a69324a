// Per-element pieces of the factorized entropy model (models/bitEstimator.py:20-42, model.py:71-73) shared by the
// entropy kernels (b2c_entropy.cu) and the convolution kernel whose epilogue quantizes and counts bits itself
// (b2c_conv_tcp.cu, the last analysis layer at inference).
#pragma once
#include "b2c_common.cuh"

namespace b2c {

struct BitPtrs {
  const float* p[B2C_NUM_BITPARAMS];
};

struct ChanConst {
  float sp[4];  // softplus(h_k)
  float b[4];
  float ta[3];  // tanh(a_k)
};

__device__ __forceinline__ float softplus_t(float h) {  // torch F.softplus, beta=1, threshold=20
  return h > 20.f ? h : log1pf(expf(h));
}
__device__ __forceinline__ ChanConst load_chan(const BitPtrs& bp, int c) {
  ChanConst k;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    k.sp[i] = softplus_t(__ldg(bp.p[3 * i] + c));
    k.b[i] = __ldg(bp.p[3 * i + 1] + c);
    k.ta[i] = tanhf(__ldg(bp.p[3 * i + 2] + c));
  }
  k.sp[3] = softplus_t(__ldg(bp.p[9] + c));
  k.b[3] = __ldg(bp.p[10] + c);
  return k;
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float cdf_fast(const float* k, float v) {   // k: sp1 b1 ta1 sp2 b2 ta2 sp3 b3 ta3 sp4 b4
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float t = fmaf(v, k[3 * i], k[3 * i + 1]);
    v = fmaf(tanh_fast(t), k[3 * i + 2], t);
  }
  const float t4 = fmaf(v, k[9], k[10]);
  return __fdividef(1.f, 1.f + __expf(-t4));
}

// integers up to 256 are exact in bf16 (8 significand bits); the NHWC bf16 latent the tensor-core synthesis transform
// reads is only valid below that.  The inference kernels raise *range_flag (sticky, never cleared here) otherwise.
constexpr float BF16_EXACT_INT = 256.f;

__device__ __forceinline__ float elem_bits_fast(const float* k, float zv) {
  const float p = cdf_fast(k, zv + 0.5f) - cdf_fast(k, zv - 0.5f);
  return fminf(fmaxf(-__log2f(p + 1e-10f), 0.f), 50.f);
}

// |z| beyond the table: the direct formula, kept out of line so the hot loop stays small (registers -> occupancy)
static __device__ __noinline__ float bits_outside_table(const BitPtrs& bp, int c, float zv) {
  const ChanConst kc = load_chan(bp, c);
  const float k[12] = {kc.sp[0], kc.b[0], kc.ta[0], kc.sp[1], kc.b[1], kc.ta[1], kc.sp[2], kc.b[2], kc.ta[2],
                       kc.sp[3], kc.b[3], 0.f};
  return elem_bits_fast(k, zv);
}


}  // namespace b2c
