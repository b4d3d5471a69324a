// Sigmoid + straight-through binarizer of the ImageCompressor_new encoder:
//   models/analysis_17_new.py:35-37  x = sigmoid(conv3(...)); code = Binarizer(x) = (x > 0.5).float()
//   models/binarizer.py:19-27        backward: identity (straight-through estimator)
// Forward is one pass: fp32 NCHW conv3 output -> sigmoid (fp32 NCHW), code (fp32 NCHW) and, for the tensor-core
// synthesis transform, the code as NHWC bf16 (0 / 1 are exact in bf16).  Backward: dL/dy = (g_code + g_sig) s (1 - s).
#include <cuda_bf16.h>

#include "b2c_common.cuh"

namespace b2c {

// one block = 32 pixels x all channels of one image (like quant_entropy_infer): coalesced NCHW reads along pixels,
// NHWC bf16 rows written through a shared-memory transpose
constexpr int SB_PIX = 32;

__global__ void __launch_bounds__(256) sigmoid_binarize_fwd_kernel(const float* __restrict__ y, float* __restrict__ sig,
                                                                   float* __restrict__ code,
                                                                   __nv_bfloat16* __restrict__ code_nhwc, int C, int HW,
                                                                   int chunks) {
  extern __shared__ __nv_bfloat16 tile[];   // [SB_PIX][C + 2]
  const int n = blockIdx.x / chunks, p0 = (blockIdx.x % chunks) * SB_PIX;
  const int px = threadIdx.x & 31, cg = threadIdx.x >> 5;    // 8 channel groups
  const int pix = p0 + px;
  const int TS = C + 2;
  for (int c = cg; c < C; c += 8) {
    float b = 0.f;
    if (pix < HW) {
      const size_t a = ((size_t)n * C + c) * HW + pix;
      const float s = 1.f / (1.f + expf(-__ldg(y + a)));       // torch.sigmoid in fp32
      b = s > 0.5f ? 1.f : 0.f;                                // models/binarizer.py:20
      sig[a] = s;
      code[a] = b;
    }
    tile[px * TS + c] = __float2bfloat16_rn(b);     // padded row stride: lanes (pixels) hit distinct banks
  }
  if (!code_nhwc) return;
  __syncthreads();
  const int rows = min(SB_PIX, HW - p0);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(tile);
  uint32_t* dst = reinterpret_cast<uint32_t*>(code_nhwc + ((size_t)n * HW + p0) * C);
  const int wpr = C / 2;
  for (int i = threadIdx.x; i < rows * wpr; i += blockDim.x) {
    const int row = i / wpr, col = i - row * wpr;
    dst[i] = src[row * (TS / 2) + col];
  }
}

// Second version for C % 64 == 0, HW % 4 == 0 (same layout as quant_entropy_infer_lut2_kernel, csrc/b2c_entropy.cu):
// CTA = 128 pixels x all channels, thread = 4 consecutive pixels of TWO adjacent channels (LDG.128 / STG.128), the
// channel pair packed as one bf16x2 word into a tile XOR-swizzled in 16-byte chunks (chunk ^ ((row >> 2) & 7)), copied
// out with LDS.128 / STG.128.  Same sigmoid expression, so sig / code are bit-identical to the first version.
constexpr int SB2_PIX = 128;

__device__ __forceinline__ float sb_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void __launch_bounds__(256) sigmoid_binarize_fwd2_kernel(const float* __restrict__ y, float* __restrict__ sig,
                                                                    float* __restrict__ code,
                                                                    __nv_bfloat16* __restrict__ code_nhwc, int C, int HW,
                                                                    int chunks) {
  extern __shared__ __align__(16) uint32_t tile2[];          // [SB2_PIX][C / 2]
  const unsigned n = blockIdx.x / chunks, pix0 = (blockIdx.x - n * chunks) * SB2_PIX;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned q = lane & 7, k = lane >> 3, half = warp >> 2;
  const unsigned lp = (warp & 3) * 32 + q * 4;
  const unsigned wpr = C / 2;
  const bool in_image = pix0 + lp < (unsigned)HW;
  size_t a0 = ((size_t)n * C + (half * 4 + k) * 2) * HW + pix0 + lp;
  const size_t cstep = (size_t)16 * HW;
  uint32_t* tp = tile2 + lp * wpr + k;
#pragma unroll 2
  for (unsigned pass = 0; pass < (unsigned)C / 16; ++pass, a0 += cstep) {
    float4 ca = make_float4(0.f, 0.f, 0.f, 0.f), cb = ca;
    if (in_image) {
      const float4 va = __ldg(reinterpret_cast<const float4*>(y + a0));
      const float4 vb = __ldg(reinterpret_cast<const float4*>(y + a0 + HW));
      const float4 sa = make_float4(sb_sigmoid(va.x), sb_sigmoid(va.y), sb_sigmoid(va.z), sb_sigmoid(va.w));
      const float4 sb = make_float4(sb_sigmoid(vb.x), sb_sigmoid(vb.y), sb_sigmoid(vb.z), sb_sigmoid(vb.w));
      ca = make_float4(sa.x > 0.5f ? 1.f : 0.f, sa.y > 0.5f ? 1.f : 0.f, sa.z > 0.5f ? 1.f : 0.f, sa.w > 0.5f ? 1.f : 0.f);
      cb = make_float4(sb.x > 0.5f ? 1.f : 0.f, sb.y > 0.5f ? 1.f : 0.f, sb.z > 0.5f ? 1.f : 0.f, sb.w > 0.5f ? 1.f : 0.f);
      *reinterpret_cast<float4*>(sig + a0) = sa;
      *reinterpret_cast<float4*>(sig + a0 + HW) = sb;
      *reinterpret_cast<float4*>(code + a0) = ca;
      *reinterpret_cast<float4*>(code + a0 + HW) = cb;
    }
    uint32_t* t0 = tp + (((pass * 2 + half) ^ q) << 2);
    const __nv_bfloat162 w0 = __floats2bfloat162_rn(ca.x, cb.x), w1 = __floats2bfloat162_rn(ca.y, cb.y);
    const __nv_bfloat162 w2 = __floats2bfloat162_rn(ca.z, cb.z), w3 = __floats2bfloat162_rn(ca.w, cb.w);
    t0[0] = *reinterpret_cast<const uint32_t*>(&w0);
    t0[wpr] = *reinterpret_cast<const uint32_t*>(&w1);
    t0[2 * wpr] = *reinterpret_cast<const uint32_t*>(&w2);
    t0[3 * wpr] = *reinterpret_cast<const uint32_t*>(&w3);
  }
  if (!code_nhwc) return;
  __syncthreads();
  const unsigned npix = min((unsigned)SB2_PIX, (unsigned)HW - pix0), cpr = C / 8;
  const uint4* src = reinterpret_cast<const uint4*>(tile2);
  uint4* dst = reinterpret_cast<uint4*>(code_nhwc + ((size_t)n * HW + pix0) * C);
  for (unsigned i = threadIdx.x; i < npix * cpr; i += 256) {
    const unsigned r = i / cpr, ch = i - r * cpr;
    dst[i] = src[r * cpr + (ch ^ ((r >> 2) & 7))];
  }
}

__global__ void __launch_bounds__(256) sigmoid_binarize_bwd_kernel(const float* __restrict__ sig,
                                                                   const float* __restrict__ g_code,
                                                                   const float* __restrict__ g_sig,
                                                                   float* __restrict__ g_y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float s = sig[i];
    const float g = (g_code ? g_code[i] : 0.f) + (g_sig ? g_sig[i] : 0.f);
    g_y[i] = g * s * (1.f - s);
  }
}

// 8-bit image -> fp32 in [0, 1]: torchvision's ToTensor (img.to(float32).div(255)), IEEE division so the result is
// bit-identical to the host-side conversion the reference's data loader performs (datasets.py:24-34)
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uchar4* __restrict__ in, float4* __restrict__ out, size_t n4) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const uchar4 v = in[i];
    out[i] = make_float4(__fdiv_rn((float)v.x, 255.f), __fdiv_rn((float)v.y, 255.f), __fdiv_rn((float)v.z, 255.f),
                         __fdiv_rn((float)v.w, 255.f));
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_u8_to_f32(const void* in_u8, float* out, size_t count, void* stream) {
  B2C_REQUIRE(in_u8 && out && count > 0 && count % 4 == 0, "u8_to_f32: bad arguments (count must be a multiple of 4)");
  B2C_REQUIRE(((uintptr_t)in_u8 & 3) == 0 && ((uintptr_t)out & 15) == 0, "u8_to_f32: misaligned buffers");
  const size_t n4 = count / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  u8_to_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uchar4*)in_u8, (float4*)out, n4);
  count_launch();
  return check_launch("u8_to_f32");
}

extern "C" int b2c_sigmoid_binarize_fwd(const float* y, float* sig, float* code, void* code_nhwc_bf16, int B, int C, int HW,
                                        void* stream) {
  B2C_REQUIRE(y && sig && code && B > 0 && C > 0 && C % 2 == 0 && HW > 0, "sigmoid_binarize_fwd: bad arguments");
  if (C % 64 == 0 && HW % 4 == 0 && C <= 256 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)sig & 15) == 0 &&
      ((uintptr_t)code & 15) == 0 && ((uintptr_t)code_nhwc_bf16 & 15) == 0) {
    const int chunks2 = (HW + SB2_PIX - 1) / SB2_PIX;
    const size_t smem2 = (size_t)SB2_PIX * C * sizeof(__nv_bfloat16);
    if (smem2 > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(sigmoid_binarize_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           SB2_PIX * 256 * 2);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
    }
    sigmoid_binarize_fwd2_kernel<<<B * chunks2, 256, smem2, (cudaStream_t)stream>>>(y, sig, code,
                                                                                   (__nv_bfloat16*)code_nhwc_bf16, C, HW, chunks2);
    count_launch();
    return check_launch("sigmoid_binarize_fwd");
  }
  const int chunks = (HW + SB_PIX - 1) / SB_PIX;
  const size_t smem = (size_t)SB_PIX * (C + 2) * sizeof(__nv_bfloat16);
  B2C_REQUIRE(smem <= 48 * 1024, "sigmoid_binarize_fwd: too many channels");
  sigmoid_binarize_fwd_kernel<<<B * chunks, 256, smem, (cudaStream_t)stream>>>(y, sig, code, (__nv_bfloat16*)code_nhwc_bf16, C,
                                                                              HW, chunks);
  count_launch();
  return check_launch("sigmoid_binarize_fwd");
}

extern "C" int b2c_sigmoid_binarize_bwd(const float* sig, const float* g_code, const float* g_sig, float* g_y, size_t count,
                                        void* stream) {
  B2C_REQUIRE(sig && g_y && count > 0, "sigmoid_binarize_bwd: bad arguments");
  size_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sigmoid_binarize_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sig, g_code, g_sig, g_y, count);
  count_launch();
  return check_launch("sigmoid_binarize_bwd");
}
