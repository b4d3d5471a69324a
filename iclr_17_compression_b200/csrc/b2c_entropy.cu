// Quantizer + factorized entropy model + distortion (all fp32, one pass each).
//   quantizer   model.py:53-56        z = y + u (training) | rint(y) half-to-even (eval)
//   Bitparm     bitEstimator.py:20-25 t = v*softplus(h)+b ; v' = t + tanh(t)*tanh(a) ; last: sigmoid(t)
//   bits        model.py:71-73        p = CDF(z+.5) - CDF(z-.5) ; bits = sum clamp(-log(p+1e-10)/ln2, 0, 50)
//   distortion  model.py:59,61        clipped = clamp(recon,0,1) ; mse on the UNclamped recon
// The reference's formula is reproduced as written (difference of two fp32 sigmoids, +1e-10 floor),
// not a numerically nicer one.  One CTA per (image, channel) plane so the per-channel constants
// softplus(h_k), tanh(a_k), b_k are computed once per CTA; sums are fixed-order (deterministic).
#include <cuda_bf16.h>

#include <mutex>

#include "b2c_common.cuh"
#include "b2c_entropy_formula.cuh"

namespace b2c {

struct BitGradPtrs {
  float* p[B2C_NUM_BITPARAMS];
};

__device__ __forceinline__ float sigmoid_t(float t) { return 1.f / (1.f + expf(-t)); }

__device__ __forceinline__ float cdf(const ChanConst& k, float v) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float t = fmaf(v, k.sp[i], k.b[i]);
    v = fmaf(tanhf(t), k.ta[i], t);
  }
  return sigmoid_t(fmaf(v, k.sp[3], k.b[3]));
}

__device__ __forceinline__ float elem_bits(const ChanConst& k, float z) {
  const float p = cdf(k, z + 0.5f) - cdf(k, z - 0.5f);
  const float bits = -logf(p + 1e-10f) / 0.6931471805599453f;
  return fminf(fmaxf(bits, 0.f), 50.f);
}

__global__ void __launch_bounds__(256) quant_entropy_fwd_kernel(const float* __restrict__ y,
                                                                const float* __restrict__ noise,
                                                                BitPtrs bp, float* __restrict__ z,
                                                                float* __restrict__ plane_bits, int C, int HW) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int plane = blockIdx.x;  // n*C + c
  const int c = plane % C;
  const ChanConst k = load_chan(bp, c);
  const size_t base = (size_t)plane * HW;
  float sum = 0.f;
  if ((HW & 3) == 0) {
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    const float4* n4 = noise ? reinterpret_cast<const float4*>(noise + base) : nullptr;
    float4* z4 = reinterpret_cast<float4*>(z + base);
    for (int i = threadIdx.x; i < (HW >> 2); i += blockDim.x) {
      float4 v = __ldg(y4 + i);
      if (n4) {
        const float4 u = __ldg(n4 + i);
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      } else {
        v.x = rintf(v.x); v.y = rintf(v.y); v.z = rintf(v.z); v.w = rintf(v.w);
      }
      z4[i] = v;
      sum += elem_bits(k, v.x);
      sum += elem_bits(k, v.y);
      sum += elem_bits(k, v.z);
      sum += elem_bits(k, v.w);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float v = __ldg(y + base + i);
      v = noise ? v + __ldg(noise + base + i) : rintf(v);
      z[base + i] = v;
      sum += elem_bits(k, v);
    }
  }
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) plane_bits[plane] = sum;
}

// image_bits[n] = sum_c plane_bits[n*C + c], fixed order
__global__ void __launch_bounds__(256) plane_reduce_kernel(const float* __restrict__ plane_vals,
                                                           float* __restrict__ out, int per_image) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < per_image; i += blockDim.x) s += plane_vals[(size_t)n * per_image + i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[n] = s;
}

__global__ void partial_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int per_image,
                                      int width);

// ---------------------------------------------------------------------------- inference variant
// Same formula (model.py:53-56,71-73) with the hardware approximations tanh.approx / ex2.approx / lg2.approx for the
// per-element chain (relative error ~2^-11 per tanh changes the total bit count by ~1e-6 relative, far inside the
// 1e-3 bpp bar; the per-channel softplus / tanh constants stay exact).  One CTA = 32 pixels x all channels of one
// image, so the rounded latent is also emitted as the NHWC bf16 tensor the tensor-core synthesis transform reads.
// (tanh_fast / cdf_fast / elem_bits_fast, BF16_EXACT_INT: b2c_entropy_formula.cuh)
constexpr int QI_PIX = 32;

__global__ void __launch_bounds__(256) quant_entropy_infer_kernel(const float* __restrict__ y, BitPtrs bp,
                                                                  float* __restrict__ z,
                                                                  __nv_bfloat16* __restrict__ z_nhwc,
                                                                  float* __restrict__ partial, int C, int HW,
                                                                  int chunks, int* __restrict__ range_flag) {
  extern __shared__ float qi_smem[];
  float* kc = qi_smem;                                   // [C][12]
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(qi_smem + (size_t)C * 12);   // [QI_PIX][C]
  __shared__ float red[32];
  const int n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
  const int pix0 = chunk * QI_PIX;
  // programmatic dependent launch: the per-channel constants are set up while the producer of y drains
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const ChanConst k = load_chan(bp, c);
    float* d = kc + c * 12;
    d[0] = k.sp[0]; d[1] = k.b[0]; d[2] = k.ta[0];
    d[3] = k.sp[1]; d[4] = k.b[1]; d[5] = k.ta[1];
    d[6] = k.sp[2]; d[7] = k.b[2]; d[8] = k.ta[2];
    d[9] = k.sp[3]; d[10] = k.b[3]; d[11] = 0.f;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int pix = pix0 + lane;
  const int TS = C + 2;       // padded tile row (bf16 elements); with stride C every lane hit the same bank (ncu:
                              // 36.6 M store bank conflicts, LSU wavefronts 91 % of peak - that bound the kernel)
  float sum = 0.f;
  for (int c = warp; c < C; c += nwarps) {
    float zv = 0.f;
    if (pix < HW) {
      const size_t a = ((size_t)n * C + c) * HW + pix;
      zv = rintf(__ldg(y + a));
      z[a] = zv;
      if (range_flag && !(fabsf(zv) <= BF16_EXACT_INT)) *range_flag = 1;     // bf16 copy of z is no longer exact
      const float* k = kc + c * 12;
      const float p = cdf_fast(k, zv + 0.5f) - cdf_fast(k, zv - 0.5f);
      sum += fminf(fmaxf(-__log2f(p + 1e-10f), 0.f), 50.f);
    }
    tile[lane * TS + c] = __float2bfloat16_rn(zv);        // row stride C + 2: lanes (pixels) land in distinct banks
  }
  sum = block_sum(sum, red);                              // also a __syncthreads() for `tile`
  if (threadIdx.x == 0) partial[blockIdx.x] = sum;
  if (z_nhwc) {
    const int npix = min(QI_PIX, HW - pix0);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tile);
    uint32_t* dst = reinterpret_cast<uint32_t*>(z_nhwc + ((size_t)n * HW + pix0) * C);
    const int wpr = C / 2;     // 32-bit words per output row
    for (int i = threadIdx.x; i < npix * wpr; i += blockDim.x) {
      const int row = i / wpr, col = i - row * wpr;
      dst[i] = src[row * (TS / 2) + col];
    }
  }
}

// ---------------------------------------------------------------------------- table-driven inference variant
// At inference z = round(y) is an integer, so the bits of an element depend only on (channel, z): a per-channel table
// bits[c][z + L], |z| <= L, built once per weight version with the SAME per-element formula as the kernel above, turns
// the pass into pure data movement (read y, write z fp32 + z bf16, one cached table read) - the SFU chain (11 MUFU ops
// per element, which held the direct kernel at 36 % of HBM peak) only runs for the rare |z| > L.
__global__ void entropy_lut_build_kernel(BitPtrs bp, float* __restrict__ lut, int C, int L) {
  const int c = blockIdx.x;
  __shared__ float k[12];
  if (threadIdx.x == 0) {
    const ChanConst kc = load_chan(bp, c);
    k[0] = kc.sp[0]; k[1] = kc.b[0]; k[2] = kc.ta[0];
    k[3] = kc.sp[1]; k[4] = kc.b[1]; k[5] = kc.ta[1];
    k[6] = kc.sp[2]; k[7] = kc.b[2]; k[8] = kc.ta[2];
    k[9] = kc.sp[3]; k[10] = kc.b[3]; k[11] = 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * L + 1; i += blockDim.x) lut[(size_t)c * (2 * L + 1) + i] = elem_bits_fast(k, (float)(i - L));
}

// CTA = LQ_PIX (128) pixels x all channels of one image: 512 contiguous bytes per channel row on the NCHW side (32-pixel
// chunks = 128-byte pieces ran at 41 % of HBM peak), a 32 KB contiguous block on the NHWC side.  Thread = 4 consecutive
// pixels of one channel (LDG.128 / STG.128); lane = (pixel quad 0..7, channel 0..3); warp w covers the 32-pixel group
// w & 3 and the channel quartet (w >> 2) of every 8-channel pass.  HW % 4 == 0 (16-byte alignment of every row).
// (Tried: the table row in registers read with warp shuffles - 4 SHFL per element bound the kernel at 2.1 TB/s.)
constexpr int LQ_PIX = 128;

__global__ void __launch_bounds__(256) quant_entropy_infer_lut_kernel(const float* __restrict__ y, BitPtrs bp,
                                                                      const float* __restrict__ lut, int L,
                                                                      float* __restrict__ z,
                                                                      __nv_bfloat16* __restrict__ z_nhwc,
                                                                      float* __restrict__ partial, int C, int HW,
                                                                      int chunks, int* __restrict__ range_flag) {
  extern __shared__ float qi_smem[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(qi_smem);   // [LQ_PIX][C + 2]
  __shared__ float red[32];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
  const unsigned pix0 = chunk * LQ_PIX;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lp = (warp & 3) * 32 + (lane & 7) * 4;      // first of this thread's 4 pixels inside the chunk
  const unsigned c0 = (warp >> 2) * 4 + (lane >> 3);         // channel inside the 8-channel pass
  const unsigned TS = C + 2, W = 2 * L + 1;
  const bool in_image = pix0 + lp < (unsigned)HW;            // whole quads: HW % 4 == 0
  const float fL = (float)L;
  float sum = 0.f;
  const float* yp = y + ((size_t)n * C + c0) * HW + pix0 + lp;
  float* zp = z + ((size_t)n * C + c0) * HW + pix0 + lp;
  const float* row = lut + (size_t)c0 * W + L;
  __nv_bfloat16* tp = tile + lp * TS + c0;
  const size_t cstep = (size_t)8 * HW;
#pragma unroll 4
  for (unsigned c = c0; c < (unsigned)C; c += 8, yp += cstep, zp += cstep, row += 8 * W, tp += 8) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_image) {
      v = __ldg(reinterpret_cast<const float4*>(yp));
      v.x = rintf(v.x); v.y = rintf(v.y); v.z = rintf(v.z); v.w = rintf(v.w);
      *reinterpret_cast<float4*>(zp) = v;
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (fabsf(e[j]) <= fL) {
          sum += __ldg(row + (int)e[j]);
        } else {                                           // rare: outside the table, same formula evaluated directly
          sum += bits_outside_table(bp, (int)c, e[j]);
          if (range_flag && !(fabsf(e[j]) <= BF16_EXACT_INT)) *range_flag = 1;
        }
      }
    }
    tp[0] = __float2bfloat16_rn(v.x);                      // rows 4q .. 4q+3, row stride C + 2: conflict-free
    tp[TS] = __float2bfloat16_rn(v.y);
    tp[2 * TS] = __float2bfloat16_rn(v.z);
    tp[3 * TS] = __float2bfloat16_rn(v.w);
  }
  sum = block_sum(sum, red);                               // also the __syncthreads() for `tile`
  if (threadIdx.x == 0) partial[blockIdx.x] = sum;
  if (z_nhwc) {
    const unsigned npix = min((unsigned)LQ_PIX, (unsigned)HW - pix0);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tile);
    uint32_t* dst = reinterpret_cast<uint32_t*>(z_nhwc + ((size_t)n * HW + pix0) * C);
    const unsigned wpr = C / 2, tsw = TS / 2;
    for (unsigned r = warp; r < npix; r += 8)              // one warp per output row: coalesced, no divisions
      for (unsigned col = lane; col < wpr; col += 32) dst[r * wpr + col] = src[r * tsw + col];
  }
}

// Second version of the table-driven pass, for C % 64 == 0 (ncu source view of the first: ~70 issued instructions per
// 4 elements - a compare + branch + reconvergence point around EVERY table read, which also serialised the four reads
// behind each other's FADD; 2-byte shared-memory stores; 4-byte copy-out).  Here a thread owns 4 consecutive pixels of
// TWO adjacent channels: all 8 table reads are issued unconditionally on clamped indices (one rare fix-up branch per 8
// elements), the two channels of a pixel are packed into one bf16x2 word (4 STS.32 per 8 elements), and the NHWC tile
// is XOR-swizzled at 16-byte granularity (chunk ^ ((row >> 2) & 7)) instead of padded, so both the 4-byte writes and
// the 16-byte row reads of the copy-out are bank-conflict free and the copy-out is LDS.128 / STG.128.
// lane = (pixel quad q 0..7, pair k 0..3); warp w: pixel group w & 3, 8-channel chunk half w >> 2; a pass = 16 channels.
// NSUB = 2: 512 threads cover 256 pixels, so the two 512-byte halves of a 1 KB piece of every channel row are requested
// together (a DRAM page sees one visit instead of two: the 128-pixel form is bound by its access pattern, not by
// instructions).
template <int NSUB>
__global__ void __launch_bounds__(256 * NSUB) quant_entropy_infer_lut2_kernel(const float* __restrict__ y, BitPtrs bp,
                                                                       const float* __restrict__ lut, int L,
                                                                       float* __restrict__ z,
                                                                       __nv_bfloat16* __restrict__ z_nhwc,
                                                                       float* __restrict__ partial, int C, int HW,
                                                                       int chunks, int* __restrict__ range_flag,
                                                                       int stream_hint) {
  extern __shared__ float qi_smem[];
  const bool STREAM = stream_hint != 0;
  uint32_t* tile = reinterpret_cast<uint32_t*>(qi_smem);     // [LQ_PIX * NSUB][C / 2] bf16x2 words, 16-byte chunks swizzled
  __shared__ float red[32];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
  const unsigned pix0 = chunk * (LQ_PIX * NSUB);
  const unsigned lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7, sub = threadIdx.x >> 8;
  const unsigned q = lane & 7, k = lane >> 3;
  const unsigned lp = sub * LQ_PIX + (warp & 3) * 32 + q * 4;   // first of this thread's 4 pixels inside the chunk
  const unsigned half = warp >> 2;
  const unsigned wpr = C / 2, W = 2 * L + 1;                 // words per tile row; table row length
  const bool in_image = pix0 + lp < (unsigned)HW;            // whole quads: HW % 4 == 0
  const float fL = (float)L;
  float sum = 0.f;
  unsigned c = (half * 4 + k) * 2;                           // first channel of the pair, pass 0
  const float* yp = y + ((size_t)n * C + c) * HW + pix0 + lp;
  float* zp = z + ((size_t)n * C + c) * HW + pix0 + lp;
  const float* row = lut + (size_t)c * W + L;
  const size_t cstep = (size_t)16 * HW;
  uint32_t* tp = tile + lp * wpr + k;                        // + ((chunk ^ q) * 4) per pass; rows lp .. lp+3 share q
#pragma unroll 2
  for (unsigned pass = 0; pass < (unsigned)C / 16; ++pass, c += 16, yp += cstep, zp += cstep, row += 16 * W) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (in_image) {
      a = STREAM ? __ldcs(reinterpret_cast<const float4*>(yp)) : __ldg(reinterpret_cast<const float4*>(yp));
      b = STREAM ? __ldcs(reinterpret_cast<const float4*>(yp + HW)) : __ldg(reinterpret_cast<const float4*>(yp + HW));
      a.x = rintf(a.x); a.y = rintf(a.y); a.z = rintf(a.z); a.w = rintf(a.w);
      b.x = rintf(b.x); b.y = rintf(b.y); b.z = rintf(b.z); b.w = rintf(b.w);
      if (STREAM) {                                          // y is read once and z is not re-read by this kernel
        __stcs(reinterpret_cast<float4*>(zp), a);
        __stcs(reinterpret_cast<float4*>(zp + HW), b);
      } else {
        *reinterpret_cast<float4*>(zp) = a;
        *reinterpret_cast<float4*>(zp + HW) = b;
      }
      const float e[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      const float* row1 = row + W;
      float bits[8];
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        amax = fmaxf(amax, fabsf(e[j]));
        const int idx = (int)fminf(fmaxf(e[j], -fL), fL);
        bits[j] = __ldg((j < 4 ? row : row1) + idx);
      }
      if (!(amax <= fL)) {                                   // rare: outside the table (or NaN), same formula directly
        if (range_flag && !(amax <= BF16_EXACT_INT)) *range_flag = 1;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(fabsf(e[j]) <= fL)) bits[j] = bits_outside_table(bp, (int)c + (j >> 2), e[j]);
      }
      sum += ((bits[0] + bits[1]) + (bits[2] + bits[3])) + ((bits[4] + bits[5]) + (bits[6] + bits[7]));
    }
    uint32_t* t0 = tp + (((pass * 2 + half) ^ q) << 2);
    __nv_bfloat162 w0 = __floats2bfloat162_rn(a.x, b.x), w1 = __floats2bfloat162_rn(a.y, b.y);
    __nv_bfloat162 w2 = __floats2bfloat162_rn(a.z, b.z), w3 = __floats2bfloat162_rn(a.w, b.w);
    t0[0] = *reinterpret_cast<uint32_t*>(&w0);
    t0[wpr] = *reinterpret_cast<uint32_t*>(&w1);
    t0[2 * wpr] = *reinterpret_cast<uint32_t*>(&w2);
    t0[3 * wpr] = *reinterpret_cast<uint32_t*>(&w3);
  }
  sum = block_sum(sum, red);                                 // also the __syncthreads() for `tile`
  if (threadIdx.x == 0) partial[blockIdx.x] = sum;
  if (z_nhwc) {
    const unsigned npix = min((unsigned)(LQ_PIX * NSUB), (unsigned)HW - pix0);
    const unsigned cpr = C / 8;                              // 16-byte chunks per row
    const uint4* src = reinterpret_cast<const uint4*>(tile);
    uint4* dst = reinterpret_cast<uint4*>(z_nhwc + ((size_t)n * HW + pix0) * C);
    for (unsigned i = threadIdx.x; i < npix * cpr; i += 256 * NSUB) {
      const unsigned r = i / cpr, ch = i - r * cpr;
      dst[i] = src[r * cpr + (ch ^ ((r >> 2) & 7))];         // (the decoder reads this next: default caching)
    }
  }
}

// ---------------------------------------------------------------------------- backward
// For each branch u = z +- 1/2:  lambda_4 = sgn * dL/dp * F(1-F) ... chain back through the 3 layers.
struct Grad11 {
  float g[B2C_NUM_BITPARAMS];
};

__device__ __forceinline__ float branch_bwd(const ChanConst& k, const float sig_h[4], float u, float dLdF, Grad11& acc) {
  float v[4], t[4], th[3];
  v[0] = u;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    t[i] = fmaf(v[i], k.sp[i], k.b[i]);
    th[i] = tanhf(t[i]);
    v[i + 1] = fmaf(th[i], k.ta[i], t[i]);
  }
  t[3] = fmaf(v[3], k.sp[3], k.b[3]);
  const float F = sigmoid_t(t[3]);
  float dt = dLdF * F * (1.f - F);
  acc.g[9] += dt * v[3] * sig_h[3];  // dh4
  acc.g[10] += dt;                   // db4
  float dv = dt * k.sp[3];
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    acc.g[3 * i + 2] += dv * th[i] * (1.f - k.ta[i] * k.ta[i]);  // da_k
    dt = dv * (1.f + (1.f - th[i] * th[i]) * k.ta[i]);
    acc.g[3 * i] += dt * v[i] * sig_h[i];  // dh_k
    acc.g[3 * i + 1] += dt;                // db_k
    dv = dt * k.sp[i];
  }
  return dv;  // dL/du
}

__global__ void __launch_bounds__(256) entropy_bwd_kernel(const float* __restrict__ z, BitPtrs bp,
                                                          const float* __restrict__ g_bits, float* __restrict__ dz,
                                                          BitGradPtrs gp, int C, int HW) {
  __shared__ float red[32];
  const int plane = blockIdx.x;
  const int c = plane % C;
  const ChanConst k = load_chan(bp, c);
  float sig_h[4];
#pragma unroll
  for (int i = 0; i < 3; ++i) sig_h[i] = sigmoid_t(__ldg(bp.p[3 * i] + c));
  sig_h[3] = sigmoid_t(__ldg(bp.p[9] + c));
  // torch softplus backward passes the gradient unchanged above the threshold
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (__ldg(bp.p[3 * i] + c) > 20.f) sig_h[i] = 1.f;
  if (__ldg(bp.p[9] + c) > 20.f) sig_h[3] = 1.f;
  const float gb = __ldg(g_bits);
  const size_t base = (size_t)plane * HW;
  Grad11 acc;
#pragma unroll
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) acc.g[i] = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float zv = __ldg(z + base + i);
    const float p = cdf(k, zv + 0.5f) - cdf(k, zv - 0.5f);
    const float q = p + 1e-10f;
    const float bits = -logf(q) / 0.6931471805599453f;
    // clamp(.,0,50) passes the gradient on the closed interval
    const float dLdp = (bits >= 0.f && bits <= 50.f) ? -gb / (q * 0.6931471805599453f) : 0.f;
    float d = 0.f;
    d += branch_bwd(k, sig_h, zv + 0.5f, dLdp, acc);
    d += branch_bwd(k, sig_h, zv - 0.5f, -dLdp, acc);
    dz[base + i] = d;
  }
#pragma unroll
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) {
    const float s = block_sum(acc.g[i], red);
    if (threadIdx.x == 0) atomicAdd(gp.p[i] + c, s);
  }
}

// --------------------------------------------------------------------------- distortion
__global__ void __launch_bounds__(256) clamp_sse_fwd_kernel(const float* __restrict__ recon,
                                                            const float* __restrict__ x, float* __restrict__ clipped,
                                                            float* __restrict__ partial, int CHW, int blocks_per_image) {
  __shared__ float red[32];
  const int n = blockIdx.x / blocks_per_image, b = blockIdx.x - n * blocks_per_image;
  const size_t base = (size_t)n * CHW;
  float s0 = 0.f, s1 = 0.f;
  for (int i = b * blockDim.x + threadIdx.x; i < CHW; i += blocks_per_image * blockDim.x) {
    const float r = __ldg(recon + base + i), xv = __ldg(x + base + i);
    const float c = fminf(fmaxf(r, 0.f), 1.f);
    clipped[base + i] = c;
    const float d0 = r - xv, d1 = c - xv;
    s0 = fmaf(d0, d0, s0);
    s1 = fmaf(d1, d1, s1);
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  if (threadIdx.x == 0) {
    partial[(size_t)blockIdx.x * 2 + 0] = s0;
    partial[(size_t)blockIdx.x * 2 + 1] = s1;
  }
}

__global__ void __launch_bounds__(256) sse_reduce_kernel(const float* __restrict__ partial, float* __restrict__ sse,
                                                         int blocks_per_image) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  float s0 = 0.f, s1 = 0.f;
  for (int i = threadIdx.x; i < blocks_per_image; i += blockDim.x) {
    s0 += partial[((size_t)n * blocks_per_image + i) * 2 + 0];
    s1 += partial[((size_t)n * blocks_per_image + i) * 2 + 1];
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  if (threadIdx.x == 0) {
    sse[2 * n + 0] = s0;
    sse[2 * n + 1] = s1;
  }
}

__global__ void __launch_bounds__(256) clamp_sse_bwd_kernel(const float* __restrict__ recon,
                                                            const float* __restrict__ x,
                                                            const float* __restrict__ g_clipped,
                                                            const float* __restrict__ g_sse, float* __restrict__ g_recon,
                                                            size_t n) {
  const float gs = 2.f * __ldg(g_sse);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float r = __ldg(recon + i);
    float g = gs * (r - __ldg(x + i));
    if (g_clipped && r >= 0.f && r <= 1.f) g += __ldg(g_clipped + i);
    g_recon[i] = g;
  }
}

__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ partial,
                                                             float* __restrict__ out, int per_image, int width) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // no-ops unless launched as a programmatic dependent
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (width == 2) {        // the squared-error pairs of the last synthesis layer: both columns per 8-byte load, eight
    const float2* src = reinterpret_cast<const float2*>(partial) + (size_t)n * per_image;   // loads in flight per thread
    float s0 = 0.f, s1 = 0.f;
    for (int i0 = threadIdx.x; i0 < per_image; i0 += 8 * blockDim.x) {
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + k * blockDim.x;
        v[k] = i < per_image ? __ldg(src + i) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { s0 += v[k].x; s1 += v[k].y; }
    }
    s0 = block_sum(s0, red);
    s1 = block_sum(s1, red);
    if (threadIdx.x == 0) { out[n * 2] = s0; out[n * 2 + 1] = s1; }
    return;
  }
  for (int j = 0; j < width; ++j) {
    float s = 0.f;
    for (int i = threadIdx.x; i < per_image; i += blockDim.x) s += partial[((size_t)n * per_image + i) * width + j];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[n * width + j] = s;
  }
}


// ---------------------------------------------------------------------------- end of an inference forward
// model.py:61,73,78 in ONE launch: block n reduces image n's per-tile squared-error partial sums in a fixed order
// (as partial_reduce_kernel, width 2); the last block to finish - a ticket in device memory, reset for the next
// launch - adds the per-image bits and squared errors in index order and forms mse and bpp.  The eval forward used to
// end with five small launches for this (partial reduction, two sums, two divisions): ~15 us of a 0.68 ms step.
__global__ void __launch_bounds__(256) infer_stats_kernel(const float* __restrict__ partial, int per_image,
                                                          float* __restrict__ image_bits,
                                                          const float* __restrict__ bits_partial, int bits_slots,
                                                          float numel, float pixels, float* __restrict__ sse,
                                                          float* __restrict__ totals, unsigned* __restrict__ ticket) {
  __shared__ float red[32];
  __shared__ unsigned s_last;
  const int n = blockIdx.x, B = gridDim.x;
  if (bits_partial) {        // image n's bits from the per-(tile, warp) sums of the last analysis layer, fixed order
    float b = 0.f;
    for (int i = threadIdx.x; i < bits_slots; i += blockDim.x) b += __ldg(bits_partial + (size_t)n * bits_slots + i);
    b = block_sum(b, red);
    if (threadIdx.x == 0) image_bits[n] = b;
  }
  const float2* src = reinterpret_cast<const float2*>(partial) + (size_t)n * per_image;
  float s0 = 0.f, s1 = 0.f;
  for (int i0 = threadIdx.x; i0 < per_image; i0 += 8 * blockDim.x) {
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * blockDim.x;
      v[k] = i < per_image ? __ldg(src + i) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { s0 += v[k].x; s1 += v[k].y; }
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  if (threadIdx.x == 0) {
    sse[n * 2] = s0;
    sse[n * 2 + 1] = s1;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == (unsigned)(B - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {                             // fixed summation tree: the same sums whichever block comes last
    __threadfence();
    float bits = 0.f, e0 = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      bits += __ldcg(image_bits + i);
      e0 += __ldcg(sse + 2 * i);
    }
    bits = block_sum(bits, red);
    e0 = block_sum(e0, red);
    if (threadIdx.x == 0) {
      totals[0] = bits;
      totals[1] = e0;
      totals[2] = e0 / numel;             // mse on the unclamped reconstruction (model.py:61)
      totals[3] = bits / pixels;          // bpp (model.py:78)
      *ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------- Bitparm as a module call
// models/bitEstimator.py:20-25 for callers that evaluate the CDF themselves (model.py:71-72 pattern
// bitEstimator(z + .5) - bitEstimator(z - .5), per-channel tables for an entropy coder).  Exact fp32 formula, one CTA
// per (image, channel) plane; the backward writes per-plane parameter-gradient partials that a fixed-order second
// kernel sums over the batch (deterministic).
__global__ void __launch_bounds__(256) bitparm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                          const float* __restrict__ b, const float* __restrict__ a,
                                                          float* __restrict__ out, int C, int HW) {
  const int plane = blockIdx.x, c = plane % C;
  const float sp = softplus_t(__ldg(h + c)), bb = __ldg(b + c);
  const float ta = a ? tanhf(__ldg(a + c)) : 0.f;
  const size_t base = (size_t)plane * HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float t = fmaf(__ldg(x + base + i), sp, bb);
    out[base + i] = a ? fmaf(tanhf(t), ta, t) : sigmoid_t(t);
  }
}

__global__ void __launch_bounds__(256) bitparm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                          const float* __restrict__ b, const float* __restrict__ a,
                                                          const float* __restrict__ g, float* __restrict__ dx,
                                                          float* __restrict__ part, int C, int HW, int planes) {
  __shared__ float red[32];
  const int plane = blockIdx.x, c = plane % C;
  const float hh = __ldg(h + c);
  const float sp = softplus_t(hh), bb = __ldg(b + c), sg = sigmoid_t(hh);      // d softplus(h) / dh = sigmoid(h)
  const float ta = a ? tanhf(__ldg(a + c)) : 0.f;
  const size_t base = (size_t)plane * HW;
  float sh = 0.f, sb = 0.f, sa = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float xv = __ldg(x + base + i), gv = __ldg(g + base + i);
    const float t = fmaf(xv, sp, bb);
    float gt;                                                                  // d loss / d t
    if (a) {
      const float th = tanhf(t);
      gt = gv * fmaf(1.f - th * th, ta, 1.f);
      sa += gv * th;
    } else {
      const float s = sigmoid_t(t);
      gt = gv * s * (1.f - s);
    }
    dx[base + i] = gt * sp;
    sb += gt;
    sh += gt * xv;
  }
  sh = block_sum(sh, red);
  sb = block_sum(sb, red);
  sa = block_sum(sa, red);
  if (threadIdx.x == 0) {
    part[plane] = sh * sg;
    part[planes + plane] = sb;
    part[2 * planes + plane] = sa * (1.f - ta * ta);
  }
}

// out_k[c] = sum_n part[k][n*C + c], k = h, b, a (fixed order)
__global__ void bitparm_param_reduce_kernel(const float* __restrict__ part, float* __restrict__ dh,
                                            float* __restrict__ db, float* __restrict__ da, int B, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int planes = B * C;
  float sh = 0.f, sb = 0.f, sa = 0.f;
  for (int n = 0; n < B; ++n) {
    sh += part[n * C + c];
    sb += part[planes + n * C + c];
    sa += part[2 * planes + n * C + c];
  }
  dh[c] = sh;
  db[c] = sb;
  if (da) da[c] = sa;
}

static int clamp_blocks_per_image(int B, int CHW) {
  int bpi = (CHW + 256 * 8 - 1) / (256 * 8);
  const int cap = (148 * 8 + B - 1) / B;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  return bpi;
}

}  // namespace b2c

using namespace b2c;

extern "C" {

int b2c_quant_entropy_fwd(const float* y, const float* noise, const float* const* host_params, float* z,
                          float* plane_bits, float* image_bits, int B, int C, int HW, void* stream) {
  B2C_REQUIRE(y && host_params && z && plane_bits && image_bits && B > 0 && C > 0 && HW > 0,
              "quant_entropy_fwd: bad arguments");
  BitPtrs bp;
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) {
    B2C_REQUIRE(host_params[i], "quant_entropy_fwd: null parameter pointer %d", i);
    bp.p[i] = host_params[i];
  }
  launch_pdl(PDL_ENTROPY, quant_entropy_fwd_kernel, dim3(B * C), dim3(256), 0, (cudaStream_t)stream, y, noise, bp, z, plane_bits, C, HW);
  launch_pdl(PDL_ENTROPY, plane_reduce_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, (const float*)plane_bits, image_bits, C);
  count_launch(2);
  return check_launch("quant_entropy_fwd");
}

size_t b2c_quant_entropy_infer_workspace_bytes(int B, int HW) {
  return (size_t)B * ((HW + QI_PIX - 1) / QI_PIX) * sizeof(float);
}

int b2c_quant_entropy_infer(const float* y, const float* const* host_params, float* z, void* z_nhwc_bf16,
                            float* image_bits, int* range_flag, int B, int C, int HW, void* workspace,
                            size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(y && host_params && z && image_bits && workspace && B > 0 && C > 0 && C % 2 == 0 && HW > 0,
              "quant_entropy_infer: bad arguments");
  if (workspace_bytes < b2c_quant_entropy_infer_workspace_bytes(B, HW)) {
    set_error("quant_entropy_infer: workspace too small");
    return B2C_EWORKSPACE;
  }
  BitPtrs bp;
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) {
    B2C_REQUIRE(host_params[i], "quant_entropy_infer: null parameter pointer %d", i);
    bp.p[i] = host_params[i];
  }
  const int chunks = (HW + QI_PIX - 1) / QI_PIX;
  const size_t smem = (size_t)C * 12 * sizeof(float) + (size_t)QI_PIX * (C + 2) * 2;
  B2C_REQUIRE(smem <= 48 * 1024, "quant_entropy_infer: too many channels");
  {   // both kernels are programmatic dependents of their predecessor (B2C_PDL=0: plain stream order)
    static const bool pdl = [] { const char* e = getenv("B2C_PDL"); return !e || atoi(e) != 0; }();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * chunks));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, quant_entropy_infer_kernel, y, bp, z, (__nv_bfloat16*)z_nhwc_bf16,
                                       (float*)workspace, C, HW, chunks, range_flag);
    if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(quant_entropy_infer): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
    cfg.gridDim = dim3((unsigned)B);
    cfg.dynamicSmemBytes = 0;
    e = cudaLaunchKernelEx(&cfg, partial_reduce_kernel, (const float*)workspace, image_bits, chunks, 1);
    if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(partial_reduce): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  }
  count_launch(2);
  return check_launch("quant_entropy_infer");
}

int b2c_entropy_lut_build(const float* const* host_params, float* lut, int C, int half_range, void* stream) {
  B2C_REQUIRE(host_params && lut && C > 0 && half_range >= 1 && half_range <= 1024, "entropy_lut_build: bad arguments");
  BitPtrs bp;
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) {
    B2C_REQUIRE(host_params[i], "entropy_lut_build: null parameter pointer %d", i);
    bp.p[i] = host_params[i];
  }
  entropy_lut_build_kernel<<<C, 128, 0, (cudaStream_t)stream>>>(bp, lut, C, half_range);
  count_launch();
  return check_launch("entropy_lut_build");
}

int b2c_quant_entropy_infer_lut(const float* y, const float* const* host_params, const float* lut, int half_range,
                                float* z, void* z_nhwc_bf16, float* image_bits, int* range_flag, int B, int C, int HW,
                                void* workspace, size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(y && host_params && lut && z && image_bits && workspace && B > 0 && C > 0 && C % 8 == 0 && HW > 0 &&
                  HW % 4 == 0 && half_range >= 1,
              "quant_entropy_infer_lut: bad arguments (C %% 8 == 0 and HW %% 4 == 0 required)");
  if (workspace_bytes < b2c_quant_entropy_infer_workspace_bytes(B, HW)) {
    set_error("quant_entropy_infer_lut: workspace too small");
    return B2C_EWORKSPACE;
  }
  BitPtrs bp;
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) {
    B2C_REQUIRE(host_params[i], "quant_entropy_infer_lut: null parameter pointer %d", i);
    bp.p[i] = host_params[i];
  }
  static const bool allow_v2 = [] { const char* e = getenv("B2C_ENTROPY_LUT_V1"); return !e || atoi(e) == 0; }();
  static const int nsub_env = [] { const char* e = getenv("B2C_ENTROPY_LUT_SUB"); return e ? atoi(e) : 2; }();
  const bool v2 = allow_v2 && C % 64 == 0 && ((uintptr_t)z_nhwc_bf16 & 15) == 0;
  // 256-pixel chunks (512 threads) when that still leaves >= 2 CTAs per SM: on the 18.9 MB latents of a Kodak batch
  // the larger CTAs (144 of them) measured slower than 288 of 128 pixels
  const int nsub = (v2 && nsub_env == 2 && (long long)B * ((HW + 2 * LQ_PIX - 1) / (2 * LQ_PIX)) >= 2 * 148) ? 2 : 1;
  const int chunks = (HW + LQ_PIX * nsub - 1) / (LQ_PIX * nsub);
  const size_t smem = v2 ? (size_t)LQ_PIX * nsub * C * 2 : (size_t)LQ_PIX * (C + 2) * 2;
  {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(quant_entropy_infer_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           LQ_PIX * (256 + 2) * 2);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      e = cudaFuncSetAttribute(quant_entropy_infer_lut2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               LQ_PIX * 256 * 2);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      e = cudaFuncSetAttribute(quant_entropy_infer_lut2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               2 * LQ_PIX * 256 * 2);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
  }
  B2C_REQUIRE(C <= 256, "quant_entropy_infer_lut: at most 256 channels");
  static const bool pdl = [] { const char* e = getenv("B2C_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * chunks));
  cfg.blockDim = dim3(256 * nsub);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  static const int stream_hint = [] { const char* e = getenv("B2C_ENTROPY_STREAM"); return e ? atoi(e) : 1; }();
  cudaError_t e;
  if (v2 && nsub == 2)
    e = cudaLaunchKernelEx(&cfg, quant_entropy_infer_lut2_kernel<2>, y, bp, lut, half_range, z, (__nv_bfloat16*)z_nhwc_bf16,
                           (float*)workspace, C, HW, chunks, range_flag, stream_hint);
  else if (v2)
    e = cudaLaunchKernelEx(&cfg, quant_entropy_infer_lut2_kernel<1>, y, bp, lut, half_range, z, (__nv_bfloat16*)z_nhwc_bf16,
                           (float*)workspace, C, HW, chunks, range_flag, stream_hint);
  else
    e = cudaLaunchKernelEx(&cfg, quant_entropy_infer_lut_kernel, y, bp, lut, half_range, z, (__nv_bfloat16*)z_nhwc_bf16,
                           (float*)workspace, C, HW, chunks, range_flag);
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(quant_entropy_infer_lut): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  cfg.gridDim = dim3((unsigned)B);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  e = cudaLaunchKernelEx(&cfg, partial_reduce_kernel, (const float*)workspace, image_bits, chunks, 1);
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(partial_reduce): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  count_launch(2);
  return check_launch("quant_entropy_infer_lut");
}

int b2c_entropy_bwd(const float* z, const float* const* host_params, const float* g_bits, float* dz,
                    float* const* host_dparams, int B, int C, int HW, void* stream) {
  B2C_REQUIRE(z && host_params && g_bits && dz && host_dparams && B > 0 && C > 0 && HW > 0,
              "entropy_bwd: bad arguments");
  BitPtrs bp;
  BitGradPtrs gp;
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) {
    B2C_REQUIRE(host_params[i] && host_dparams[i], "entropy_bwd: null parameter pointer %d", i);
    bp.p[i] = host_params[i];
    gp.p[i] = host_dparams[i];
  }
  entropy_bwd_kernel<<<B * C, 256, 0, (cudaStream_t)stream>>>(z, bp, g_bits, dz, gp, C, HW);
  count_launch();
  return check_launch("entropy_bwd");
}

size_t b2c_clamp_sse_workspace_bytes(int B, int CHW) {
  return (size_t)B * clamp_blocks_per_image(B, CHW) * 2 * sizeof(float);
}

int b2c_clamp_sse_fwd(const float* recon, const float* x, float* clipped, float* sse, int B, int CHW,
                      void* workspace, size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(recon && x && clipped && sse && workspace && B > 0 && CHW > 0, "clamp_sse_fwd: bad arguments");
  if (workspace_bytes < b2c_clamp_sse_workspace_bytes(B, CHW)) {
    set_error("clamp_sse_fwd: workspace too small");
    return B2C_EWORKSPACE;
  }
  const int bpi = clamp_blocks_per_image(B, CHW);
  clamp_sse_fwd_kernel<<<B * bpi, 256, 0, (cudaStream_t)stream>>>(recon, x, clipped, (float*)workspace, CHW, bpi);
  sse_reduce_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, sse, bpi);
  count_launch(2);
  return check_launch("clamp_sse_fwd");
}

int b2c_partial_reduce(const float* partial, float* out, int B, int per_image, int width, void* stream) {
  B2C_REQUIRE(partial && out && B > 0 && per_image > 0 && width > 0 && width <= 4, "partial_reduce: bad arguments");
  partial_reduce_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(partial, out, per_image, width);
  count_launch();
  return check_launch("partial_reduce");
}

int b2c_infer_stats(const float* sse_partial, int B, int tiles, float* image_bits, const float* bits_partial,
                    int bits_slots, double numel, double pixels, float* sse, float* totals, unsigned* ticket,
                    void* stream) {
  B2C_REQUIRE(sse_partial && image_bits && sse && totals && ticket && B > 0 && B <= 65535 && tiles > 0 && numel > 0 &&
                  pixels > 0 && (!bits_partial || bits_slots > 0), "infer_stats: bad arguments");
  infer_stats_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(sse_partial, tiles, image_bits, bits_partial,
                                                                    bits_slots, (float)numel, (float)pixels, sse, totals,
                                                                    ticket);
  count_launch();
  return check_launch("infer_stats");
}

int b2c_clamp_sse_bwd(const float* recon, const float* x, const float* g_clipped, const float* g_sse,
                      float* g_recon, size_t n, void* stream) {
  B2C_REQUIRE(recon && x && g_sse && g_recon && n > 0, "clamp_sse_bwd: bad arguments");
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  clamp_sse_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(recon, x, g_clipped, g_sse, g_recon, n);
  count_launch();
  return check_launch("clamp_sse_bwd");
}

int b2c_bitparm_fwd(const float* x, const float* h, const float* b, const float* a, float* out, int B, int C, int HW,
                    void* stream) {
  B2C_REQUIRE(x && h && b && out && B > 0 && C > 0 && HW > 0, "bitparm_fwd: bad arguments");
  bitparm_fwd_kernel<<<B * C, 256, 0, (cudaStream_t)stream>>>(x, h, b, a, out, C, HW);
  count_launch();
  return check_launch("bitparm_fwd");
}

size_t b2c_bitparm_bwd_workspace_bytes(int B, int C) { return (size_t)3 * B * C * sizeof(float); }

int b2c_bitparm_bwd(const float* x, const float* h, const float* b, const float* a, const float* g, float* dx,
                    float* dh, float* db, float* da, int B, int C, int HW, void* workspace, size_t workspace_bytes,
                    void* stream) {
  B2C_REQUIRE(x && h && b && g && dx && dh && db && B > 0 && C > 0 && HW > 0, "bitparm_bwd: bad arguments");
  B2C_REQUIRE((a == nullptr) == (da == nullptr), "bitparm_bwd: a and da must both be given (or both NULL for the final layer)");
  if (!workspace || workspace_bytes < b2c_bitparm_bwd_workspace_bytes(B, C)) {
    set_error("bitparm_bwd: workspace too small");
    return B2C_EWORKSPACE;
  }
  float* part = static_cast<float*>(workspace);
  bitparm_bwd_kernel<<<B * C, 256, 0, (cudaStream_t)stream>>>(x, h, b, a, g, dx, part, C, HW, B * C);
  bitparm_param_reduce_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(part, dh, db, da, B, C);
  count_launch(2);
  return check_launch("bitparm_bwd");
}

}  // extern "C"
