// Training input transform on the device: the reference's per-image CPU pipeline
//     RandomResizedCrop(image_size) -> RandomHorizontalFlip -> RandomVerticalFlip -> ToTensor        (datasets.py:24-34)
// as ONE launch over a batch of decoded 8-bit RGB images resident in HBM.  The crop boxes and flips are drawn on the
// host (same torch RNG calls as torchvision); the arithmetic follows Pillow's ImagingResample (src/libImaging/
// Resample.c), which torchvision's resized_crop runs through img.crop(box).resize(size, BILINEAR): separable
// antialiased bilinear resampling of 8-bit channels, filter taps in 22-bit fixed point computed in IEEE double,
// horizontal pass first, each pass rounded (+2^21, >> 22) and clipped to uint8; ToTensor divides by 255 in fp32.
// Integer work: bit-exact against the real packages (tests/golden/augment_ref.npz) and oracle/resample_oracle.py.
//
// One CTA = AUG_T output rows of one image, one thread per output column.  Both passes are fused: the CTA streams the
// crop rows its AUG_T outputs touch through shared memory (16-byte cp.async of the aligned span, double buffered),
// each thread resamples the row horizontally for its column (its taps in shared memory, laid out [tap][column]),
// rounds to uint8 exactly as the intermediate image of the two-pass original would hold it, and accumulates that value
// into the (at most AUG_T) vertical sums whose windows contain the row.  HBM traffic: every crop byte read once per
// row tile that needs it (window overlap between neighbouring tiles is served by L2) + 12 B per output pixel written.
#include "b2c_common.cuh"

namespace b2c {

constexpr int AUG_T = 8;
constexpr int AUG_PREC = 22;            // Resample.c PRECISION_BITS = 32 - 8 - 2
constexpr int AUG_ST_BAD = 1;           // status bit: descriptor outside its image / above the launch maxima

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output index xx of out_size over [0, in_size), bilinear
// (support 1).  Explicit round-to-nearest intrinsics: no FMA contraction, so the doubles equal the host library's.
__device__ __forceinline__ void aug_taps(int in_size, int out_size, int xx, int& first, int& count, int* k, int stride) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = fscale;
  const double ss = __ddiv_rn(1.0, fscale);
  const double center = __dmul_rn((double)xx + 0.5, scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  const int n = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    double v = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
    ww = __dadd_rn(ww, v < 1.0 ? __dsub_rn(1.0, v) : 0.0);
  }
  for (int x = 0; x < n; ++x) {
    double v = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
    double w = v < 1.0 ? __dsub_rn(1.0, v) : 0.0;
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    k[x * stride] = (int)__dadd_rn(0.5, __dmul_rn(w, (double)(1 << AUG_PREC)));      // bilinear taps are >= 0
  }
  first = xmin;
  count = n;
}

__device__ __forceinline__ int aug_clip8(int acc) { return min(max(acc >> AUG_PREC, 0), 255); }

__device__ __forceinline__ void aug_cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

// desc[n] = {src byte offset, H, W, top, left, crop_h, crop_w, flags}; out fp32 [B][3][S][S]
__global__ void __launch_bounds__(512) resized_crop_flip_kernel(const uint8_t* __restrict__ src, size_t src_bytes,
                                                                const long long* __restrict__ desc, int S, int kh_cap,
                                                                int kv_cap, int row_cap, float* __restrict__ out,
                                                                int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t aug_smem[];
  int* kh = (int*)aug_smem;                         // [kh_cap][S]
  int* kv = kh + (size_t)kh_cap * S;                // [AUG_T][kv_cap]
  int* bv = kv + AUG_T * kv_cap;                    // [AUG_T][2]
  uint8_t* rows = (uint8_t*)(bv + 2 * AUG_T);       // 2 x row_cap (row_cap % 16 == 0; the int counts above are multiples of 4)

  const int xx = threadIdx.x, n = blockIdx.y, tile = blockIdx.x;
  const long long* d = desc + (size_t)n * 8;
  const size_t off = (size_t)d[0];
  const int H = (int)d[1], W = (int)d[2], top = (int)d[3], left = (int)d[4], ch = (int)d[5], cw = (int)d[6];
  const int flags = (int)d[7];
  const bool ok = d[0] >= 0 && (d[0] & 15) == 0 && H > 0 && W > 0 && top >= 0 && left >= 0 && ch > 0 && cw > 0 &&
                  (long long)top + ch <= H && (long long)left + cw <= W && off + (size_t)H * W * 3 <= src_bytes &&
                  (size_t)cw * 3 + 32 <= (size_t)row_cap;
  int hfirst = 0, hcount = 0;
  bool fits = ok;
  if (ok) {
    // tap counts are bounded by ceil(max(scale, 1)) * 2 + 1 (Resample.c ksize)
    const int need_h = (cw + S - 1) / S * 2 + 1, need_v = (ch + S - 1) / S * 2 + 1;
    fits = need_h <= kh_cap && need_v <= kv_cap;
  }
  if (!fits) {                                      // uniform over the CTA: descriptor data only
    if (xx == 0 && tile == 0) atomicOr(status, AUG_ST_BAD);
    return;
  }
  aug_taps(cw, S, xx, hfirst, hcount, kh + xx, S);
  if (xx < AUG_T) {
    int f, c;
    aug_taps(ch, S, tile * AUG_T + xx, f, c, kv + xx * kv_cap, 1);
    bv[2 * xx] = f;
    bv[2 * xx + 1] = c;
  }
  __syncthreads();
  int vfirst[AUG_T], vcount[AUG_T], acc[AUG_T][3];
#pragma unroll
  for (int t = 0; t < AUG_T; ++t) {
    vfirst[t] = bv[2 * t];
    vcount[t] = bv[2 * t + 1];
    acc[t][0] = acc[t][1] = acc[t][2] = 1 << (AUG_PREC - 1);
  }
  const int y_begin = vfirst[0], y_end = vfirst[AUG_T - 1] + vcount[AUG_T - 1];     // windows are monotone in the row

  auto issue = [&](int y, int buf) {
    const size_t g0 = off + ((size_t)(top + y) * W + left) * 3;
    const size_t a0 = g0 & ~(size_t)15, a1 = (g0 + (size_t)cw * 3 + 15) & ~(size_t)15;
    uint8_t* dst = rows + (size_t)buf * row_cap;
    for (size_t a = a0 + (size_t)xx * 16; a < a1; a += (size_t)blockDim.x * 16) {
      if (a + 16 <= src_bytes) {
        aug_cp16(dst + (a - a0), src + a);
      } else {                                      // tail of the whole buffer: byte-wise, zero filled
        for (int b = 0; b < 16; ++b) dst[a - a0 + b] = a + b < src_bytes ? src[a + b] : (uint8_t)0;
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  issue(y_begin, 0);
  for (int y = y_begin; y < y_end; ++y) {
    const int buf = (y - y_begin) & 1;
    if (y + 1 < y_end) {
      issue(y + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    const size_t g0 = off + ((size_t)(top + y) * W + left) * 3;
    const uint8_t* r = rows + (size_t)buf * row_cap + (g0 & 15) + (size_t)hfirst * 3;
    int s0 = 1 << (AUG_PREC - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < hcount; ++t) {
      const int k = kh[t * S + xx];
      s0 += (int)r[3 * t] * k;
      s1 += (int)r[3 * t + 1] * k;
      s2 += (int)r[3 * t + 2] * k;
    }
    const int h0 = aug_clip8(s0), h1 = aug_clip8(s1), h2 = aug_clip8(s2);
#pragma unroll
    for (int t = 0; t < AUG_T; ++t) {
      const int dy = y - vfirst[t];
      if ((unsigned)dy < (unsigned)vcount[t]) {
        const int k = kv[t * kv_cap + dy];
        acc[t][0] += h0 * k;
        acc[t][1] += h1 * k;
        acc[t][2] += h2 * k;
      }
    }
    __syncthreads();                                // the buffer read here is refilled by the next iteration's issue
  }

  const int ox = (flags & 1) ? S - 1 - xx : xx;
#pragma unroll
  for (int t = 0; t < AUG_T; ++t) {
    const int oy0 = tile * AUG_T + t, oy = (flags & 2) ? S - 1 - oy0 : oy0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[(((size_t)n * 3 + c) * S + oy) * S + ox] = __fdiv_rn((float)aug_clip8(acc[t][c]), 255.f);   // ToTensor
  }
}

struct AugGeom {
  int kh_cap, kv_cap, row_cap;
  size_t smem;
};
static bool aug_geom(int S, int max_crop_h, int max_crop_w, AugGeom& g) {
  if (S < 32 || S > 512 || S % 32 || max_crop_h <= 0 || max_crop_w <= 0 || max_crop_h > (1 << 20) || max_crop_w > (1 << 20))
    return false;
  g.kh_cap = (max_crop_w + S - 1) / S * 2 + 1;
  g.kv_cap = ((max_crop_h + S - 1) / S * 2 + 1 + 3) & ~3;
  g.row_cap = (max_crop_w * 3 + 32 + 15) & ~15;
  g.smem = ((size_t)g.kh_cap * S + (size_t)AUG_T * g.kv_cap + 2 * AUG_T) * sizeof(int) + 2 * (size_t)g.row_cap;
  return true;
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_resized_crop_flip(const uint8_t* src, size_t src_bytes, const int64_t* desc, int B, int out_size,
                                     int max_crop_h, int max_crop_w, float* out, int* status, void* stream) {
  AugGeom g;
  B2C_REQUIRE(src && desc && out && status && B > 0 && B <= 65535 && aug_geom(out_size, max_crop_h, max_crop_w, g),
              "resized_crop_flip: bad arguments (out_size a multiple of 32 in [32, 512], positive crop maxima)");
  B2C_REQUIRE(((uintptr_t)src & 15) == 0, "resized_crop_flip: the image buffer must be 16-byte aligned");
  B2C_REQUIRE(g.smem <= 227 * 1024, "resized_crop_flip: crops up to %d x %d need %zu bytes of shared memory (limit 232448)",
              max_crop_h, max_crop_w, g.smem);
  if (g.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(resized_crop_flip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("resized_crop_flip: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return B2C_ECUDA;
    }
  }
  resized_crop_flip_kernel<<<dim3(out_size / AUG_T, B), out_size, g.smem, (cudaStream_t)stream>>>(
      src, src_bytes, (const long long*)desc, out_size, g.kh_cap, g.kv_cap, g.row_cap, out, status);
  count_launch();
  return check_launch("resized_crop_flip");
}
