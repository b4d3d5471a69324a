// tcgen05 weight-gradient kernels for the bf16 training path (out_channel_N = 128).
//
//   dW[m][n][tap] += sum_{pixels} S[pix][m] * G[stride*pix - pad + tap][n]
//
// Both operands are NHWC bf16 activations, i.e. the reduction index (pixels) is the row index of
// the TMA-staged tiles and the channel index is contiguous: the MMA operands are therefore
// MN-major (UMMA a_major = b_major = 1) SWIZZLE_128B tiles, 16 pixels (K = 16) per instruction,
// M = N = 128 channels = two 64-channel boxes 16 KB apart.  The tiles are the very TMA boxes the
// forward convolution loads (traversal stride 2 for the stride-2 layers, OOB zero-fill = padding
// and tile masking).  One CTA owns up to 4 taps (4 x 128 TMEM columns) and a strided share of the
// pixel tiles; partial sums are added to the fp32 gradient with red.global.
//
// Covers conv2/conv3/deconv1/deconv2 weight gradients (25 taps, stride 2) and the GDN gamma
// gradient d gamma[i][j] = sum_pix dn[pix][i] * x^2[pix][j] (1 tap, stride 1).
#include <cstdlib>
#include <mutex>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"

namespace b2c {
namespace tc {

constexpr int W_TH = 8, W_TW = 16;
constexpr int W_PAIR_BYTES = 2 * 128 * 64 * 2;  // two 64-channel boxes of 128 pixels: 32 KB
constexpr int W_STAGES = 6;
constexpr int W_TAPS = 4;                       // taps per CTA (4 x 128 fp32 columns = all of TMEM)
constexpr int W_THREADS = 192;                  // producer warp, MMA warp, 4 epilogue warps
constexpr int W_ROW_PAD = 528;                  // bytes between staged output rows (512 + 16: conflict-free 16-byte stores)

struct WgParams {
  int B, hs, ws;          // small (tile) grid
  int tiles_x, tiles_per_image, total_tiles;
  int ksz, stride, pad, ntaps;
  float* dW;              // [ntaps][128][128] (tap-major), accumulated with red.global.add when partial == nullptr
  float* partial;         // two-stage reduction: [nsplit][ntaps][128 n][128 m] per-CTA sums (no atomics), or nullptr
  int bulk_reduce;        // 1: 512-byte bulk reductions from staged rows, 0: red.global.add.v4.f32 from registers
};

// MN-major SWIZZLE_128B operand: 64 channels (128 B) x 8 pixels per swizzle atom; atoms along the channel
// (M/N) direction are `lbo` bytes apart, atoms along the pixel (K) direction 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(W_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmG, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W_STAGES * W_PAIR_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + W_STAGES;
  uint64_t* done = bars + 2 * W_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * W_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap0 = blockIdx.x * W_TAPS;
  const int ntap = min(W_TAPS, p.ntaps - tap0);
  const int split = blockIdx.y, nsplit = gridDim.y;
  const int my_tiles = (p.total_tiles - split + nsplit - 1) / nsplit;   // tiles split, split+nsplit, ...

  if (threadIdx.x == 0) {
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int slot_ctr = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = split + t * nsplit;
        const int n_img = tile / p.tiles_per_image, tt = tile - n_img * p.tiles_per_image;
        const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
        const int y0 = ty * W_TH, x0 = tx * W_TW;
        for (int j = -1; j < ntap; ++j, ++slot_ctr) {     // j = -1: the S tile, then one G tile per tap
          const int slot = slot_ctr % W_STAGES, it = slot_ctr / W_STAGES;
          mbar_wait(&empty[slot], (it & 1) ^ 1);
          uint8_t* dst = smem + slot * W_PAIR_BYTES;
          mbar_arrive_expect_tx(&full[slot], (uint32_t)W_PAIR_BYTES);
          if (j < 0) {
            tma_load_4d(dst, &tmS, &full[slot], 0, x0, y0, n_img);
            tma_load_4d(dst + W_PAIR_BYTES / 2, &tmS, &full[slot], 64, x0, y0, n_img);
          } else {
            const int tap = tap0 + j, ky = tap / p.ksz, kx = tap - ky * p.ksz;
            const int cx = x0 * p.stride - p.pad + kx, cy = y0 * p.stride - p.pad + ky;
            tma_load_4d(dst, &tmG, &full[slot], 0, cx, cy, n_img);
            tma_load_4d(dst + W_PAIR_BYTES / 2, &tmG, &full[slot], 64, cx, cy, n_img);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_mn(128);
      int slot_ctr = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int s_slot = slot_ctr % W_STAGES, s_it = slot_ctr / W_STAGES;
        ++slot_ctr;
        mbar_wait(&full[s_slot], s_it & 1);
        const uint32_t a_base = smem_u32(smem + s_slot * W_PAIR_BYTES);
        for (int j = 0; j < ntap; ++j, ++slot_ctr) {
          const int slot = slot_ctr % W_STAGES, it = slot_ctr / W_STAGES;
          mbar_wait(&full[slot], it & 1);
          tcgen05_fence_after();
          const uint32_t b_base = smem_u32(smem + slot * W_PAIR_BYTES);
#pragma unroll
          for (int k = 0; k < 8; ++k) {                   // 8 x 16 pixels
            const uint64_t da = umma_desc_mn_sw128(a_base + k * 2048, W_PAIR_BYTES / 2);
            const uint64_t db = umma_desc_mn_sw128(b_base + k * 2048, W_PAIR_BYTES / 2);
            umma_bf16(tmem_base + (uint32_t)(j * 128), da, db, idesc, (t > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[slot]);
        }
        umma_commit(&empty[s_slot]);                      // S tile is free once its last tap has been issued
      }
      umma_commit(done);
    }
  } else {
    // epilogue: 4 warps, thread = output row m (TMEM lane), all 128 columns n of each tap
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(done, 0);
    tcgen05_fence_after();
    if (p.partial) {
      // two-stage form: this CTA's sums go to its own tile [tap][n][m] (lanes = consecutive m: every store instruction
      // is one 128-byte line); wgrad_reduce_kernel adds the splits in a fixed order and writes the final layout
      for (int j = 0; j < ntap; ++j) {
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
          uint32_t raw[32];
          if (my_tiles > 0) {
            tmem_ld32(t_row + (uint32_t)(j * 128 + h * 32), raw);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) raw[c] = 0u;
          }
          float* dst = p.partial + ((((size_t)split * p.ntaps + tap0 + j) * 128 + h * 32) * 128) + m;
#pragma unroll
          for (int c = 0; c < 32; ++c) dst[(size_t)c * 128] = __uint_as_float(raw[c]);
        }
      }
    } else if (my_tiles > 0) {
      // tap-major partial layout dW[tap][m][n].  Each thread stages its row (128 floats) in the now idle pipeline
      // buffers and adds it to global memory with ONE 512-byte bulk reduction (cp.reduce.async.bulk ... .add.f32:
      // four full lines per operation).  The first form issued 32 red.global.add.v4.f32 per thread and tap, a warp
      // instruction touching 32 different lines 16 bytes each - the split-K reduction, not the MMAs, was most of the
      // kernel.  Rows are 528 bytes apart in shared memory (16-byte stores of a warp then spread over all banks);
      // two tap buffers alternate.
      static_assert(2 * 128 * W_ROW_PAD <= W_STAGES * W_PAIR_BYTES, "staging fits the pipeline buffers");
      const bool bulk = p.bulk_reduce != 0;
      for (int j = 0; j < ntap; ++j) {
        uint8_t* row = smem + (j & 1) * (128 * W_ROW_PAD) + m * W_ROW_PAD;
        if (bulk && j >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // this buffer's previous row
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
          uint32_t raw[32];
          tmem_ld32(t_row + (uint32_t)(j * 128 + h * 32), raw);
          tmem_ld_wait();
          float* dst = p.dW + ((size_t)(tap0 + j) * 128 + m) * 128 + h * 32;
          if (bulk) {
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              *reinterpret_cast<uint4*>(row + (h * 32 + c) * 4) = make_uint4(raw[c], raw[c + 1], raw[c + 2], raw[c + 3]);
          } else {
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(__uint_as_float(raw[c])),
                           "f"(__uint_as_float(raw[c + 1])), "f"(__uint_as_float(raw[c + 2])),
                           "f"(__uint_as_float(raw[c + 3]))
                           : "memory");
          }
        }
        if (bulk) {
          fence_proxy_async_smem();      // this thread's generic-proxy writes of its own row -> visible to the bulk copy
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(
                           p.dW + ((size_t)(tap0 + j) * 128 + m) * 128),
                       "r"(smem_u32(row))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (bulk) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 get_encode2() {
  static EncodeTiledFn2 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn2)ptr;
  });
  return fn;
}

static int make_nhwc_map(CUtensorMap* m, const void* base, int B, int H, int W, int S) {
  EncodeTiledFn2 enc = get_encode2();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[4] = {128, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {256, (cuuint64_t)W * 256, (cuuint64_t)H * W * 256};
  cuuint32_t box[4] = {64, (cuuint32_t)(W_TW * S), (cuuint32_t)(W_TH * S), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)S, (cuuint32_t)S, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// 9x9 / stride-4 weight gradients (conv1: OIHW [128][3][9][9]; deconv3: IOHW [128][3][9][9]):
//   dW[m][k] += sum_pix S[pix][m] * im2col(image)[pix][k],  k = c*81 + ky*9 + kx  (243 -> 256)
// S: NHWC bf16 [B][h][w][128] (TMA, MN-major A operand); image: fp32 NCHW [B][3][4h][4w]; its im2col rows are
// built in shared memory from a TMA-staged patch exactly as the CONV1 forward does, and used as an MN-major B
// operand with N = 256 (four 64-wide k-block tiles).  Persistent CTAs, double-buffered S / patch.
constexpr int W9_THREADS = 64 + 256;
constexpr int W9_S_BYTES = 2 * IM_SLOT_BYTES;   // 32 KB
constexpr int W9_B_BYTES = 4 * IM_SLOT_BYTES;   // 64 KB

struct Wg9Params {
  int tiles_x, tiles_per_image, total_tiles;
  float* dW;        // [128][243]
  float* partial;   // two-stage reduction: [ctas][256 k][128 m] per-CTA sums (no atomics), or nullptr
};

__global__ void __launch_bounds__(W9_THREADS, 1)
wgrad9_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmI, const Wg9Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_buf = smem;                                  // [2][32 KB]
  uint8_t* p_buf = smem + 2 * W9_S_BYTES;                 // [2][32 KB] fp32 patches
  uint8_t* b_buf = smem + 2 * W9_S_BYTES + 2 * P1_SLOT;   // 64 KB im2col tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + W9_B_BYTES);
  uint64_t* s_full = bars;          // [2]
  uint64_t* p_full = bars + 2;      // [2]
  uint64_t* s_empty = bars + 4;     // [2] MMA done with S[b]
  uint64_t* p_empty = bars + 6;     // [2] builders done with patch[b]
  uint64_t* b_full = bars + 8;      // builders -> MMA
  uint64_t* b_empty = bars + 9;     // MMA -> builders
  uint64_t* done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 1); mbar_init(&s_empty[i], 1); mbar_init(&p_empty[i], 256);
    }
    mbar_init(b_full, 256); mbar_init(b_empty, 1); mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmI);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  pdl_trigger();
  pdl_wait();              // the prologue above overlapped the predecessor's tail (programmatic dependent launch)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = blockIdx.x + t * gridDim.x;
        const int n_img = tile / p.tiles_per_image, tt = tile - n_img * p.tiles_per_image;
        const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
        const int y0 = ty * W_TH, x0 = tx * W_TW;
        const int b = t & 1, it = t >> 1;
        mbar_wait(&s_empty[b], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&s_full[b], (uint32_t)W9_S_BYTES);
        tma_load_4d(s_buf + b * W9_S_BYTES, &tmS, &s_full[b], 0, x0, y0, n_img);
        tma_load_4d(s_buf + b * W9_S_BYTES + IM_SLOT_BYTES, &tmS, &s_full[b], 64, x0, y0, n_img);
        mbar_wait(&p_empty[b], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&p_full[b], (uint32_t)P1_BYTES);
        tma_load_4d(p_buf + b * P1_SLOT, &tmI, &p_full[b], x0 * 4 - 4, y0 * 4 - 4, 0, n_img);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_mn(256);
      for (int t = 0; t < my_tiles; ++t) {
        const int b = t & 1, it = t >> 1;
        mbar_wait(&s_full[b], it & 1);
        mbar_wait(b_full, t & 1);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(s_buf + b * W9_S_BYTES), b_base = smem_u32(b_buf);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = umma_desc_mn_sw128(a_base + k * 2048, IM_SLOT_BYTES);
          const uint64_t db = umma_desc_mn_sw128(b_base + k * 2048, IM_SLOT_BYTES);
          umma_bf16(tmem_base, da, db, idesc, (t > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&s_empty[b]);
        umma_commit(b_empty);
      }
      umma_commit(done);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ty = r / W_TW, tx = r - ty * W_TW;
    for (int t = 0; t < my_tiles; ++t) {
      const int b = t & 1, it = t >> 1;
      const float* pp = reinterpret_cast<const float*>(p_buf + b * P1_SLOT) + (4 * ty) * P1_W + 4 * tx;
      mbar_wait(&p_full[b], it & 1);
      mbar_wait(b_empty, (t & 1) ^ 1);
      if (half == 0) {
        conv1_build_row<0, 0, false>(pp, b_buf + 0 * IM_SLOT_BYTES, b_buf, r);
        conv1_build_row<1, 0, false>(pp, b_buf + 1 * IM_SLOT_BYTES, b_buf, r);
        conv1_build_row<2, 0, false>(pp, b_buf + 2 * IM_SLOT_BYTES, b_buf, r);
        conv1_build_row<3, 0, false>(pp, b_buf + 3 * IM_SLOT_BYTES, b_buf, r);
      } else {
        conv1_build_row<0, 1, false>(pp, b_buf + 0 * IM_SLOT_BYTES, b_buf, r);
        conv1_build_row<1, 1, false>(pp, b_buf + 1 * IM_SLOT_BYTES, b_buf, r);
        conv1_build_row<2, 1, false>(pp, b_buf + 2 * IM_SLOT_BYTES, b_buf, r);
        conv1_build_row<3, 1, false>(pp, b_buf + 3 * IM_SLOT_BYTES, b_buf, r);
      }
      fence_proxy_async_smem();
      mbar_arrive(b_full);
      mbar_arrive(&p_empty[b]);
    }
    mbar_wait(done, 0);
    tcgen05_fence_after();
    if (my_tiles > 0) {
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const int m = r;
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        uint32_t raw[32];
        const int k0 = half * 128 + h * 32;
        tmem_ld32(t_row + (uint32_t)k0, raw);
        tmem_ld_wait();
        if (p.partial) {     // k-major per-CTA tile: lanes = consecutive m -> every store instruction is one 128-byte line
          float* dst = p.partial + ((size_t)blockIdx.x * 256 + k0) * 128 + m;
#pragma unroll
          for (int c = 0; c < 32; ++c) dst[(size_t)c * 128] = __uint_as_float(raw[c]);
        } else {
          float* dst = p.dW + (size_t)m * 243 + k0;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (k0 + c < 243) atomicAdd(dst + c, __uint_as_float(raw[c]));
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int make_image_map9(CUtensorMap* m, const void* base, int B, int H, int W) {
  EncodeTiledFn2 enc = get_encode2();
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

// second stage of the 5x5 / 1x1 weight gradient: out[m][n][tap] += sum_split partial[split][tap][n][m], fixed order;
// one block per (tap, n), thread = m.  `out` is the state_dict layout [128][128][k][k].
__global__ void __launch_bounds__(128) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                           int nsplit, int ntaps) {
  const int tap = blockIdx.x, n = blockIdx.y, m = threadIdx.x;
  const size_t stride = (size_t)ntaps * 128 * 128;
  const float* src = partial + ((size_t)tap * 128 + n) * 128 + m;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int j = 0;
  for (; j + 4 <= nsplit; j += 4) {
    s0 += __ldg(src + (size_t)(j + 0) * stride);
    s1 += __ldg(src + (size_t)(j + 1) * stride);
    s2 += __ldg(src + (size_t)(j + 2) * stride);
    s3 += __ldg(src + (size_t)(j + 3) * stride);
  }
  for (; j < nsplit; ++j) s0 += __ldg(src + (size_t)j * stride);
  out[((size_t)m * 128 + n) * ntaps + tap] += (s0 + s1) + (s2 + s3);
}

// second stage of the 9x9 weight gradient: dW[m][k] += sum_j partial[j][k][m] in a fixed order (deterministic);
// one block per k, thread = (m, j mod 4).  Every thread's loads are independent and issued together (the first form
// walked the partial tiles four at a time: 16 dependent round trips to L2 = 10 us at the end of the training step).
__global__ void __launch_bounds__(512) wgrad9_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW,
                                                            int nparts) {
  __shared__ float red[4][128];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int k = blockIdx.x, m = threadIdx.x & 127, jg = threadIdx.x >> 7;
  const float* src = partial + (size_t)k * 128 + m;
  float s = 0.f;
  for (int j0 = jg; j0 < nparts; j0 += 64) {          // 16 loads in flight per thread, summed in index order
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (j0 + 4 * i < nparts) ? __ldg(src + (size_t)(j0 + 4 * i) * 256 * 128) : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
  }
  red[jg][m] = s;
  __syncthreads();
  if (jg == 0) dW[(size_t)m * 243 + k] += (red[0][m] + red[1][m]) + (red[2][m] + red[3][m]);
}

// x^2 of the GDN input from the saved (y, n): GDN x^2 = y^2 n, IGDN x^2 = y^2 / n   (bf16, 8 per thread)
__global__ void __launch_bounds__(256) xsq_kernel(const uint4* __restrict__ y, const uint4* __restrict__ n,
                                                  uint4* __restrict__ out, size_t nvec, int inverse) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 qy = __ldg(y + i), qn = __ldg(n + i);
    const uint32_t wy[4] = {qy.x, qy.y, qy.z, qy.w}, wn[4] = {qn.x, qn.y, qn.z, qn.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float y0 = __uint_as_float(wy[k] << 16), y1 = __uint_as_float(wy[k] & 0xffff0000u);
      const float n0 = __uint_as_float(wn[k] << 16), n1 = __uint_as_float(wn[k] & 0xffff0000u);
      const float r0 = inverse ? y0 * y0 / n0 : y0 * y0 * n0;
      const float r1 = inverse ? y1 * y1 / n1 : y1 * y1 * n1;
      o[k] = pack_bf16x2(__float2bfloat16_rn(r0), __float2bfloat16_rn(r1));
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[c] += sum over pixel rows of a bf16 [P][128] tensor (bias / beta gradients); out zero-filled by the caller.
// thread = (8-channel group, row lane): 16-byte loads, all of a thread's rows requested before the first add (the
// first form read 2 bytes per thread and row, 128 dependent iterations: 12 us for 8 MB, ten launches per step).
constexpr int CS_ROWS = 128;       // rows per block
__global__ void __launch_bounds__(256) nhwc_colsum_kernel(const __nv_bfloat16* __restrict__ t, float* __restrict__ out,
                                                          size_t P, int rows_per_block) {
  __shared__ float red[16][129];
  const int cg = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const size_t r0 = (size_t)blockIdx.x * rows_per_block;
  const size_t r1 = min(P, r0 + (size_t)rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (size_t rb = r0 + rl; rb < r1; rb += 16 * 8) {
    uint4 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const size_t r = rb + (size_t)i * 16;
      q[i] = r < r1 ? __ldg(reinterpret_cast<const uint4*>(t + r * 128) + cg) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t w[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[2 * k] += __uint_as_float(w[k] << 16);
        s[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cg * 8 + k] = s[k];
  __syncthreads();
  if (threadIdx.x < 128) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a += red[i][threadIdx.x];
    atomicAdd(out + threadIdx.x, a);
  }
}

// tap-major weight-gradient accumulator [taps][128 m][128 n] -> the parameter's layout [128 m][128 n][taps]:
// one block per m; 25 coalesced 512-byte row reads into shared memory, one contiguous 12.8 KB write
// (torch's generic strided copy took 6-18 us per layer inside the step)
__global__ void __launch_bounds__(256) wgrad_untap_kernel(const float* __restrict__ tmp, float* __restrict__ out,
                                                          int ntaps, int accumulate) {
  extern __shared__ float s_t[];      // [ntaps][129]
  const int m = blockIdx.x;
  for (int i = threadIdx.x; i < ntaps * 128; i += blockDim.x) {
    const int t = i >> 7, n = i & 127;
    s_t[t * 129 + n] = __ldg(tmp + ((size_t)t * 128 + m) * 128 + n);
  }
  __syncthreads();
  float* dst = out + (size_t)m * 128 * ntaps;
  for (int i = threadIdx.x; i < ntaps * 128; i += blockDim.x) {
    const int n = i / ntaps, t = i - n * ntaps;
    const float v = s_t[t * 129 + n];
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

}  // namespace tc
}  // namespace b2c

using namespace b2c;
using namespace b2c::tc;

extern "C" {

static int wgrad9_ctas(int total_tiles) {
  static const int tiles_per_cta = [] { const char* e = getenv("B2C_WG9_TILES_PER_CTA"); int v = e ? atoi(e) : 4; return v > 0 ? v : 4; }();
  int ctas = (total_tiles + tiles_per_cta - 1) / tiles_per_cta;
  if (ctas > 148) ctas = 148;
  if (ctas < 1) ctas = 1;
  return ctas;
}

size_t b2c_tc_wgrad9_workspace_bytes(int B, int h, int w) {
  if (B <= 0 || h <= 0 || w <= 0) return 0;
  const int total_tiles = B * ((w + W_TW - 1) / W_TW) * ((h + W_TH - 1) / W_TH);
  return (size_t)wgrad9_ctas(total_tiles) * 256 * 128 * sizeof(float);
}

int b2c_tc_wgrad9(const void* small_t, const float* image, float* dW, int B, int h, int w, void* workspace,
                  size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(small_t && image && dW && B > 0 && h > 0 && w > 0, "tc_wgrad9: bad arguments");
  constexpr size_t SMEM = 2 * W9_S_BYTES + 2 * P1_SLOT + W9_B_BYTES + 1024 + 256;
  {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(wgrad9_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
  }
  Wg9Params p;
  p.tiles_x = (w + W_TW - 1) / W_TW;
  p.tiles_per_image = p.tiles_x * ((h + W_TH - 1) / W_TH);
  p.total_tiles = B * p.tiles_per_image;
  p.dW = dW;
  CUtensorMap mS, mI;
  int rc;
  if ((rc = make_nhwc_map(&mS, small_t, B, h, w, 1))) return rc;
  if ((rc = make_image_map9(&mI, image, B, 4 * h, 4 * w))) return rc;
  int ctas;
  if (workspace) {
    // two-stage, atomic-free: per-CTA partial tiles, then a fixed-order reduction (deterministic)
    ctas = wgrad9_ctas(p.total_tiles);
    if (workspace_bytes < (size_t)ctas * 256 * 128 * sizeof(float)) {
      set_error("tc_wgrad9: workspace too small");
      return B2C_EWORKSPACE;
    }
    p.partial = (float*)workspace;
  } else {
    ctas = (p.total_tiles + 3) / 4;   // >= 4 tiles per CTA: the scalar red.global.add of the 128 x 243 tile dominate
    if (ctas > 148) ctas = 148;
    p.partial = nullptr;
  }
  {
    cudaError_t e = launch_pdl(PDL_WGRAD9, wgrad9_tc_kernel, dim3(ctas), dim3(W9_THREADS), SMEM, (cudaStream_t)stream, mS, mI, p);
    if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(wgrad9): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  }
  count_launch();
  if (workspace) {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(243);
    cfg.blockDim = dim3(512);
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_small_enabled(PDL_WGRAD9) ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad9_reduce_kernel, (const float*)workspace, dW, ctas);
    if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(wgrad9_reduce): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
    count_launch();
  }
  return check_launch("tc_wgrad9");
}

int b2c_tc_wgrad_untap(const float* tmp, float* out, int ntaps, int accumulate, void* stream) {
  B2C_REQUIRE(tmp && out && ntaps > 0 && ntaps <= 81, "tc_wgrad_untap: bad arguments");
  wgrad_untap_kernel<<<128, 256, (size_t)ntaps * 129 * sizeof(float), (cudaStream_t)stream>>>(tmp, out, ntaps, accumulate);
  count_launch();
  return check_launch("tc_wgrad_untap");
}

int b2c_tc_xsq(const void* y, const void* n, void* out, size_t count, int inverse, void* stream) {
  B2C_REQUIRE(y && n && out && count > 0 && count % 8 == 0, "tc_xsq: bad arguments");
  const size_t nvec = count / 8;
  size_t blocks = (nvec + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  xsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)y, (const uint4*)n, (uint4*)out, nvec,
                                                                 inverse);
  count_launch();
  return check_launch("tc_xsq");
}

int b2c_tc_nhwc_colsum(const void* t, float* out, size_t P, void* stream) {
  B2C_REQUIRE(t && out && P > 0, "tc_nhwc_colsum: bad arguments");
  const int rows_per_block = CS_ROWS;
  const size_t blocks = (P + rows_per_block - 1) / rows_per_block;
  nhwc_colsum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)t, out, P,
                                                                        rows_per_block);
  count_launch();
  return check_launch("tc_nhwc_colsum");
}

static int wgrad_splits(int total_tiles, int ntaps) {
  const int groups = (ntaps + W_TAPS - 1) / W_TAPS;
  static const int forced = [] { const char* e = getenv("B2C_WG_SPLITS"); return e ? atoi(e) : 0; }();
  if (forced > 0) return forced < total_tiles ? forced : total_tiles;
  int splits = 148 / groups;
  // >= 4 pixel tiles per CTA: fewer reductions (B2C_WG_MIN_TILES: for the 25-tap layers)
  static const int min_tiles5 = [] { const char* e = getenv("B2C_WG_MIN_TILES"); const int v = e ? atoi(e) : 4; return v > 0 ? v : 4; }();
  static const int min_tiles1 = [] { const char* e = getenv("B2C_WG_MIN_TILES1"); const int v = e ? atoi(e) : 4; return v > 0 ? v : 4; }();
  const int mt = ntaps > 1 ? min_tiles5 : min_tiles1;
  if (splits > (total_tiles + mt - 1) / mt) splits = (total_tiles + mt - 1) / mt;
  if (splits < 1) splits = 1;
  return splits;
}

size_t b2c_tc_wgrad_workspace_bytes(int B, int hs, int ws, int k) {
  if (B <= 0 || hs <= 0 || ws <= 0 || (k != 1 && k != 5)) return 0;
  const int total_tiles = B * ((ws + W_TW - 1) / W_TW) * ((hs + W_TH - 1) / W_TH);
  return (size_t)wgrad_splits(total_tiles, k * k) * k * k * 128 * 128 * sizeof(float);
}

int b2c_tc_wgrad(const void* small_t, const void* big_t, float* dW, int B, int hs, int ws, int hb, int wb, int k,
                 int stride, int pad, void* workspace, size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(small_t && big_t && dW, "tc_wgrad: null pointer");
  B2C_REQUIRE(B > 0 && hs > 0 && ws > 0 && hb > 0 && wb > 0 && (k == 1 || k == 5) && (stride == 1 || stride == 2) &&
                  pad >= 0,
              "tc_wgrad: unsupported geometry (k in {1,5}, stride in {1,2})");
  constexpr size_t SMEM = (size_t)W_STAGES * W_PAIR_BYTES + 1024 + 256;
  {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
  }
  WgParams p;
  p.B = B; p.hs = hs; p.ws = ws;
  p.tiles_x = (ws + W_TW - 1) / W_TW;
  p.tiles_per_image = p.tiles_x * ((hs + W_TH - 1) / W_TH);
  p.total_tiles = B * p.tiles_per_image;
  p.ksz = k; p.stride = stride; p.pad = pad; p.ntaps = k * k;
  p.dW = dW;
  p.partial = nullptr;
  static const int bulk = [] { const char* e = getenv("B2C_WG_BULK"); return e ? atoi(e) : 1; }();
  p.bulk_reduce = bulk;
  CUtensorMap mS, mG;
  int rc;
  if ((rc = make_nhwc_map(&mS, small_t, B, hs, ws, 1))) return rc;
  if ((rc = make_nhwc_map(&mG, big_t, B, hb, wb, stride))) return rc;
  const int groups = (p.ntaps + W_TAPS - 1) / W_TAPS;
  const int splits = wgrad_splits(p.total_tiles, p.ntaps);
  if (workspace) {
    if (workspace_bytes < (size_t)splits * p.ntaps * 128 * 128 * sizeof(float)) {
      set_error("tc_wgrad: workspace too small");
      return B2C_EWORKSPACE;
    }
    p.partial = (float*)workspace;
  }
  dim3 grid(groups, splits);
  wgrad_tc_kernel<<<grid, W_THREADS, SMEM, (cudaStream_t)stream>>>(mS, mG, p);
  count_launch();
  if (workspace) {
    wgrad_reduce_kernel<<<dim3(p.ntaps, 128), 128, 0, (cudaStream_t)stream>>>((const float*)workspace, dW, splits, p.ntaps);
    count_launch();
  }
  return check_launch("tc_wgrad");
}

}  // extern "C"
