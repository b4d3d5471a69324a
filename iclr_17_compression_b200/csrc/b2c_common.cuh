// Shared helpers for libb2c: error reporting, launch accounting and the exact-fp32
// 128x128x8 register-tiled GEMM core that the fp32 convolution / GDN kernels build on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2c.h"

namespace b2c {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError -> B2C_ECUDA

#define B2C_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      b2c::set_error(__VA_ARGS__);             \
      return B2C_EINVAL;                       \
    }                                          \
  } while (0)

// ---------------------------------------------------------------------------------
// fp32 tile GEMM:  C[m][n] = sum_k A[m][k] * B[k][n]
// CTA tile 128(m) x 128(n), k step 8, 256 threads, 8x8 accumulators per thread.
// Operands are produced by loader functors so the same core serves im2col-gathered
// convolutions, sub-pixel transposed convolutions, weight gradients and the GDN
// channel-mixing product.  Loader contract (each thread owns ONE row/column of its
// operand tile and 4 consecutive k per step):
//     void init(int idx_global, int k_first);   // idx = m (A) or n (B); k_first = first k of this thread
//     void fetch4(float v[4]);                   // values at k, k+1, k+2, k+3 ; then k += 8
// Out-of-range idx / k must yield 0.
// ---------------------------------------------------------------------------------
constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 8, TG_THREADS = 256;

struct TileAcc {
  float v[8][8];  // [i: m sub-index][j: n sub-index]
  // thread's m indices: m0 + i (i<4), m0 + 64 + (i-4) (i>=4) with m0 = tile_m + tx*4
  // thread's n indices: n0 + j (j<4), n0 + 64 + (j-4) (j>=4) with n0 = tile_n + ty*4
  int m0, n0;
  __device__ __forceinline__ int m(int i) const { return m0 + (i & 3) + ((i >> 2) << 6); }
  __device__ __forceinline__ int n(int j) const { return n0 + (j & 3) + ((j >> 2) << 6); }
};

template <class ALoader, class BLoader>
__device__ __forceinline__ void tile_gemm(ALoader& la, BLoader& lb, int tile_m, int tile_n,
                                          int k_begin, int k_end, TileAcc& acc) {
  __shared__ __align__(16) float As[2][TG_BK][TG_BM];
  __shared__ __align__(16) float Bs[2][TG_BK][TG_BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid & 127, lk = (tid >> 7) * 4;
  acc.m0 = tile_m + tx * 4;
  acc.n0 = tile_n + ty * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[i][j] = 0.f;

  la.init(tile_m + lrow, k_begin + lk);
  lb.init(tile_n + lrow, k_begin + lk);
  const int nsteps = (k_end - k_begin + TG_BK - 1) / TG_BK;
  float ra[4], rb[4];
  if (nsteps > 0) {
    la.fetch4(ra);
    lb.fetch4(rb);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[0][lk + j][lrow] = ra[j];
      Bs[0][lk + j][lrow] = rb[j];
    }
  }
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) {
      la.fetch4(ra);
      lb.fetch4(rb);
    }
#pragma unroll
    for (int kk = 0; kk < TG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[i][j] = fmaf(a[i], b[j], acc.v[i][j]);
    }
    if (s + 1 < nsteps) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[buf ^ 1][lk + j][lrow] = ra[j];
        Bs[buf ^ 1][lk + j][lrow] = rb[j];
      }
    }
    __syncthreads();
  }
}

// block-wide sum of one float per thread (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ float block_sum(float v, float* smem32) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem32[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

}  // namespace b2c

// ---------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL) for the small kernels of a launch-latency-bound chain (the training step):
// the kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization, so its blocks may be scheduled
// while the predecessor in the stream drains; it must execute pdl_wait() before touching global memory and may
// call pdl_trigger() to let ITS successor do the same.  B2C_PDL=0 restores plain stream order.
#include <cstdlib>
#ifndef B2C_PDL_SMALL_DEFAULT
#define B2C_PDL_SMALL_DEFAULT 2
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

namespace b2c {
// B2C_PDL_SMALL: bit mask of the small-kernel groups launched as programmatic dependents
//   1 = quantizer/entropy forward, 2 = NCHW->NHWC transposes, 4 = 9x9 weight gradient.
// Measured on the batch-8 training step: none 357 us, entropy 360, layout 350, wgrad9 376 (its early CTAs hold
// shared memory / TMEM next to the last dgrad kernel), all 382 -> default = layout only.
enum { PDL_ENTROPY = 1, PDL_LAYOUT = 2, PDL_WGRAD9 = 4 };
static inline bool pdl_small_enabled(int group) {
  static const int mask = [] {
    const char* e = getenv("B2C_PDL");
    if (e && atoi(e) == 0) return 0;
    const char* f = getenv("B2C_PDL_SMALL");
    return f ? atoi(f) : B2C_PDL_SMALL_DEFAULT;
  }();
  return (mask & group) != 0;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(int group, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t stream, Args... args) {
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_small_enabled(group) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
}  // namespace b2c
