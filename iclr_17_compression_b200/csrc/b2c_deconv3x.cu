// DECONV3X: the last synthesis layer (9x9 s4 p4 op3 transposed conv 128 -> 3, models/synthesis_17.py:23,30) with the
// HORIZONTAL input shifts folded into the GEMM's N dimension.
//
// The 3x3-shift form of b2c_conv_tcp.cu (kind DECONV3) issues, per 8x16 tile, 9 shifts x 8 K-slices = 72 MMAs of
// N = 48 (16 sub-pixel phases x 3 channels).  An SS-mode MMA reads its 4 KB A slice from shared memory whatever N is,
// so those MMAs ran at 70-80 clk instead of 24 and the layer sat at 30 % tensor-pipe utilisation, bound by
// shared-memory bandwidth (profiles/r2_notes.md).  Here every A slice is used by ONE MMA of N = 112:
//
//   E[(m, n'), (dx, ry, rx, o)] = sum_{dy, i} in[i, m + dy, n'] * W[i, o, ry + 4 - 4 dy, rx + 4 - 4 dx]
//   out[o, 4m + ry, 4n + rx]    = bias[o] + sum_{dx in {-1, 0, +1}} E[(m, n + dx), (dx, ry, rx, o)]
//
// dx = 0 and dx = +1 use all four rx (kx = rx + 4, rx), dx = -1 only rx = 0 (kx = 8): 48 + 48 + 12 = 108 -> 112 columns.
// The three VERTICAL shifts stay MMA accumulations over one TMA box of TH + 2 tile rows (the A operand of tap dy is the
// box advanced by whole 2 KB tile rows, as in b2c_conv_tcp.cu), so a tile is 3 taps x 2 k-blocks x 4 K-slices = 24
// MMAs and two A boxes (40 KB instead of 120 KB of fill).  The horizontal sum is done by the epilogue: a tile row is
// 16 consecutive lanes of a warp, so the right / left neighbours' partial sums arrive by warp shuffle.  The two edge
// columns of a tile have no neighbour inside the warp: tiles therefore advance by TW - 2 = 14 columns and the edge
// lanes only supply partial sums (87.5 % of the MMA rows produce output).
//
// One persistent CTA per SM, resident weights (84 KB), a 4-deep ring of A boxes, two TMEM accumulator sets so the
// epilogue of tile i (shuffles, bias, clamp, squared errors, fp32 NCHW stores) overlaps the MMAs of tile i + 1.
// With the fused distortion the reference image is 30 % of the layer's HBM traffic: its 3 x 32 x 64-pixel window of
// every tile comes through a TMA ring of its own, two tiles ahead (per-thread global loads issued at the top of a
// tile's epilogue left one tile of latency hiding: 4.6 k clk per tile in the epilogue against 1.7 k of MMAs).
#include <mutex>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_host.cuh"
#include "b2c_tcp_util.cuh"
#include "b2c_deconv3x_pack.cuh"

namespace b2c {
namespace tcp3x {

using namespace b2c::tc;

constexpr int NWORKERS = 512;                       // 16 epilogue warps: (TMEM lane quarter q, sub-pixel row ry)
constexpr int NWARPS_W = NWORKERS / 32;
constexpr int NTHREADS = 64 + NWORKERS;
constexpr int BOX_ROWS = TH + 2;
constexpr int ROW_BYTES = TW * 128;                 // one tile row of a k-block: 2 KB
constexpr int A_SLOT = BOX_ROWS * ROW_BYTES;        // 20 KB
constexpr int NX = X_NX;                            // GEMM N: 4 sub-pixel rows x (12 dx=0, 12 dx=+1, 3 dx=-1, 1 zero)
constexpr int PART_COLS = X_PART_COLS;              // 28 accumulator columns per sub-pixel row ry
constexpr int B_BYTES = NX * 128;                   // one (tap, k-block) weight tile: 14 KB
constexpr int TWX = TW - 2;                         // output block columns per tile
constexpr int R_W = 4 * TW, R_H = 4 * TH;           // 64 x 32 pixels x 3 channels of fp32
constexpr int R_SLOT = 3 * R_H * R_W * 4;           // 24 KB
constexpr int EXTRA = 1024 /*align*/ + 512 /*bias, reduction scratch*/ + 512 /*barriers*/;
// channels C in {128, 192, 256} (model.py:39; the wider-latent models of BASELINE config 5): C / 64 k-blocks, resident
// weights 84 / 126 / 168 KB.  What is left of the shared memory goes to the A ring and - except at 256 channels,
// where it does not fit - the reference-image ring (there the epilogue fetches its reference pixels itself).
template <int C>
struct XCfg {
  static constexpr int NCB = C / 64;
  static constexpr int W_RES = 3 * NCB * B_BYTES;
  static constexpr int NR = C <= 192 ? 2 : 0;                 // reference-image ring slots
  static constexpr int NA = C == 128 ? 4 : 2;                 // A-box ring slots
  static constexpr int SMEM = NA * A_SLOT + W_RES + NR * R_SLOT + EXTRA;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};
constexpr uint32_t ACC_STRIDE = 128;
constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;

struct XParams {
  int B, Hg, Wg;            // input grid = output block grid
  int tiles_x, tiles_per_image, n_tiles;
  int Hfull, Wfull;         // 4 Hg, 4 Wg
  const float* bias;        // [3]
  float* out_f32;           // [B][3][Hfull][Wfull]
  const float* ref_image;   // fused distortion (model.py:59,61; train.py:172) or nullptr
  float* sse_partial;       // [B][tiles_per_image][16 worker warps][2]
  float* grad_out;          // training: grad_scale * (recon - x) or nullptr
  float grad_scale;
  float* dbias;             // training: dbias[o] += sum of grad_out over (n, y, x) - the bias gradient - or nullptr
  long long* trace;         // debug: clock64 timeline of CTA 0: [3 roles][64 marks]
};

template <int C>
__global__ void __launch_bounds__(NTHREADS, 1)
deconv3x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmR, const XParams p) {
  using Cfg = XCfg<C>;
  constexpr int NCB = Cfg::NCB, NA = Cfg::NA, NR = Cfg::NR, W_RES = Cfg::W_RES;
  constexpr int NRX = NR > 0 ? NR : 1;
  constexpr bool REF_RING = NR > 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wres = smem + NA * A_SLOT;                 // [tap 3][kb 2] tiles of 112 rows x 128 B
  uint8_t* rring = wres + W_RES;                      // [NR] reference-image windows
  float* s_bias = reinterpret_cast<float*>(rring + NR * R_SLOT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 128);
  uint64_t* a_full = bars;                 // [NA]
  uint64_t* a_empty = a_full + NA;         // [NA]
  uint64_t* accum_full = a_empty + NA;     // [2]
  uint64_t* acc_empty = accum_full + 2;    // [2]  one arrival per worker warp
  uint64_t* res_full = acc_empty + 2;
  uint64_t* r_full = res_full + 1;         // [NR]
  uint64_t* r_empty = r_full + NRX;        // [NR] one arrival per worker warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + NRX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit0 = (int)blockIdx.x, unit_stride = (int)gridDim.x;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  int tmark = 0;
#define X_MARK(role, cond)                                                           \
  do {                                                                               \
    if (tracing && (cond) && tmark < 64) p.trace[(role) * 64 + tmark++] = clock64(); \
  } while (0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accum_full[i], 1); mbar_init(&acc_empty[i], NWARPS_W); }
    for (int i = 0; i < NR; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], NWARPS_W); }
    mbar_init(res_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (REF_RING && p.ref_image) tma_prefetch_desc(&tmR);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      griddep_wait();     // the packed weights and the previous layer's output are complete and visible from here on
      mbar_arrive_expect_tx(res_full, (uint32_t)W_RES);
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb)
          tma_load_2d(wres + (t * NCB + cb) * B_BYTES, &tmB, res_full, cb * KBLK, t * NX);
      uint32_t ga = 0, gr = 0;
      for (int tile = unit0; tile < p.n_tiles; tile += unit_stride, ++gr) {
        const int n_img = tile / p.tiles_per_image;
        const int trem = tile - n_img * p.tiles_per_image;
        const int tile_y = trem / p.tiles_x, tile_x = trem - tile_y * p.tiles_x;
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb, ++ga) {
          const uint32_t as = ga % NA;
          mbar_wait(&a_empty[as], ((ga / NA) & 1) ^ 1);
          X_MARK(0, true);
          mbar_arrive_expect_tx(&a_full[as], (uint32_t)A_SLOT);
          tma_load_4d(smem + as * A_SLOT, &tmA, &a_full[as], cb * KBLK, tile_x * TWX - 1, tile_y * TH - 1, n_img);
        }
        if (REF_RING && p.ref_image) {   // the window of the reference image under this tile (out-of-image parts zero-filled)
          const uint32_t rs = gr % NRX;
          mbar_wait(&r_empty[rs], ((gr / NRX) & 1) ^ 1);
          mbar_arrive_expect_tx(&r_full[rs], (uint32_t)R_SLOT);
          tma_load_4d(rring + rs * R_SLOT, &tmR, &r_full[rs], (tile_x * TWX - 1) * 4, tile_y * R_H, 0, n_img);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (whole warp, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_bf16(NX);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint64_t desc0 = umma_desc_sw128(sbase);
    const uint64_t wres_d = desc0 + ((uint32_t)(NA * A_SLOT) >> 4);
    uint32_t ga = 0, st = 0;
    mbar_wait(res_full, 0);
    for (int tile = unit0; tile < p.n_tiles; tile += unit_stride, ++st) {
      const uint32_t buf = st & 1u;
      mbar_wait(&acc_empty[buf], ((st >> 1) & 1u) ^ 1u);     // the workers have drained this accumulator set
      tcgen05_fence_after();
      const uint32_t d = tb + buf * ACC_STRIDE;
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb, ++ga) {
        const uint32_t as = ga % NA;
        mbar_wait(&a_full[as], (ga / NA) & 1);
        tcgen05_fence_after();
        X_MARK(1, lane == 0);
        const uint64_t a_box = desc0 + ((as * (uint32_t)A_SLOT) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 3; ++t) {               // vertical taps: dy = t - 1, A = the box advanced by t tile rows
            const uint64_t a = a_box + ((uint32_t)(t * ROW_BYTES) >> 4);
            const uint64_t b = wres_d + ((uint32_t)((t * NCB + cb) * B_BYTES) >> 4);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16(d, a + 2 * k4, b + 2 * k4, idesc, (cb > 0 || t > 0 || k4 > 0) ? 1u : 0u);
          }
          umma_commit(&a_empty[as]);
          if (cb == NCB - 1) umma_commit(&accum_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================ row workers: thread = (pixel r, sub-pixel row ry)
    // ry selects 28 accumulator columns; lanes 0..15 / 16..31 of a warp are the 16 columns of two tile rows, so a
    // column's horizontal neighbours are lane - 1 and lane + 1.  (Sixteen warps: with eight - two sub-pixel rows per
    // thread - the epilogue took 2 k clk per tile against 1.7 k clk of MMAs and bounded the kernel.)
    const int q = warp & 3;
    const int ry = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ty = r / TW, tx = r - ty * TW;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ry * PART_COLS);
    griddep_wait();                    // before any global read / write (the bias may come from an optimizer step)
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 3) s_bias[threadIdx.x - 64] = p.bias ? __ldg(p.bias + threadIdx.x - 64) : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(NWORKERS) : "memory");
    uint32_t st = 0;
    float sb[3] = {0.f, 0.f, 0.f};     // running sums of (recon - x) per channel over this thread's tiles (bias gradient)
    for (int tile = unit0; tile < p.n_tiles; tile += unit_stride, ++st) {
      const int n_img = tile / p.tiles_per_image;
      const int trem = tile - n_img * p.tiles_per_image;
      const int tile_y = trem / p.tiles_x, tile_x = trem - tile_y * p.tiles_x;
      const int oy = tile_y * TH + ty, ox = tile_x * TWX - 1 + tx;       // output block (= input pixel) of this thread
      const bool valid = tx >= 1 && tx <= TWX && oy < p.Hg && ox < p.Wg;
      const uint32_t buf = st & 1u, par = (st >> 1) & 1u;

      // without the reference ring (256 channels): requested before waiting for the accumulator
      float4 xpre[3];
#pragma unroll
      for (int o = 0; o < 3; ++o) xpre[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!REF_RING && p.ref_image && valid) {
#pragma unroll
        for (int o = 0; o < 3; ++o)
          xpre[o] = __ldg(reinterpret_cast<const float4*>(
              p.ref_image + (((size_t)n_img * 3 + o) * p.Hfull + (oy * 4 + ry)) * p.Wfull + ox * 4));
      }
      X_MARK(2, threadIdx.x == 64);
      mbar_wait(&accum_full[buf], par);
      tcgen05_fence_after();
      X_MARK(2, threadIdx.x == 64);
      // columns: [0,12) dx = 0 (own), [12,24) dx = +1 (for the left neighbour), [24,27) dx = -1 (for the right one);
      // the 32-column load also covers the first columns of the next sub-pixel row (unused; < 128 allocated columns)
      uint32_t v[32];
      const uint32_t t_acc = t_lane + buf * ACC_STRIDE;
      tmem_ld32(t_acc, v);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);        // this warp's TMEM reads of the set are complete
      float acc[12];                                      // [rx][o]
#pragma unroll
      for (int j = 0; j < 12; ++j)
        acc[j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[12 + j]), 1);
#pragma unroll
      for (int o = 0; o < 3; ++o)                         // dx = -1 reaches rx = 0 only
        acc[o] += __shfl_up_sync(0xffffffffu, __uint_as_float(v[24 + o]), 1);

      // fused distortion: this thread's 3 channels x 4 pixels of the reference window (TMA ring)
      float4 xref[3];
      if (!REF_RING) {
#pragma unroll
        for (int o = 0; o < 3; ++o) xref[o] = xpre[o];
      } else if (p.ref_image) {
        const uint32_t rs = st % NRX;
        mbar_wait(&r_full[rs], (st / NRX) & 1u);
        const float* rw = reinterpret_cast<const float*>(rring + rs * R_SLOT) + (ty * 4 + ry) * R_W + tx * 4;
#pragma unroll
        for (int o = 0; o < 3; ++o) xref[o] = *reinterpret_cast<const float4*>(rw + o * R_H * R_W);
        __syncwarp();
        if (lane == 0) mbar_arrive(&r_empty[rs]);
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) xref[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }

      float s0 = 0.f, s1 = 0.f;
      if (valid) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const float bv = s_bias[o];
          {
            float4 v4 = make_float4(acc[0 * 3 + o] + bv, acc[1 * 3 + o] + bv, acc[2 * 3 + o] + bv, acc[3 * 3 + o] + bv);
            const size_t off = (((size_t)n_img * 3 + o) * p.Hfull + (oy * 4 + ry)) * p.Wfull + ox * 4;
            if (p.ref_image) {
              const float4 x4 = xref[o];
              const float4 c4 = make_float4(fminf(fmaxf(v4.x, 0.f), 1.f), fminf(fmaxf(v4.y, 0.f), 1.f),
                                            fminf(fmaxf(v4.z, 0.f), 1.f), fminf(fmaxf(v4.w, 0.f), 1.f));
              float dd;
              dd = v4.x - x4.x; s0 = fmaf(dd, dd, s0); dd = v4.y - x4.y; s0 = fmaf(dd, dd, s0);
              dd = v4.z - x4.z; s0 = fmaf(dd, dd, s0); dd = v4.w - x4.w; s0 = fmaf(dd, dd, s0);
              dd = c4.x - x4.x; s1 = fmaf(dd, dd, s1); dd = c4.y - x4.y; s1 = fmaf(dd, dd, s1);
              dd = c4.z - x4.z; s1 = fmaf(dd, dd, s1); dd = c4.w - x4.w; s1 = fmaf(dd, dd, s1);
              if (p.grad_out) {   // d/d recon of (grad_scale / 2) * sum (recon - x)^2: the distortion backward, fused
                *reinterpret_cast<float4*>(p.grad_out + off) =
                    make_float4(p.grad_scale * (v4.x - x4.x), p.grad_scale * (v4.y - x4.y),
                                p.grad_scale * (v4.z - x4.z), p.grad_scale * (v4.w - x4.w));
                sb[o] += ((v4.x - x4.x) + (v4.y - x4.y)) + ((v4.z - x4.z) + (v4.w - x4.w));
              }
              v4 = c4;
            }
            *reinterpret_cast<float4*>(p.out_f32 + off) = v4;
          }
        }
      }
      if (p.ref_image) {   // fixed-order reduction: butterfly inside the warp, one partial-sum slot per (tile, warp) -
#pragma unroll         // the worker warps never meet on a block barrier
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0)
          *reinterpret_cast<float2*>(p.sse_partial +
                                     (((size_t)n_img * p.tiles_per_image + trem) * NWARPS_W + (warp - 2)) * 2) =
              make_float2(s0, s1);
      }
      X_MARK(2, threadIdx.x == 64);
    }
    if (p.dbias && p.grad_out) {     // bias gradient = sum of grad_out: per-CTA sum, three atomic adds per CTA
#pragma unroll
      for (int o = 0; o < 3; ++o) {
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) sb[o] += __shfl_xor_sync(0xffffffffu, sb[o], sh);
      }
      float* red = s_bias + 16;      // [16 warps][3]
      if (lane == 0) { red[(warp - 2) * 3] = sb[0]; red[(warp - 2) * 3 + 1] = sb[1]; red[(warp - 2) * 3 + 2] = sb[2]; }
      asm volatile("bar.sync 1, %0;" ::"n"(NWORKERS) : "memory");
      if (warp == 2 && lane < 3) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < NWARPS_W; ++w) a += red[w * 3 + lane];
        atomicAdd(p.dbias + lane, p.grad_scale * a);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
#undef X_MARK
}

__global__ void pack_deconv3x_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int total, int C) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  out[idx] = __float2bfloat16_rn(deconv3x_weight(w, idx, C));
}

size_t packed_elems(int channels) { return (size_t)3 * NX * channels; }

int pack(const float* w, void* out, int channels, cudaStream_t stream) {
  const int n = (int)packed_elems(channels);
  pack_deconv3x_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, (__nv_bfloat16*)out, n, channels);
  count_launch();
  return check_launch("tc_pack_weights(deconv3x)");
}

// partial-sum slots per image: one per (tile, worker warp)
int tiles_per_image(int Hin, int Win) { return ((Win + TWX - 1) / TWX) * ((Hin + TH - 1) / TH) * NWARPS_W; }

static int make_box_map(CUtensorMap* m, const void* base, int B, int H, int W, int C) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)KBLK, (cuuint32_t)TW, (cuuint32_t)BOX_ROWS, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(deconv3x box) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

// fp32 NCHW reference image [B][3][H][W]: box = 64 x 32 pixels x 3 channels (the output window of one tile)
static int make_ref_map(CUtensorMap* m, const void* base, int B, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  cuuint32_t box[4] = {(cuuint32_t)R_W, (cuuint32_t)R_H, 3, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(reference image) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

template <int C>
static int launch_c(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  constexpr int SMEM = XCfg<C>::SMEM;
  XParams p;
  memset(&p, 0, sizeof(p));
  p.B = L->B; p.Hg = L->Hin; p.Wg = L->Win;
  p.Hfull = 4 * p.Hg; p.Wfull = 4 * p.Wg;
  p.tiles_x = (p.Wg + TWX - 1) / TWX;
  p.tiles_per_image = p.tiles_x * ((p.Hg + TH - 1) / TH);
  p.n_tiles = p.B * p.tiles_per_image;
  p.bias = L->bias;
  p.out_f32 = (float*)L->out_hi;
  p.ref_image = L->ref_image; p.sse_partial = L->sse_partial;
  p.grad_out = L->ref_image ? L->grad_out : nullptr; p.grad_scale = L->grad_scale;
  p.dbias = p.grad_out ? L->dbias_out : nullptr;
  p.trace = trace;

  static std::mutex mu;
  static int n_sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int sms = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !n_sms[dev]) {
      cudaError_t e = cudaFuncSetAttribute(deconv3x_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (const char* env = getenv("B2C_TCP_MAX_CTAS")) { const int v = atoi(env); if (v > 0 && v < sms) sms = v; }
      if (sms <= 0) { set_error("deconv3x: no SMs"); return B2C_ECUDA; }
      if (dev >= 0 && dev < 64) n_sms[dev] = sms;
    } else {
      sms = n_sms[dev];
    }
  }
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;

  CUtensorMap mA, mB, mR;
  memset(&mR, 0, sizeof(mR));
  int rc;
  if ((rc = make_box_map(&mA, L->in_hi, L->B, L->Hin, L->Win, C))) return rc;
  if ((rc = make_mat_map(&mB, L->w_hi, 3 * NX, C, NX))) return rc;
  if (XCfg<C>::NR > 0 && L->ref_image && (rc = make_ref_map(&mR, L->ref_image, L->B, p.Hfull, p.Wfull))) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, deconv3x_kernel<C>, mA, mB, mR, p);
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(deconv3x): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  count_launch();
  return check_launch("tc_layer_run(deconv3x)");
}

int launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  switch (L->channels ? L->channels : 128) {
    case 128: return launch_c<128>(L, stream, trace);
    case 192: return launch_c<192>(L, stream, trace);
    case 256: return launch_c<256>(L, stream, trace);
    default: set_error("deconv3x: channels must be 128, 192 or 256"); return B2C_EINVAL;
  }
}

}  // namespace tcp3x
}  // namespace b2c
