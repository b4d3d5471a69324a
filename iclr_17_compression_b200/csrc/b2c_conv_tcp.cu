// Persistent tcgen05 implicit-GEMM kernels for the 128-channel inference layers whose A operand comes straight from
// TMA boxes (CONV5S2, DECONV5S2, DECONV3).  One CTA (or cta_group::2 CTA pair) per SM walks a static round-robin
// list of 8x16 output tiles.  What bounds these layers is the L2 -> shared-memory fill rate (~43 B/clk/SM measured
// chip-wide), not the tensor pipe, so the kernel is organised around operand reuse in shared memory:
//   * ROW-SHARED A BOXES: all taps that differ only in their vertical offset read ONE TMA box of (8 + 2) tile rows;
//     a tap's A operand is that box advanced by whole tile rows (16 pixels x 128 B = 2 KB, a multiple of the 1 KB
//     swizzle atom, so the UMMA descriptor just moves its start address).  A traffic drops 2.1x (5x5/2 conv:
//     ky = 0,2,4 share the even input rows, ky = 1,3 the odd ones), 2.4x (transposed 5x5/2) and 2.4x (9x9/4
//     transposed as 3x3).  Weights (one 128 x 64 tile per tap and k-block) stream through their own ring.
//   * DECONV3 keeps its whole 108 KB weight panel resident in shared memory.
//   * a TMA ring that runs ahead across tile boundaries (no per-tile prologue / pipeline refill),
//   * TWO TMEM accumulator sets, so the workers' epilogue of tile i (x^2 -> gamma MMA -> rsqrt -> bf16 (hi, lo) ->
//     TMA store, or clamp + squared error + fp32 stores) overlaps the MMA main loop of tile i+1,
//   * gamma resident in shared memory; the gamma MMAs of tile i are slipped between main-loop steps of tile i+1 as
//     soon as the workers have written x^2 (polled, never blocking the tensor pipe),
//   * DECONV5S2: the four sub-pixel phases of one spatial tile are consecutive sub-tiles of the same CTA.
// The math, operand layouts and the C ABI are those of b2c_conv_tc.cu (same packed weights).
#include <mutex>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"
#include "b2c_tc_host.cuh"
#include "b2c_tcp_util.cuh"
#include "b2c_tc_ts.cuh"
#include "b2c_entropy_formula.cuh"

#ifndef B2C_TCP_TS_EPI
#define B2C_TCP_TS_EPI 1    /* 0: x^2 through shared memory and block-wide staging, as in round 1 */
#endif

namespace b2c {
namespace tcp {

using namespace b2c::tc;

constexpr int C = 128;
constexpr int NCB = 2;
constexpr int NWORKERS = 256;
constexpr int NTHREADS = 64 + NWORKERS;
constexpr int XS_BYTES = 2 * SLOT_BYTES;            // x^2 tile / output staging (aliased)
constexpr int BOX_ROWS = TH + 2;                    // tile rows per A box
constexpr int ROW_BYTES = TW * 128;                 // one tile row of a k-block: 2 KB
constexpr int A_PLANE = BOX_ROWS * ROW_BYTES;       // 20 KB

enum { OUT_BF16 = 0, OUT_F32_NCHW = 1, OUT_F32_NCHW_PS4 = 2 };

struct PParams {
  int B, Hg, Wg;        // tile grid (pixels): conv -> output dims; deconv -> input dims
  int tiles_x, tiles_per_image;
  int n_tiles;          // B * tiles_per_image
  int Hfull, Wfull, ps;
  const float* bias;
  const float* beta_eff;
  float* out_f32;
  const float* ref_image;   // DECONV3 fused distortion
  float* sse_partial;
  float* grad_out;          // DECONV3 + ref_image (training): grad_scale * (recon - x), fp32 NCHW, or nullptr
  float grad_scale;
  // last analysis layer at inference (CONV5S2, fp32 NCHW output, no GDN): with `lut` the epilogue is the quantizer and
  // the entropy model too (model.py:55-56,71-73) - out_f32 receives z = rint(y), z_nhwc its bf16 NHWC copy (what the
  // synthesis transform reads), bits_partial[B][tiles][8 warps] the bits of every (tile, epilogue warp)
  const float* lut;         // bits[c][z + L], |z| <= L (b2c_entropy_lut_build) or nullptr
  int lut_L;
  BitPtrs bp;               // the 11 Bitparm vectors, for |z| > L
  __nv_bfloat16* z_nhwc;
  float* bits_partial;
  int* range_flag;          // raised when some |z| > 256 (the bf16 copy is inexact beyond), may be nullptr
  long long* trace;         // debug: clock64 timeline of CTA 0: [3 roles][64 marks]
};

struct POutMaps {
  CUtensorMap m[8];     // [0..3] hi plane (deconv: per phase), [4..7] lo plane
};

// ---- A groups: (x tap, k-block[, row parity]) -> one TMA box shared by `ntaps` vertical taps
template <int KIND>
__device__ __forceinline__ int p_ngroups(int phase) {
  if (KIND == B2C_TC_CONV5S2) return 20;
  if (KIND == B2C_TC_DECONV3) return 6;
  return (3 - (phase & 1)) * NCB;
}
// box origin (x_off, y_off) relative to the tile origin in input pixels, taps in the group
template <int KIND>
__device__ __forceinline__ void p_group(int gi, int phase, int& cb, int& x_off, int& y_off, int& ntaps) {
  cb = gi & 1;
  if (KIND == B2C_TC_CONV5S2) {
    const int par = (gi >> 1) & 1, kx = gi >> 2;
    x_off = kx - 2; y_off = par - 2; ntaps = 3 - par;
  } else if (KIND == B2C_TC_DECONV3) {
    x_off = (gi >> 1) - 1; y_off = -1; ntaps = 3;
  } else {
    const int ry = phase >> 1;
    x_off = 1 - (gi >> 1); y_off = ry - 1; ntaps = 3 - ry;
  }
}
// tap t of group gi: weight panel index and the tile-row offset of its A operand inside the box
template <int KIND>
__device__ __forceinline__ void p_tap(int gi, int t, int phase, int& wgroup, int& row_off) {
  if (KIND == B2C_TC_CONV5S2) {
    const int par = (gi >> 1) & 1, kx = gi >> 2;
    wgroup = (2 * t + par) * 5 + kx; row_off = t;
  } else if (KIND == B2C_TC_DECONV3) {
    wgroup = t * 3 + (gi >> 1); row_off = t;
  } else {
    const int ry = phase >> 1, rx = phase & 1, jx = gi >> 1;
    wgroup = (ry + 2 * t) * 5 + rx + 2 * jx; row_off = 2 - ry - t;
  }
}

// ---- DECONV5S2 with RESIDENT A boxes (pair form).  Tap t of sub-pixel row ry reads the input rows starting at
// oy0 + 1 - t whatever ry is, so ONE box per (x shift, k-block) with origin row oy0 - 1 and TH + 2 rows serves all four
// phases of a spatial tile: six boxes per tile instead of twenty (A fill 400 -> 120 KB per tile; once the epilogue is
// off the critical path these layers are bound by the L2 -> shared-memory fill).  The six boxes own fixed slots;
// phases run in the order (ry,rx) = (0,0) (1,0) (0,1) (1,1) so that the x shift -1 boxes (only used when rx = 0) are
// released half way and the next tile's first boxes arrive under the rx = 1 phases.
__device__ __forceinline__ int r_phase(int pi) { return ((pi & 1) << 1) | (pi >> 1); }        // pi -> phase = ry*2 + rx
__device__ __forceinline__ int r_ngroups(int phase) { return (phase & 1) ? 2 * NCB : 3 * NCB; }
// group gi of a phase: k-block cb, x shift index jx (x_off = 1 - jx); rx = 0 visits jx = 2, 0, 1, rx = 1 visits 0, 1
__device__ __forceinline__ void r_group(int gi, int phase, int& cb, int& jx) {
  cb = gi & 1;
  const int j = gi >> 1;
  jx = (phase & 1) ? j : (j == 0 ? 2 : j - 1);
}
__device__ __forceinline__ void r_tap(int jx, int t, int phase, int& wgroup, int& row_off) {
  const int ry = phase >> 1, rx = phase & 1;
  wgroup = (ry + 2 * t) * 5 + rx + 2 * jx;
  row_off = 2 - t;
}

template <int KIND, bool SPLIT3, bool PAIR, int GDN, bool ARES>
struct PCfg {
  static constexpr bool A_RES = ARES;                                           // resident A boxes (6 fixed slots)
  static_assert(!ARES || (KIND == B2C_TC_DECONV5S2 && PAIR && !SPLIT3), "resident A boxes: bf16 pair DECONV5S2 only");
  static constexpr bool B_RES = KIND == B2C_TC_DECONV3;                         // weights resident
  static constexpr int N_TILE = B_RES ? 48 : C;
  static constexpr int B_BYTES = N_TILE * 128 / (PAIR ? 2 : 1);                 // one weight / gamma k-block tile
  static constexpr int A_SLOT = (SPLIT3 ? 2 : 1) * A_PLANE;
  static constexpr int B_SLOT = (SPLIT3 ? 2 : 1) * B_BYTES;
  static constexpr int W_RES = B_RES ? 9 * NCB * B_BYTES : 0;
  static constexpr int GAMMA_BYTES = GDN ? (SPLIT3 ? 2 : 1) * NCB * B_BYTES : 0;
  // x^2 tile / output staging; the TS epilogue (bf16 pair layers, see TS_EPI in the kernel) only stages 2 KB per
  // worker warp at a time, which buys two more weight-ring slots
  static constexpr bool TS_LAYOUT = !SPLIT3 && GDN != 0 && PAIR && B2C_TCP_TS_EPI;
  static constexpr int XS = GDN ? (TS_LAYOUT ? XS_BYTES / 2 : XS_BYTES) : 0;
  static constexpr int EXTRA = 1024 /*align*/ + 1024 /*bias, beta*/ + 512 /*barriers*/;
  static constexpr int MAX_SMEM = 227 * 1024;
  static constexpr int LEFT = MAX_SMEM - EXTRA - GAMMA_BYTES - XS - W_RES;
  static constexpr int NA = A_RES ? 3 * NCB : (SPLIT3 ? (GDN ? 2 : 3) : (B_RES ? 5 : (PAIR ? 4 : 3)));
  static constexpr int NB_RAW = B_RES ? 0 : (LEFT - NA * A_SLOT) / B_SLOT;
  static constexpr int NB = NB_RAW > 12 ? 12 : NB_RAW;     // (weight tiles arrive ~2.6 k clk after their request:
                                                           // a ring of 7-8 four-MMA tiles ran at 70 % of the tensor rate)
  static_assert(B_RES || NB >= 4, "weight ring too shallow");
  static constexpr int SMEM = NA * A_SLOT + NB * B_SLOT + W_RES + GAMMA_BYTES + XS + EXTRA;
  static_assert(SMEM <= MAX_SMEM, "shared memory budget");
  static constexpr int ACC_STRIDE = GDN ? 256 : (N_TILE <= 64 ? 64 : 128);      // TMEM columns per accumulator set
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
};

template <int KIND, bool SPLIT3, int GDN, int OUT_MODE, bool PAIR, bool ARES = false>
__global__ void __launch_bounds__(NTHREADS, 1)
igemm_tcp_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmG_hi, const __grid_constant__ CUtensorMap tmG_lo,
                 const __grid_constant__ POutMaps tmO, const PParams p) {
  using Cfg = PCfg<KIND, SPLIT3, PAIR, GDN, ARES>;
  constexpr int NA = Cfg::NA, NB = Cfg::NB;
  constexpr int NBX = NB > 0 ? NB : 1;
  constexpr bool B_RES = Cfg::B_RES;
  constexpr bool A_RES = Cfg::A_RES;
  constexpr int N_TILE = Cfg::N_TILE;
  constexpr uint32_t A_SLOT = Cfg::A_SLOT, B_SLOT = Cfg::B_SLOT, B_BYTES = Cfg::B_BYTES;
  constexpr int IN_STRIDE = (KIND == B2C_TC_CONV5S2) ? 2 : 1;
  constexpr int NPH = (KIND == B2C_TC_DECONV5S2) ? 4 : 1;
  constexpr int XS_ROUNDS = SPLIT3 ? 2 : 1;
  constexpr uint32_t NPAIR = PAIR ? 2u : 1u;
  constexpr uint32_t ACC_STRIDE = Cfg::ACC_STRIDE;
  constexpr bool PREFETCH = false;   // measured slower (TMA queue shared with the demand loads): conv2 0.250 -> 0.258 ms
  // TS epilogue (bf16 pair layers with a fused GDN / IGDN and NHWC bf16 output = the two transposed 5x5 convs): the
  // worker epilogue - not the main loop - bounded these layers (deconv2: x^2 1.9 k + norm wait 1.9 k + final 2.2 k clk
  // per sub-tile against 3.7 k clk of MMAs, profiles/r2_notes.md).  As in b2c_conv1_ts.cu the x^2 operand of the gamma
  // product is written over the accumulator columns it was read from (tcgen05.mma takes A from TMEM), the workers
  // never meet on a block barrier (per-warp mbarrier arrivals) and every warp stores its own 32 pixels x 64 channels
  // through its own 4 KB staging slice and TMA box.
  constexpr bool TS_EPI = !SPLIT3 && GDN != 0 && OUT_MODE == OUT_BF16 && PAIR && B2C_TCP_TS_EPI;
  constexpr uint32_t NWARPS_W = NWORKERS / 32;
  static_assert(OUT_MODE != OUT_BF16 || GDN != 0, "bf16 NHWC output stages through the x^2 buffer");
  static_assert(!SPLIT3 || PAIR, "the bf16x3 rings only fit as a CTA pair");
  static_assert(!B_RES || (!PAIR && !SPLIT3 && GDN == 0 && OUT_MODE == OUT_F32_NCHW_PS4), "DECONV3 form");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps LDS/STS
  uint8_t* b_ring = smem + NA * A_SLOT;
  uint8_t* wres = b_ring + NB * B_SLOT;                // DECONV3: [tap 9][kb 2] tiles of 48 rows
  uint8_t* gam = wres + Cfg::W_RES;                    // [hi kb0][hi kb1]([lo kb0][lo kb1])
  uint8_t* xs = gam + Cfg::GAMMA_BYTES;                // x^2 A tile, later output staging
  float* s_bias = reinterpret_cast<float*>(xs + Cfg::XS);
  float* s_beta = s_bias + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + 128);
  uint64_t* a_full = bars;                     // [NA]
  uint64_t* a_empty = a_full + NA;             // [NA]
  uint64_t* b_full = a_empty + NA;             // [NB]
  uint64_t* b_empty = b_full + NBX;            // [NB]
  uint64_t* accum_full = b_empty + NBX;        // [2]
  uint64_t* norm_full = accum_full + 2;        // [2]
  uint64_t* acc_empty = norm_full + 2;         // [2]  (leader's: all workers of the pair)
  uint64_t* xs_full = acc_empty + 2;           // (leader's)
  uint64_t* xs_empty = xs_full + 1;
  uint64_t* res_full = xs_empty + 1;           // resident gamma / weights landed (leader's)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int n_units = (p.n_tiles + (int)NPAIR - 1) / (int)NPAIR;           // tiles, or pairs of adjacent tiles
  const int unit0 = (int)(blockIdx.x / NPAIR), unit_stride = (int)(gridDim.x / NPAIR);
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  int tmark = 0;
#ifdef B2C_TCP_FINE_TRACE
#define TCP_FINE() TCP_MARK(2, threadIdx.x == 64)
#define TCP_FINE_MMA() TCP_MARK(1, lane == 0)
#else
#define TCP_FINE() do {} while (0)
#define TCP_FINE_MMA() do {} while (0)
#endif
#define TCP_MARK(role, cond)                                                        \
  do {                                                                              \
    if (tracing && (cond) && tmark < 64) p.trace[(role) * 64 + tmark++] = clock64(); \
  } while (0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&accum_full[i], 1);
      mbar_init(&norm_full[i], 1);
      mbar_init(&acc_empty[i], TS_EPI ? NPAIR * NWARPS_W : NPAIR);   // one elected worker per CTA (after a named
    }                                                                // barrier), or every worker warp (TS epilogue)
    mbar_init(xs_full, TS_EPI ? NPAIR * NWARPS_W : NPAIR);
    mbar_init(xs_empty, 1);
    mbar_init(res_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA_hi);
    if (SPLIT3) tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    if (SPLIT3) tma_prefetch_desc(&tmB_lo);
    if (GDN) { tma_prefetch_desc(&tmG_hi); if (SPLIT3) tma_prefetch_desc(&tmG_lo); }
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + C) {
    const int c = threadIdx.x - 64;
    s_bias[c] = (p.bias && c < (B_RES ? 3 : C)) ? __ldg(p.bias + c) : 0.f;
    s_beta[c] = (GDN && p.beta_eff) ? __ldg(p.beta_eff + c) : 1.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();       // the next layer's prologue may overlap this kernel's tail

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      if (GDN) {   // gamma stays resident: NCB k-blocks (hi[, lo]); a pair stages half of the rows per CTA
        if (leader) mbar_arrive_expect_tx(res_full, (uint32_t)Cfg::GAMMA_BYTES * NPAIR);
#pragma unroll
        for (int g = 0; g < NCB; ++g) {
          if (PAIR) {
            tma_load_2d_pair(gam + g * B_BYTES, &tmG_hi, res_full, g * KBLK, (int)cta_rank * (C / 2));
            if (SPLIT3) tma_load_2d_pair(gam + (NCB + g) * B_BYTES, &tmG_lo, res_full, g * KBLK, (int)cta_rank * (C / 2));
          } else {
            tma_load_2d(gam + g * B_BYTES, &tmG_hi, res_full, g * KBLK, 0);
            if (SPLIT3) tma_load_2d(gam + (NCB + g) * B_BYTES, &tmG_lo, res_full, g * KBLK, 0);
          }
        }
      }
      if (B_RES) {
        mbar_arrive_expect_tx(res_full, (uint32_t)Cfg::W_RES);
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int cb = 0; cb < NCB; ++cb)
            tma_load_2d(wres + (t * NCB + cb) * B_BYTES, &tmB_hi, res_full, cb * KBLK, t * N_TILE);
      }
      griddep_wait();             // the previous layer's output is complete and visible from here on
      uint32_t ga = 0, gb = 0;    // A boxes / weight tiles issued so far
      for (int unit = unit0; unit < n_units; unit += unit_stride) {
        const int tile = unit * (int)NPAIR + (int)cta_rank;
        const int n_img = tile / p.tiles_per_image;          // == B for the masked tile of an odd count: all OOB
        const int trem = tile - n_img * p.tiles_per_image;
        const int tile_y = trem / p.tiles_x, tile_x = trem - tile_y * p.tiles_x;
        const int oy0 = tile_y * TH, ox0 = tile_x * TW;
        // the next unit's input window -> L2 while this unit is computed (its first boxes would otherwise pay the
        // full HBM latency with only NA boxes of run-ahead)
        int pf_img = -1, pf_x = 0, pf_y = 0;
        if (PREFETCH && unit + unit_stride < n_units) {
          const int ntile = (unit + unit_stride) * (int)NPAIR + (int)cta_rank;
          pf_img = ntile / p.tiles_per_image;
          const int nrem = ntile - pf_img * p.tiles_per_image;
          const int nty = nrem / p.tiles_x;
          pf_y = nty * TH * IN_STRIDE; pf_x = (nrem - nty * p.tiles_x) * TW * IN_STRIDE;
          if (pf_img >= p.B) pf_img = -1;
        }
        if (A_RES) {
          const uint32_t ut = (uint32_t)((unit - unit0) / unit_stride);     // spatial tiles this CTA has started
          for (int pi = 0; pi < NPH; ++pi) {
            const int phase = r_phase(pi);
            const int ng = r_ngroups(phase), ntaps = 3 - (phase >> 1);
            for (int gi = 0; gi < ng; ++gi) {
              int cb, jx;
              r_group(gi, phase, cb, jx);
              if (pi == 0) {                      // first visit of this tile's box (jx, cb): fixed slot
                const uint32_t as = (uint32_t)(jx * NCB + cb);
                mbar_wait(&a_empty[as], (ut & 1u) ^ 1u);
                TCP_MARK(0, true);
                uint8_t* a_hi = smem + as * A_SLOT;
                if (leader) mbar_arrive_expect_tx(&a_full[as], 2u * A_SLOT);
                tma_load_4d_pair(a_hi, &tmA_hi, &a_full[as], cb * KBLK, ox0 + 1 - jx, oy0 - 1, n_img);
              }
              for (int t = 0; t < ntaps; ++t, ++gb) {
                int wgroup, row_off;
                r_tap(jx, t, phase, wgroup, row_off);
                const uint32_t bs = gb % NBX;
                mbar_wait(&b_empty[bs], ((gb / NBX) & 1) ^ 1);
                uint8_t* b_hi = b_ring + bs * B_SLOT;
                if (leader) mbar_arrive_expect_tx(&b_full[bs], 2u * B_SLOT);
                tma_load_2d_pair(b_hi, &tmB_hi, &b_full[bs], cb * KBLK, wgroup * C + (int)cta_rank * (C / 2));
              }
            }
          }
          continue;
        }
        for (int phase = 0; phase < NPH; ++phase) {
          const int ng = p_ngroups<KIND>(phase);
          for (int gi = 0; gi < ng; ++gi, ++ga) {
            int cb, x_off, y_off, ntaps;
            p_group<KIND>(gi, phase, cb, x_off, y_off, ntaps);
            if (PREFETCH && pf_img >= 0 && phase == 0) {
              tma_prefetch_4d(&tmA_hi, cb * KBLK, pf_x + x_off, pf_y + y_off, pf_img);
              if (SPLIT3) tma_prefetch_4d(&tmA_lo, cb * KBLK, pf_x + x_off, pf_y + y_off, pf_img);
            }
            const uint32_t as = ga % NA;
            mbar_wait(&a_empty[as], ((ga / NA) & 1) ^ 1);
            TCP_MARK(0, true);
            uint8_t* a_hi = smem + as * A_SLOT;
            const int cx = ox0 * IN_STRIDE + x_off, cy = oy0 * IN_STRIDE + y_off;
            if (PAIR) {
              if (leader) mbar_arrive_expect_tx(&a_full[as], 2u * A_SLOT);
              tma_load_4d_pair(a_hi, &tmA_hi, &a_full[as], cb * KBLK, cx, cy, n_img);
              if (SPLIT3) tma_load_4d_pair(a_hi + A_PLANE, &tmA_lo, &a_full[as], cb * KBLK, cx, cy, n_img);
            } else {
              mbar_arrive_expect_tx(&a_full[as], A_SLOT);
              tma_load_4d(a_hi, &tmA_hi, &a_full[as], cb * KBLK, cx, cy, n_img);
              if (SPLIT3) tma_load_4d(a_hi + A_PLANE, &tmA_lo, &a_full[as], cb * KBLK, cx, cy, n_img);
            }
            if (!B_RES) {
              for (int t = 0; t < ntaps; ++t, ++gb) {
                int wgroup, row_off;
                p_tap<KIND>(gi, t, phase, wgroup, row_off);
                const uint32_t bs = gb % NBX;
                mbar_wait(&b_empty[bs], ((gb / NBX) & 1) ^ 1);
                uint8_t* b_hi = b_ring + bs * B_SLOT;
                if (PAIR) {
                  if (leader) mbar_arrive_expect_tx(&b_full[bs], 2u * B_SLOT);
                  const int brow = wgroup * C + (int)cta_rank * (C / 2);
                  tma_load_2d_pair(b_hi, &tmB_hi, &b_full[bs], cb * KBLK, brow);
                  if (SPLIT3) tma_load_2d_pair(b_hi + B_BYTES, &tmB_lo, &b_full[bs], cb * KBLK, brow);
                } else {
                  mbar_arrive_expect_tx(&b_full[bs], B_SLOT);
                  tma_load_2d(b_hi, &tmB_hi, &b_full[bs], cb * KBLK, wgroup * C);
                  if (SPLIT3) tma_load_2d(b_hi + B_BYTES, &tmB_lo, &b_full[bs], cb * KBLK, wgroup * C);
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (whole warp, one elected lane issues)
    if (leader) {
      constexpr uint32_t idesc = PAIR ? umma_idesc_bf16_m256(N_TILE) : umma_idesc_bf16(N_TILE);
      constexpr uint32_t idesc_g = PAIR ? umma_idesc_bf16_m256(C) : umma_idesc_bf16(C);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint64_t desc0 = umma_desc_sw128(sbase);
      const uint64_t bring_d = desc0 + ((uint32_t)(NA * A_SLOT) >> 4);
      const uint64_t wres_d = bring_d + ((uint32_t)(NB * B_SLOT) >> 4);
      const uint64_t gam_d = wres_d + ((uint32_t)Cfg::W_RES >> 4);
      const uint64_t xs_d = gam_d + ((uint32_t)Cfg::GAMMA_BYTES >> 4);
      uint32_t ga = 0, gb = 0, st = 0;
      int pend_buf = -1;        // accumulator set whose gamma MMAs are still to be issued
      int pend_round = 0;
      uint32_t xs_cnt = 0;      // completed x^2 rounds (parity of xs_full)

      // issue the gamma MMAs of one x^2 round (SPLIT3: k-block `round`, hi/lo; bf16: both k-blocks)
      auto issue_gdn_round = [&]() {
        tcgen05_fence_after();
        const uint32_t d = tb + (uint32_t)pend_buf * ACC_STRIDE + 128u;
        if (elect_one()) {
          if (SPLIT3) {
            const uint64_t a_hi = xs_d, a_lo = xs_d + (SLOT_BYTES >> 4);
            const uint64_t b_hi = gam_d + ((uint32_t)pend_round * B_BYTES >> 4);
            const uint64_t b_lo = gam_d + ((uint32_t)(NCB + pend_round) * B_BYTES >> 4);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t acc = (pend_round > 0 || k4 > 0) ? 1u : 0u;
              if (PAIR) {
                umma_bf16_pair(d, a_hi + 2 * k4, b_hi + 2 * k4, idesc_g, acc);
                umma_bf16_pair(d, a_hi + 2 * k4, b_lo + 2 * k4, idesc_g, 1u);
                umma_bf16_pair(d, a_lo + 2 * k4, b_hi + 2 * k4, idesc_g, 1u);
              } else {
                umma_bf16(d, a_hi + 2 * k4, b_hi + 2 * k4, idesc_g, acc);
                umma_bf16(d, a_hi + 2 * k4, b_lo + 2 * k4, idesc_g, 1u);
                umma_bf16(d, a_lo + 2 * k4, b_hi + 2 * k4, idesc_g, 1u);
              }
            }
          } else if (TS_EPI) {
            // x^2 of worker half h = kk >> 2 sits in the accumulator's own columns 64h .. 64h + 31 (two bf16 per column)
            const uint32_t a0 = tb + (uint32_t)pend_buf * ACC_STRIDE;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t b = gam_d + ((uint32_t)(kk >> 2) * B_BYTES >> 4) + 2 * (kk & 3);
              umma_bf16_pair_ts(d, a0 + (uint32_t)((kk >> 2) * 64 + (kk & 3) * 8), b, idesc_g, kk > 0 ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t a = xs_d + ((uint32_t)(kk >> 2) * SLOT_BYTES >> 4) + 2 * (kk & 3);
              const uint64_t b = gam_d + ((uint32_t)(kk >> 2) * B_BYTES >> 4) + 2 * (kk & 3);
              if (PAIR) umma_bf16_pair(d, a, b, idesc_g, kk > 0 ? 1u : 0u);
              else umma_bf16(d, a, b, idesc_g, kk > 0 ? 1u : 0u);
            }
          }
          const bool last = pend_round == XS_ROUNDS - 1;
          if (PAIR) {
            if (last) umma_commit_pair(&norm_full[pend_buf]);
            else umma_commit_pair(xs_empty);
          } else {
            if (last) umma_commit(&norm_full[pend_buf]);
            else umma_commit(xs_empty);
          }
        }
        __syncwarp();
        ++xs_cnt;
        if (++pend_round == XS_ROUNDS) pend_buf = -1;
      };
      auto xs_ready = [&]() -> bool {
        uint32_t ok = mbar_test_wait(xs_full, xs_cnt & 1u, PAIR);
        return __shfl_sync(0xffffffffu, ok, 0) != 0;
      };

      if (GDN || B_RES) mbar_wait(res_full, 0);
      for (int unit = unit0; unit < n_units; unit += unit_stride) {
        for (int pi = 0; pi < NPH; ++pi, ++st) {
          const int phase = A_RES ? r_phase(pi) : pi;
          const uint32_t buf = st & 1u;
          const int ng = A_RES ? r_ngroups(phase) : p_ngroups<KIND>(phase);
          // the workers must have drained this accumulator set (sub-tile st-2)
          TCP_FINE_MMA();
          if (PAIR) mbar_wait_cluster(&acc_empty[buf], ((st >> 1) & 1u) ^ 1u);
          else mbar_wait(&acc_empty[buf], ((st >> 1) & 1u) ^ 1u);
          tcgen05_fence_after();
          TCP_FINE_MMA();
          const uint32_t d = tb + buf * ACC_STRIDE;
          for (int gi = 0; gi < ng; ++gi, ++ga) {
            int cb, x_off, y_off, ntaps, jx = 0;
            if (A_RES) {
              r_group(gi, phase, cb, jx);
              ntaps = 3 - (phase >> 1);
            } else {
              p_group<KIND>(gi, phase, cb, x_off, y_off, ntaps);
            }
            if (GDN && pend_buf >= 0 && xs_ready()) issue_gdn_round();
            const uint32_t ut = (uint32_t)((unit - unit0) / unit_stride);
            const uint32_t as = A_RES ? (uint32_t)(jx * NCB + cb) : ga % NA;
            if (!A_RES) mbar_wait(&a_full[as], (ga / NA) & 1);
            else if (pi == 0) mbar_wait(&a_full[as], ut & 1u);       // later phases re-use the resident box
            if (!B_RES) {   // all weight tiles of the group (the producer fills them without waiting on anything else)
#pragma unroll
              for (int t = 0; t < 3; ++t)
                if (t < ntaps) mbar_wait(&b_full[(gb + t) % NBX], ((gb + t) / NBX) & 1);
            }
            tcgen05_fence_after();
            TCP_MARK(1, lane == 0);
            const uint64_t a_box = desc0 + ((as * A_SLOT) >> 4);
            const bool last_group = gi == ng - 1;
            // ONE elected region per group: up to 3 taps x 4 k-slices x (1 | 3) MMAs and their commits back to back
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                if (t < ntaps) {
                  int wgroup, row_off;
                  if (A_RES) r_tap(jx, t, phase, wgroup, row_off);
                  else p_tap<KIND>(gi, t, phase, wgroup, row_off);
                  const uint32_t bs = (gb + t) % NBX;
                  const uint64_t a_hi = a_box + ((uint32_t)(row_off * ROW_BYTES) >> 4), a_lo = a_hi + (A_PLANE >> 4);
                  const uint64_t b_hi = B_RES ? wres_d + ((uint32_t)((wgroup * NCB + cb) * B_BYTES) >> 4)
                                              : bring_d + ((bs * B_SLOT) >> 4);
                  const uint64_t b_lo = b_hi + (B_BYTES >> 4);
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    const uint32_t acc = (gi > 0 || t > 0 || k4 > 0) ? 1u : 0u;
                    if (PAIR) {
                      umma_bf16_pair(d, a_hi + 2 * k4, b_hi + 2 * k4, idesc, acc);
                      if (SPLIT3) {
                        umma_bf16_pair(d, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1u);
                        umma_bf16_pair(d, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1u);
                      }
                    } else {
                      umma_bf16(d, a_hi + 2 * k4, b_hi + 2 * k4, idesc, acc);
                      if (SPLIT3) {
                        umma_bf16(d, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1u);
                        umma_bf16(d, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1u);
                      }
                    }
                  }
                  if (!B_RES) {
                    if (PAIR) umma_commit_pair(&b_empty[bs]);
                    else umma_commit(&b_empty[bs]);
                  }
                }
              }
              // resident boxes are handed back after their last phase: x shift -1 after the second rx = 0 phase
              const bool a_done = !A_RES || (jx == 2 ? pi == 1 : pi == NPH - 1);
              if (PAIR) {
                if (a_done) umma_commit_pair(&a_empty[as]);
                if (last_group) umma_commit_pair(&accum_full[buf]);
              } else {
                umma_commit(&a_empty[as]);
                if (last_group) umma_commit(&accum_full[buf]);
              }
            }
            __syncwarp();
            gb += (uint32_t)ntaps;
          }
          if (GDN) {
            // the previous sub-tile's gamma product must be out before this one's can be queued
            TCP_FINE_MMA();
            while (pend_buf >= 0) {
              if (PAIR) mbar_wait_cluster(xs_full, xs_cnt & 1u);
              else mbar_wait(xs_full, xs_cnt & 1u);
              issue_gdn_round();
            }
            pend_buf = (int)buf;
            pend_round = 0;
          }
        }
      }
      if (GDN) {
        while (pend_buf >= 0) {
          if (PAIR) mbar_wait_cluster(xs_full, xs_cnt & 1u);
          else mbar_wait(xs_full, xs_cnt & 1u);
          issue_gdn_round();
        }
      }
    }
  } else {
    // ================================================================ row workers (8 warps): thread = (pixel r, half)
    // Worker -> MMA-warp signals (x^2 written, accumulator drained) are ONE mbarrier arrival per CTA: every worker
    // fences, the workers meet on a named barrier, one elected thread arrives (a per-thread release.cluster arrive
    // costs a MEMBAR.GPU each and dominated the epilogue).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ty = r / TW, tx = r - ty * TW;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool elected = warp == 2 && lane == 0;
    bool store_pending = false;        // elected thread: a TMA store may still be reading the staging buffer
    uint32_t st = 0;
    griddep_wait();                    // before any global read / write of activations
#define WORKER_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(NWORKERS) : "memory")
    for (int unit = unit0; unit < n_units; unit += unit_stride) {
      const int tile = unit * (int)NPAIR + (int)cta_rank;
      const int n_img = tile / p.tiles_per_image;
      const int trem = tile - n_img * p.tiles_per_image;
      const int tile_y = trem / p.tiles_x, tile_x = trem - tile_y * p.tiles_x;
      const int oy0 = tile_y * TH, ox0 = tile_x * TW;
      const int oy = oy0 + ty, ox = ox0 + tx;
      const bool valid = n_img < p.B && oy < p.Hg && ox < p.Wg;

      // DECONV3 fused distortion: fetch this thread's 2 rows x 3 channels x 4 pixels of the reference image before
      // waiting for the accumulator, so the HBM latency hides behind the main loop
      float4 xref[OUT_MODE == OUT_F32_NCHW_PS4 ? 6 : 1];
      if (OUT_MODE == OUT_F32_NCHW_PS4) {
#pragma unroll
        for (int i = 0; i < 6; ++i) xref[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.ref_image && valid) {
#pragma unroll
          for (int o = 0; o < 3; ++o)
#pragma unroll
            for (int ryl = 0; ryl < 2; ++ryl)
              xref[o * 2 + ryl] = __ldg(reinterpret_cast<const float4*>(
                  p.ref_image + (((size_t)n_img * 3 + o) * p.Hfull + (oy * 4 + half * 2 + ryl)) * p.Wfull + ox * 4));
        }
      }

      for (int pi = 0; pi < NPH; ++pi, ++st) {
        const int phase = A_RES ? r_phase(pi) : pi;
        const uint32_t buf = st & 1u, par = (st >> 1) & 1u;
        const uint32_t t_acc = t_lane + buf * ACC_STRIDE, t_nrm = t_acc + 128u;
        TCP_MARK(2, threadIdx.x == 64);
        mbar_wait(&accum_full[buf], par);
        tcgen05_fence_after();
        TCP_MARK(2, threadIdx.x == 64);

        // x = acc + bias of this thread's 64 channels stays in registers between the two passes (TMEM reads are
        // 64 B/clk per SM: re-reading the accumulator would cost as much as the gamma MMAs)
        //   SPLIT3: xk[g*16 + i] = channels g*64 + half*32 + 2i, +1     bf16: xk[hh*16 + i] = half*64 + hh*32 + 2i, +1
        uint64_t xk[GDN ? 32 : 1];
        if constexpr (TS_EPI) {
          // ---- x^2 of this thread's 64 channels over accumulator columns it has itself read: xk[hh*16 + i] = channels
          // half*64 + hh*32 + 2i, +1; the 32 packed words land in columns half*64 .. half*64 + 31
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t raw[32], w[16];
            tmem_ld32(t_acc + (uint32_t)(half * 64 + hh * 32), raw);
            tmem_ld_wait();
            const uint64_t* bias2 = reinterpret_cast<const uint64_t*>(s_bias + half * 64 + hh * 32);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t x2 = add2(pk2(raw[2 * i], raw[2 * i + 1]), bias2[i]);
              xk[hh * 16 + i] = x2;
              float v0, v1;
              upk2(mul2(x2, x2), v0, v1);
              w[i] = cvt_bf16x2(v0, v1);
            }
            tmem_st16(t_acc + (uint32_t)(half * 64 + hh * 16), w);
          }
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta0(xs_full);
          TCP_MARK(2, threadIdx.x == 64);
          mbar_wait(&norm_full[buf], par);
          tcgen05_fence_after();
          TCP_MARK(2, threadIdx.x == 64);
          // ---- final pass: y = x * sqrt(beta + n) (IGDN) or x * rsqrt(beta + n) (GDN) -> this warp's 2 KB staging
          // slice (32 pixels x 32 channels, 64-byte rows, SWIZZLE_64B) -> one TMA store of a {32 ch, 16, 2} box per
          // 32-channel block; the slice is re-used for the second block once the first store has read it
          uint8_t* my_stg = xs + (uint32_t)(warp - 2) * 2048u;
          const uint32_t st_row = (uint32_t)lane * 64u, st_x = ((uint32_t)lane >> 1) & 3u;
          const int mi = (KIND == B2C_TC_DECONV5S2) ? phase : 0;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t rawn[32];
            tmem_ld32(t_nrm + (uint32_t)(half * 64 + hh * 32), rawn);
            tmem_ld_wait();
            if (hh == 1) tcgen05_fence_before();
            const uint64_t* beta2 = reinterpret_cast<const uint64_t*>(s_beta + half * 64 + hh * 32);
            uint32_t hp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float n0, n1, y0, y1;
              upk2(add2(pk2(rawn[2 * i], rawn[2 * i + 1]), beta2[i]), n0, n1);
              const uint64_t sc = (GDN == 2) ? pkf2(sqrt_fast(n0), sqrt_fast(n1)) : pkf2(rsqrt_fast(n0), rsqrt_fast(n1));
              upk2(mul2(xk[hh * 16 + i], sc), y0, y1);
              hp[i] = cvt_bf16x2(y0, y1);
            }
            if (lane == 0 && store_pending) { tma_store_wait_read(); store_pending = false; }
            __syncwarp();
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
              *reinterpret_cast<uint4*>(my_stg + st_row + ((((uint32_t)ci) ^ st_x) << 4)) =
                  make_uint4(hp[ci * 4 + 0], hp[ci * 4 + 1], hp[ci * 4 + 2], hp[ci * 4 + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (hh == 1) mbar_arrive_cta0(&acc_empty[buf]);       // this warp's TMEM reads of the set are complete
              tma_store_4d(&tmO.m[4 + mi], my_stg, half * KBLK + hh * 32, ox0, oy0 + 2 * q, n_img);
              tma_store_commit();
              store_pending = true;
            }
          }
          continue;
        }
        if (GDN) {
          // ---- x^2 as the K-major swizzled A tile of the gamma product
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t raw[32];
            const int c0 = SPLIT3 ? g * 64 + half * 32 : half * 64 + g * 32;
            tmem_ld32(t_acc + (uint32_t)c0, raw);
            tmem_ld_wait();
            TCP_FINE();
            if (g == 0) {                      // the previous sub-tile's (last) store must have left the buffer
              if (elected && store_pending) { tma_store_wait_read(); store_pending = false; }
              WORKER_SYNC();
              TCP_FINE();
            } else if (SPLIT3) {
              mbar_wait(xs_empty, st & 1u);    // the gamma MMAs of round 0 have read it
            }
            const uint64_t* bias2 = reinterpret_cast<const uint64_t*>(s_bias + c0);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t x2 = add2(pk2(raw[ci * 8 + 2 * j], raw[ci * 8 + 2 * j + 1]), bias2[ci * 4 + j]);
                xk[g * 16 + ci * 4 + j] = x2;
                float v0, v1;
                upk2(mul2(x2, x2), v0, v1);
                h[j] = cvt_bf16x2(v0, v1);
                if (SPLIT3) l[j] = lo_of(v0, v1, h[j]);
              }
              if (SPLIT3) {
                sts128(xs, r, half * 4 + ci, h[0], h[1], h[2], h[3]);
                sts128(xs + SLOT_BYTES, r, half * 4 + ci, l[0], l[1], l[2], l[3]);
              } else {
                sts128(xs + half * SLOT_BYTES, r, g * 4 + ci, h[0], h[1], h[2], h[3]);
              }
            }
            TCP_FINE();
            if (SPLIT3 || g == 1) {
              fence_proxy_async_smem();
              TCP_FINE();
              WORKER_SYNC();
              TCP_FINE();
              if (elected) {
                if (PAIR) mbar_arrive_cta0(xs_full);
                else mbar_arrive(xs_full);
              }
            }
          }
          TCP_MARK(2, threadIdx.x == 64);
          mbar_wait(&norm_full[buf], par);     // also: every gamma MMA has finished reading the x^2 tile
          tcgen05_fence_after();
          TCP_MARK(2, threadIdx.x == 64);
        }

        // ---- final pass: y = (acc + bias) [* or / sqrt(beta + norm)]
        const int oyf = oy * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase >> 1) : 0);
        const int oxf = ox * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase & 1) : 0);
        if (OUT_MODE == OUT_F32_NCHW_PS4) {
          // DECONV3: column = (ry*4+rx)*3 + o -> out[n, o, 4*oy+ry, 4*ox+rx]; this thread owns ry = 2*half, 2*half+1.
          // With a reference image: out = clamp(r, 0, 1) (model.py:59) and the tile's partial sums of (r - x)^2
          // (model.py:61) and (clamp(r) - x)^2 (train.py:172) go to sse_partial.
          uint32_t raw[24];
          tmem_ld8(t_acc + (uint32_t)(half * 24), *reinterpret_cast<uint32_t(*)[8]>(&raw[0]));
          tmem_ld8(t_acc + (uint32_t)(half * 24 + 8), *reinterpret_cast<uint32_t(*)[8]>(&raw[8]));
          tmem_ld8(t_acc + (uint32_t)(half * 24 + 16), *reinterpret_cast<uint32_t(*)[8]>(&raw[16]));
          tmem_ld_wait();
          tcgen05_fence_before();
          WORKER_SYNC();
          if (elected) mbar_arrive(&acc_empty[buf]);
          float s0 = 0.f, s1 = 0.f;
          if (valid) {
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              const float bv = s_bias[o];
#pragma unroll
              for (int ryl = 0; ryl < 2; ++ryl) {
                const int ry = half * 2 + ryl;
                float4 v4 = make_float4(__uint_as_float(raw[(ryl * 4 + 0) * 3 + o]) + bv,
                                        __uint_as_float(raw[(ryl * 4 + 1) * 3 + o]) + bv,
                                        __uint_as_float(raw[(ryl * 4 + 2) * 3 + o]) + bv,
                                        __uint_as_float(raw[(ryl * 4 + 3) * 3 + o]) + bv);
                const size_t off = (((size_t)n_img * 3 + o) * p.Hfull + (oy * 4 + ry)) * p.Wfull + ox * 4;
                if (p.ref_image) {
                  const float4 x4 = xref[o * 2 + ryl];
                  const float4 c4 = make_float4(fminf(fmaxf(v4.x, 0.f), 1.f), fminf(fmaxf(v4.y, 0.f), 1.f),
                                                fminf(fmaxf(v4.z, 0.f), 1.f), fminf(fmaxf(v4.w, 0.f), 1.f));
                  float dd;
                  dd = v4.x - x4.x; s0 = fmaf(dd, dd, s0); dd = v4.y - x4.y; s0 = fmaf(dd, dd, s0);
                  dd = v4.z - x4.z; s0 = fmaf(dd, dd, s0); dd = v4.w - x4.w; s0 = fmaf(dd, dd, s0);
                  dd = c4.x - x4.x; s1 = fmaf(dd, dd, s1); dd = c4.y - x4.y; s1 = fmaf(dd, dd, s1);
                  dd = c4.z - x4.z; s1 = fmaf(dd, dd, s1); dd = c4.w - x4.w; s1 = fmaf(dd, dd, s1);
                  if (p.grad_out)   // d/d recon of (grad_scale / 2) * sum (recon - x)^2: the distortion backward, fused
                    *reinterpret_cast<float4*>(p.grad_out + off) =
                        make_float4(p.grad_scale * (v4.x - x4.x), p.grad_scale * (v4.y - x4.y),
                                    p.grad_scale * (v4.z - x4.z), p.grad_scale * (v4.w - x4.w));
                  v4 = c4;
                }
                *reinterpret_cast<float4*>(p.out_f32 + off) = v4;
              }
            }
          }
          if (p.ref_image) {   // fixed-order CTA reduction: warp shuffle, then the 8 worker warps through smem
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              s0 += __shfl_xor_sync(0xffffffffu, s0, o);
              s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            float* red = s_beta + (st & 1u) * 16;   // s_beta is unused by DECONV3 (no GDN); double-buffered per tile
            if (lane == 0) { red[(warp - 2) * 2] = s0; red[(warp - 2) * 2 + 1] = s1; }
            WORKER_SYNC();
            if (elected && n_img < p.B) {
              float t0 = 0.f, t1 = 0.f;
#pragma unroll
              for (int w = 0; w < NWORKERS / 32; ++w) { t0 += red[w * 2]; t1 += red[w * 2 + 1]; }
              float* dst = p.sse_partial + ((size_t)n_img * p.tiles_per_image + trem) * 2;
              dst[0] = t0;
              dst[1] = t1;
            }
          }
        } else if (OUT_MODE == OUT_F32_NCHW) {
          const bool latent = !GDN && KIND == B2C_TC_CONV5S2 && p.lut != nullptr;
          float bits_sum = 0.f;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int c0 = half * 64 + h * 32;
            uint32_t raw[32], rawn[32];
            tmem_ld32(t_acc + (uint32_t)c0, raw);
            if (GDN) tmem_ld32(t_nrm + (uint32_t)c0, rawn);
            tmem_ld_wait();
            if (h == 1) {                        // last TMEM read of this sub-tile: hand the accumulator set back
              tcgen05_fence_before();
              WORKER_SYNC();
              if (elected) {
                if (PAIR) mbar_arrive_cta0(&acc_empty[buf]);
                else mbar_arrive(&acc_empty[buf]);
              }
            }
            if (!GDN && KIND == B2C_TC_CONV5S2 && latent) {
              // quantizer + entropy model in the epilogue: z = rint(y) (half to even, model.py:56), its bits from the
              // per-channel table (one cached read per element; the formula itself only beyond the table), z as fp32
              // NCHW (the API tensor) and as bf16 NHWC (this thread's 32 channels = 64 contiguous bytes)
              if (valid) {
                const float fL = (float)p.lut_L;
                const int Wt = 2 * p.lut_L + 1;
                const float* row = p.lut + (size_t)c0 * Wt + p.lut_L;
                float zv[32], amax = 0.f, bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  zv[j] = rintf(__uint_as_float(raw[j]) + s_bias[c0 + j]);
                  amax = fmaxf(amax, fabsf(zv[j]));
                  const int idx = (int)fminf(fmaxf(zv[j], -fL), fL);
                  bs[j & 3] += __ldg(row + (size_t)j * Wt + idx);
                  p.out_f32[(((size_t)n_img * C + c0 + j) * p.Hfull + oyf) * p.Wfull + oxf] = zv[j];
                }
                float bsum = (bs[0] + bs[1]) + (bs[2] + bs[3]);
                if (!(amax <= fL)) {             // rare: outside the table (or NaN) - the same formula directly
                  if (p.range_flag && !(amax <= BF16_EXACT_INT)) *p.range_flag = 1;
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (!(fabsf(zv[j]) <= fL)) {
                      const int idx = (int)fminf(fmaxf(zv[j], -fL), fL);     // (NaN -> -L, as above)
                      bsum += bits_outside_table(p.bp, c0 + j, zv[j]) - __ldg(row + (size_t)j * Wt + idx);
                    }
                }
                bits_sum += bsum;
                uint4* dst = reinterpret_cast<uint4*>(p.z_nhwc + (((size_t)n_img * p.Hfull + oyf) * p.Wfull + oxf) * C + c0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  dst[k] = make_uint4(cvt_bf16x2(zv[8 * k], zv[8 * k + 1]), cvt_bf16x2(zv[8 * k + 2], zv[8 * k + 3]),
                                      cvt_bf16x2(zv[8 * k + 4], zv[8 * k + 5]), cvt_bf16x2(zv[8 * k + 6], zv[8 * k + 7]));
              }
              continue;
            }
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float xv = __uint_as_float(raw[j]) + s_bias[c0 + j];
                if (GDN) {
                  const float nb = __uint_as_float(rawn[j]) + s_beta[c0 + j];
                  const float rs = rsqrt_fast(nb);
                  xv = (GDN == 2) ? xv * (nb * rs) : xv * rs;
                }
                p.out_f32[(((size_t)n_img * C + c0 + j) * p.Hfull + oyf) * p.Wfull + oxf] = xv;
              }
            }
          }
          if (!GDN && KIND == B2C_TC_CONV5S2 && latent) {     // bits of this (tile, warp): fixed-order butterfly
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bits_sum += __shfl_xor_sync(0xffffffffu, bits_sum, o);
            if (lane == 0 && n_img < p.B)
              p.bits_partial[((size_t)n_img * p.tiles_per_image + trem) * (NWORKERS / 32) + (warp - 2)] = bits_sum;
          }
        } else {
          // NHWC bf16 through the (now idle) x^2 buffer + TMA stores: hi plane first, lo plane second
          uint32_t lo_pack[SPLIT3 ? 32 : 1];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = SPLIT3 ? h * 64 + half * 32 : half * 64 + h * 32;
            uint32_t rawn[32];
            tmem_ld32(t_nrm + (uint32_t)c0, rawn);
            tmem_ld_wait();
            TCP_FINE();
            if (h == 1) tcgen05_fence_before();
            const uint64_t* beta2 = reinterpret_cast<const uint64_t*>(s_beta + c0);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              uint32_t hp[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float n0, n1, y0, y1;
                upk2(add2(pk2(rawn[ci * 8 + 2 * j], rawn[ci * 8 + 2 * j + 1]), beta2[ci * 4 + j]), n0, n1);
                const uint64_t sc = (GDN == 2) ? pkf2(sqrt_fast(n0), sqrt_fast(n1)) : pkf2(rsqrt_fast(n0), rsqrt_fast(n1));
                upk2(mul2(xk[h * 16 + ci * 4 + j], sc), y0, y1);
                hp[j] = cvt_bf16x2(y0, y1);
                if (SPLIT3) lo_pack[h * 16 + ci * 4 + j] = lo_of(y0, y1, hp[j]);
              }
              if (SPLIT3) sts128(xs + h * SLOT_BYTES, r, half * 4 + ci, hp[0], hp[1], hp[2], hp[3]);
              else sts128(xs + half * SLOT_BYTES, r, h * 4 + ci, hp[0], hp[1], hp[2], hp[3]);
            }
          }
          const int mi = (KIND == B2C_TC_DECONV5S2) ? phase : 0;
          TCP_FINE();
          fence_proxy_async_smem();
          TCP_FINE();
          WORKER_SYNC();                        // staging complete; every worker has finished its TMEM reads
          TCP_FINE();
          if (elected) {
            if (PAIR) mbar_arrive_cta0(&acc_empty[buf]);
            else mbar_arrive(&acc_empty[buf]);
            TCP_FINE();
#ifndef B2C_TCP_NO_STORE   /* timing experiment only */
            tma_store_4d(&tmO.m[mi], xs, 0, ox0, oy0, n_img);
            tma_store_4d(&tmO.m[mi], xs + SLOT_BYTES, KBLK, ox0, oy0, n_img);
#endif
            tma_store_commit();
            if (SPLIT3) tma_store_wait_read();
            else store_pending = true;          // waited for just before the buffer is written again
          }
          if (SPLIT3) {
            WORKER_SYNC();
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int ci = 0; ci < 4; ++ci)
                sts128(xs + h * SLOT_BYTES, r, half * 4 + ci, lo_pack[h * 16 + ci * 4 + 0], lo_pack[h * 16 + ci * 4 + 1],
                       lo_pack[h * 16 + ci * 4 + 2], lo_pack[h * 16 + ci * 4 + 3]);
            fence_proxy_async_smem();
            WORKER_SYNC();
            if (elected) {
              tma_store_4d(&tmO.m[4 + mi], xs, 0, ox0, oy0, n_img);
              tma_store_4d(&tmO.m[4 + mi], xs + SLOT_BYTES, KBLK, ox0, oy0, n_img);
              tma_store_commit();
              store_pending = true;
            }
          }
        }
      }
    }
    if ((TS_EPI ? lane == 0 : elected) && store_pending) tma_store_wait_read();
    TCP_MARK(2, threadIdx.x == 64);
#undef WORKER_SYNC
  }

  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
// NHWC bf16 activation [B][H][W][128]: box = 64 channels x TW x (TH + 2) pixels with traversal stride S
static int make_act_box_map(CUtensorMap* m, const void* base, int B, int H, int W, int S) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return B2C_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)KBLK, (cuuint32_t)(TW * S), (cuuint32_t)(BOX_ROWS * S), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)S, (cuuint32_t)S, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation box) failed: %d", (int)r); return B2C_ECUDA; }
  return B2C_OK;
}

template <int KIND, bool SPLIT3, int GDN, int OUT_MODE, bool PAIR, bool ARES = false>
static int launch_p(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  using Cfg = PCfg<KIND, SPLIT3, PAIR, GDN, ARES>;
  constexpr size_t SMEM = Cfg::SMEM;
  constexpr int N_TILE = Cfg::N_TILE;
  auto kern = igemm_tcp_kernel<KIND, SPLIT3, GDN, OUT_MODE, PAIR, ARES>;

  PParams p;
  memset(&p, 0, sizeof(p));
  p.B = L->B;
  if (KIND == B2C_TC_CONV5S2) { p.Hg = (L->Hin - 1) / 2 + 1; p.Wg = (L->Win - 1) / 2 + 1; p.ps = 1; }
  else { p.Hg = L->Hin; p.Wg = L->Win; p.ps = (KIND == B2C_TC_DECONV3) ? 4 : 2; }
  p.Hfull = p.Hg * p.ps; p.Wfull = p.Wg * p.ps;
  p.tiles_x = (p.Wg + TW - 1) / TW;
  p.tiles_per_image = p.tiles_x * ((p.Hg + TH - 1) / TH);
  p.n_tiles = p.B * p.tiles_per_image;
  p.bias = L->bias; p.beta_eff = L->beta_eff;
  p.out_f32 = (float*)L->out_hi;
  p.ref_image = L->ref_image; p.sse_partial = L->sse_partial;
  p.grad_out = L->ref_image ? L->grad_out : nullptr; p.grad_scale = L->grad_scale;
  if (L->lut) {
    p.lut = L->lut; p.lut_L = L->lut_half_range;
    for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) p.bp.p[i] = L->bit_params[i];
    p.z_nhwc = (__nv_bfloat16*)L->z_nhwc_out; p.bits_partial = L->bits_partial; p.range_flag = L->range_flag;
  }
  p.trace = trace;

  // grid: one CTA (pair) per SM, never more than can be co-resident (a persistent CTA that waits for a slot
  // would serialise its whole share of the tiles)
  static std::mutex mu;
  static int max_ctas[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int resident = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !max_ctas[dev]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      int sms = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      int n = sms;
      if (PAIR) {
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(sms & ~1);
        qc.blockDim = dim3(NTHREADS);
        qc.dynamicSmemBytes = SMEM;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        qc.attrs = qa; qc.numAttrs = 1;
        int ncl = 0;
        e = cudaOccupancyMaxActiveClusters(&ncl, kern, &qc);
        if (e != cudaSuccess || ncl <= 0) { set_error("cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
        n = 2 * ncl;
        if (n > (sms & ~1)) n = sms & ~1;
      }
      if (const char* env = getenv("B2C_TCP_MAX_CTAS")) { const int v = atoi(env); if (v > 0 && v < n) n = PAIR ? (v & ~1) : v; }
      if (n <= 0) { set_error("tcp: no resident CTAs"); return B2C_ECUDA; }
      if (dev >= 0 && dev < 64) max_ctas[dev] = n;
      resident = n;
    } else {
      resident = max_ctas[dev];
    }
  }
  const int n_units = (p.n_tiles + (PAIR ? 1 : 0)) / (PAIR ? 2 : 1);
  int grid = n_units * (PAIR ? 2 : 1);
  if (grid > resident) grid = resident;

  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mG_hi, mG_lo;
  memset(&mA_lo, 0, sizeof(mA_lo)); memset(&mB_lo, 0, sizeof(mB_lo));
  memset(&mG_hi, 0, sizeof(mG_hi)); memset(&mG_lo, 0, sizeof(mG_lo));
  int rc;
  constexpr int S = (KIND == B2C_TC_CONV5S2) ? 2 : 1;
  if ((rc = make_act_box_map(&mA_hi, L->in_hi, L->B, L->Hin, L->Win, S))) return rc;
  if (SPLIT3 && (rc = make_act_box_map(&mA_lo, L->in_lo, L->B, L->Hin, L->Win, S))) return rc;
  constexpr int BROWS = PAIR ? N_TILE / 2 : N_TILE;
  const int wgroups = (KIND == B2C_TC_DECONV3) ? 9 : 25;
  if ((rc = make_mat_map(&mB_hi, L->w_hi, wgroups * N_TILE, C, BROWS))) return rc;
  if (SPLIT3 && (rc = make_mat_map(&mB_lo, L->w_lo, wgroups * N_TILE, C, BROWS))) return rc;
  if (GDN) {
    if ((rc = make_mat_map(&mG_hi, L->gamma_hi, C, C, PAIR ? C / 2 : C))) return rc;
    if (SPLIT3 && (rc = make_mat_map(&mG_lo, L->gamma_lo, C, C, PAIR ? C / 2 : C))) return rc;
  }
  POutMaps mO;
  memset(&mO, 0, sizeof(mO));
  if (OUT_MODE == OUT_BF16) {
    const int nph = (KIND == B2C_TC_DECONV5S2) ? 4 : 1;
    for (int ph = 0; ph < nph; ++ph) {
      const int ry = nph == 4 ? (ph >> 1) : 0, rx = nph == 4 ? (ph & 1) : 0;
      if ((rc = make_out_map(&mO.m[ph], L->out_hi, L->B, p.Hg, p.Wg, p.Hfull, p.Wfull, p.ps, ry, rx, C))) return rc;
      if (SPLIT3 && (rc = make_out_map(&mO.m[4 + ph], L->out_lo, L->B, p.Hg, p.Wg, p.Hfull, p.Wfull, p.ps, ry, rx, C)))
        return rc;
      // TS epilogue: the same lattice with per-warp boxes (32 channels x 16 x 2 pixels, SWIZZLE_64B) in the (otherwise
      // unused) lo maps
      if (!SPLIT3 && GDN != 0 && PAIR &&
          (rc = make_out_map(&mO.m[4 + ph], L->out_hi, L->B, p.Hg, p.Wg, p.Hfull, p.Wfull, p.ps, ry, rx, C, 2, 32)))
        return rc;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mA_hi, mA_lo, mB_hi, mB_lo, mG_hi, mG_lo, mO, p);
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(tcp): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  count_launch();
  return check_launch("tc_layer_run(persistent)");
}

// Returns 1 when the layer is not one the persistent kernels cover (the caller then uses the per-tile kernels),
// otherwise the launch status.
int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  static const int mode = [] {   // 0: off, 1: on (default); bit 1: transposed 5x5 convs as independent CTAs
    const char* e = getenv("B2C_TC_PERSIST");
    return e ? atoi(e) : 1;
  }();
  if (!mode) return 1;
  const int channels = L->channels ? L->channels : 128;
  if (channels != 128 || L->saved_y) return 1;
  if (L->kind == B2C_TC_CONV5S2 && L->split3) {
    if (L->gdn == 1 && !L->out_f32_nchw && L->out_lo)
      return launch_p<B2C_TC_CONV5S2, true, 1, OUT_BF16, true>(L, stream, trace);
    if (L->gdn == 0 && L->out_f32_nchw) {
      if (L->lut) {
        B2C_REQUIRE(L->lut_half_range >= 1 && L->bit_params && L->z_nhwc_out && L->bits_partial,
                    "tc_layer_run: the fused quantizer / entropy epilogue needs lut_half_range, bit_params, z_nhwc_out, "
                    "bits_partial");
        for (int i = 0; i < B2C_NUM_BITPARAMS; ++i)
          B2C_REQUIRE(L->bit_params[i], "tc_layer_run: null Bitparm pointer %d", i);
      }
      return launch_p<B2C_TC_CONV5S2, true, 0, OUT_F32_NCHW, true>(L, stream, trace);
    }
    return 1;
  }
  if (L->kind == B2C_TC_DECONV5S2 && !L->split3 && L->gdn == 2 && !L->out_f32_nchw && !L->out_lo) {
    if (mode & 2) return launch_p<B2C_TC_DECONV5S2, false, 2, OUT_BF16, false>(L, stream, trace);
    // B2C_TCP_A_RES=1: six resident A boxes per spatial tile + a 7-deep weight ring; 0: a streaming ring of 4 A boxes
    // (one per phase, x shift and k-block) + a 12-deep weight ring
    static const int a_res = [] { const char* e = getenv("B2C_TCP_A_RES"); return e ? atoi(e) : 1; }();
    if (a_res) return launch_p<B2C_TC_DECONV5S2, false, 2, OUT_BF16, true, true>(L, stream, trace);
    return launch_p<B2C_TC_DECONV5S2, false, 2, OUT_BF16, true>(L, stream, trace);
  }
  if (L->kind == B2C_TC_DECONV3 && !L->split3 && L->gdn == 0 && L->out_f32_nchw && !(mode & 4))
    return launch_p<B2C_TC_DECONV3, false, 0, OUT_F32_NCHW_PS4, false>(L, stream, trace);
  return 1;
}

}  // namespace tcp
}  // namespace b2c
