// Persistent tcgen05 kernel for the first analysis layer, role-specialised form of b2c_conv1_ts.cu: conv 9x9 / 4,
// 3 -> 128 (+bias) fused with GDN1 (models/analysis_17.py:14,17,33; models/GDN.py:64-94), bf16x3 arithmetic, NHWC bf16
// (hi, lo) output, both A operands of the MMAs in TMEM.
//
// In b2c_conv1_ts.cu the same eight warps build the im2col operand AND run the epilogue, phase after phase; every phase
// hands over to the MMA warp and waits for it, so the per-tile chain (build 2.3 k clk, x^2 1.1 k, final pass 2.9 k, waits
// 1.5 k) is serial and no pipe is ever more than 30 % busy (ncu: issue slots 37 %, LSU 30 %, MUFU 13 %, tensor 52 %).
// Here the two jobs belong to different warps that only meet through the MMA warp:
//   * 4 BUILDER warps (one per TMEM lane quadrant, thread = output pixel): fp32 patch -> bf16 (hi | lo) im2col rows in
//     registers -> tcgen05.st into a two-slot ring of TMEM columns, one 64-wide k-block per slot; they run ahead of
//     the epilogue by as much as the ring and the two patch slots allow;
//   * 8 EPILOGUE warps (thread = (pixel, 32-channel half)): x^2 (hi | lo) written over the accumulator columns it was
//     read from, then y = x * rsqrt(beta + gamma x^2) through per-warp staging slices and per-warp TMA stores;
//   * the MMA warp issues whatever is ready - conv k-blocks of tile i, gamma rounds of tile i-1 - polling both kinds
//     of barrier without blocking (a blocked wait on one would stall the tensor pipe while the other is ready).
// Shared-memory loads of the builders thus overlap the TMEM loads / MUFU / staging of the epilogue warps, and the
// per-tile time is the longest of the three roles instead of their sum.
// TMEM (512 columns): conv accumulator x 2 | norm accumulator x 1 | A ring 2 x 64 (hi words 0-31, lo words 32-63).
//
// TRAIN = true is the forward launch of the bf16 training step (train_tc.step_forward): plain bf16 operands (one MMA
// per product instead of three, no lo words anywhere) and the second output plane carries the norm n = beta + gamma x^2
// the GDN backward needs instead of the lo half of y.  The per-tile kernel that ran this layer before built its
// im2col tile in shared memory and ran prologue, main loop and epilogue of a tile one after the other: 18.6 us for the
// 256 tiles of a batch-8 step (two waves), 9 % tensor pipe.
#include <mutex>
#include <type_traits>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"
#include "b2c_tc_host.cuh"
#include "b2c_tcp_util.cuh"
#include "b2c_tc_ts.cuh"

namespace b2c {
namespace tcp1rs {

using namespace b2c::tc;

constexpr int C = 128;
constexpr int NCB = 2;                               // gamma k-blocks
constexpr int NKB = 4;                               // conv k-blocks (K = 243 -> 256)
constexpr int NP = 2;                                // patch slots
constexpr int NEPI_W = 8, NBLD_W = 4;                // epilogue / builder warps
constexpr int NTHREADS = 64 + 32 * (NEPI_W + NBLD_W);   // 448
constexpr int BLD_WARP0 = 2 + NEPI_W;                // first builder warp
constexpr uint32_t B_BYTES = (C / 2) * 128;          // this CTA's half of a 128-row weight / gamma tile
constexpr uint32_t W_RES = NKB * 2 * B_BYTES;        // 64 KB
constexpr uint32_t GAMMA_BYTES = NCB * 2 * B_BYTES;  // 32 KB
constexpr uint32_t STG_PLANE = 32 * 64;              // per-warp staging: 32 pixels x 32 channels of one plane
constexpr uint32_t STG_WARP = 2 /*blocks*/ * 2 /*hi, lo*/ * STG_PLANE;
constexpr uint32_t STG_BYTES = NEPI_W * STG_WARP;    // 64 KB
constexpr uint32_t SMEM = W_RES + GAMMA_BYTES + STG_BYTES + NP * P1_SLOT + 1024 /*align*/ + 1024 /*bias, beta*/ +
                          512 /*barriers*/;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t T_ACC = 0, T_NORM = 256, T_A = 384, T_COLS = 512;

struct Params {
  int B, Hg, Wg;
  int tiles_x, tiles_per_image, n_tiles;
  const float* bias;
  const float* beta_eff;
  long long* trace;      // debug: clock64 timeline of CTA 0: [4 roles: producer, mma, epilogue, builder][64 marks]
};

template <bool TRAIN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv1_gdn_rs_kernel(const __grid_constant__ CUtensorMap tmImg, const __grid_constant__ CUtensorMap tmB_hi,
                    const __grid_constant__ CUtensorMap tmB_lo, const __grid_constant__ CUtensorMap tmG_hi,
                    const __grid_constant__ CUtensorMap tmG_lo, const __grid_constant__ CUtensorMap tmO_hi,
                    const __grid_constant__ CUtensorMap tmO_lo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wres = smem;                                // [hi kb0..3][lo kb0..3], 64 rows each
  uint8_t* gam = wres + W_RES;                         // [hi kb0][hi kb1][lo kb0][lo kb1]
  uint8_t* stg = gam + GAMMA_BYTES;                    // per-warp output staging (1024-byte aligned slices)
  float* patch = reinterpret_cast<float*>(stg + STG_BYTES);            // NP slots of P1_SLOT bytes
  float* s_bias = reinterpret_cast<float*>(stg + STG_BYTES + NP * P1_SLOT);
  float* s_beta = s_bias + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + 128);
  uint64_t* a_full = bars;                     // [2] (leader's: one arrival per builder warp of the pair)
  uint64_t* a_empty = a_full + 2;              // [2] (multicast commit)
  uint64_t* accum_full = a_empty + 2;          // [2] (multicast commit)
  uint64_t* norm_full = accum_full + 2;        // [1] (multicast commit)
  uint64_t* acc_empty = norm_full + 1;         // [2] (leader's: one arrival per epilogue warp of the pair)
  uint64_t* xs_full = acc_empty + 2;           // [2] (leader's): x^2 round g of the current epilogue tile is in TMEM
  uint64_t* res_full = xs_full + 2;            // resident weights + gamma landed (leader's)
  uint64_t* patch_full = res_full + 1;         // [NP] local
  uint64_t* patch_empty = patch_full + NP;     // [NP] local: one arrival per builder warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(patch_empty + NP);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_units = (p.n_tiles + 1) / 2;
  const int unit0 = (int)(blockIdx.x / 2), unit_stride = (int)(gridDim.x / 2);
  const int n_it = (n_units - unit0 + unit_stride - 1) / unit_stride;     // >= 1 (grid <= 2 * n_units)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  int tmark = 0;
#define T1_MARK(role, cond)                                                          \
  do {                                                                               \
    if (tracing && (cond) && tmark < 64) p.trace[(role) * 64 + tmark++] = clock64(); \
  } while (0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 2 * NBLD_W);
      mbar_init(&a_empty[i], 1);
      mbar_init(&accum_full[i], 1);
      mbar_init(&acc_empty[i], 2 * NEPI_W);
      mbar_init(&xs_full[i], 2 * NEPI_W);
    }
    mbar_init(norm_full, 1);
    mbar_init(res_full, 1);
    for (int i = 0; i < NP; ++i) { mbar_init(&patch_full[i], 1); mbar_init(&patch_empty[i], NBLD_W); }
    fence_barrier_init();
    tma_prefetch_desc(&tmImg);
    tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    tma_prefetch_desc(&tmG_hi); tma_prefetch_desc(&tmG_lo);
    tma_prefetch_desc(&tmO_hi); tma_prefetch_desc(&tmO_lo);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, T_COLS);
  if (TRAIN) griddep_wait();     // training: beta_eff (and the packed operands) come from the launch right in front of this one
  if (threadIdx.x >= 64 && threadIdx.x < 64 + C) {
    const int c = threadIdx.x - 64;
    s_bias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
    s_beta[c] = p.beta_eff ? __ldg(p.beta_eff + c) : 1.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();       // the next layer's prologue may overlap this kernel's tail

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      if (TRAIN) griddep_wait();  // training: the operands were packed by the launch in front of this one
      if (leader) mbar_arrive_expect_tx(res_full, (TRAIN ? 1u : 2u) * (W_RES + GAMMA_BYTES));
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        tma_load_2d_pair(wres + kb * B_BYTES, &tmB_hi, res_full, 0, kb * C + (int)cta_rank * (C / 2));
        if (!TRAIN) tma_load_2d_pair(wres + (NKB + kb) * B_BYTES, &tmB_lo, res_full, 0, kb * C + (int)cta_rank * (C / 2));
      }
#pragma unroll
      for (int g = 0; g < NCB; ++g) {
        tma_load_2d_pair(gam + g * B_BYTES, &tmG_hi, res_full, g * KBLK, (int)cta_rank * (C / 2));
        if (!TRAIN) tma_load_2d_pair(gam + (NCB + g) * B_BYTES, &tmG_lo, res_full, g * KBLK, (int)cta_rank * (C / 2));
      }
      griddep_wait();             // the image (and the buffers this kernel overwrites) are free from here on
      for (int it = 0; it < n_it; ++it) {
        const int tile = (unit0 + it * unit_stride) * 2 + (int)cta_rank;
        const int n_img = tile / p.tiles_per_image;          // == B for the masked tile of an odd count: all OOB (zeros)
        const int trem = tile - n_img * p.tiles_per_image;
        const int tile_y = trem / p.tiles_x, tile_x = trem - tile_y * p.tiles_x;
        const int ps = it % NP;
        mbar_wait(&patch_empty[ps], (uint32_t)((it / NP) & 1) ^ 1u);
        T1_MARK(0, true);
        mbar_arrive_expect_tx(&patch_full[ps], (uint32_t)P1_BYTES);
        tma_load_4d(reinterpret_cast<uint8_t*>(patch) + ps * P1_SLOT, &tmImg, &patch_full[ps], tile_x * TW * 4 - 4,
                    tile_y * TH * 4 - 4, 0, n_img);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (whole warp, one elected lane issues)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16_m256(C);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint64_t wres_d = umma_desc_sw128(sbase);
      const uint64_t gam_d = wres_d + (W_RES >> 4);

      // conv k-block kb (global count kcnt) of the tile accumulated into set `buf`
      auto conv_kb = [&](int kb, uint32_t buf) {
        tcgen05_fence_after();
        T1_MARK(1, lane == 0);
        if (elect_one()) {
          const uint32_t d = tb + T_ACC + buf * 128u;
          const int slot = kb & 1;
          const uint32_t a0 = tb + T_A + (uint32_t)slot * 64u;
          const uint64_t b_hi = wres_d + ((uint32_t)(kb * B_BYTES) >> 4);
          const uint64_t b_lo = wres_d + ((uint32_t)((NKB + kb) * B_BYTES) >> 4);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
            const uint32_t a_hi = a0 + (uint32_t)(k4 * 8), a_lo = a_hi + 32u;
            umma_bf16_pair_ts(d, a_hi, b_hi + 2 * k4, idesc, acc);
            if (!TRAIN) {
              umma_bf16_pair_ts(d, a_hi, b_lo + 2 * k4, idesc, 1u);
              umma_bf16_pair_ts(d, a_lo, b_hi + 2 * k4, idesc, 1u);
            }
          }
          umma_commit_pair(&a_empty[slot]);
          if (kb == NKB - 1) umma_commit_pair(&accum_full[buf]);
        }
        __syncwarp();
      };
      // gamma k-block `round` of the tile in set `ebuf`: x^2 (hi | lo) of epilogue half h sits in the accumulator's own
      // columns round*64 + 32h: 16 hi words then 16 lo words
      auto gamma_round = [&](int round, uint32_t ebuf) {
        tcgen05_fence_after();
        T1_MARK(1, lane == 0);
        if (elect_one()) {
          const uint32_t d = tb + T_NORM;
          const uint32_t a0 = tb + T_ACC + ebuf * 128u + (uint32_t)round * 64u;
          const uint64_t b_hi = gam_d + ((uint32_t)(round * B_BYTES) >> 4);
          const uint64_t b_lo = gam_d + ((uint32_t)((NCB + round) * B_BYTES) >> 4);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acc = (round > 0 || k4 > 0) ? 1u : 0u;
            const uint32_t a_hi = a0 + (uint32_t)((k4 >> 1) * 32 + (k4 & 1) * 8), a_lo = a_hi + 16u;
            umma_bf16_pair_ts(d, a_hi, b_hi + 2 * k4, idesc, acc);
            if (!TRAIN) {
              umma_bf16_pair_ts(d, a_hi, b_lo + 2 * k4, idesc, 1u);
              umma_bf16_pair_ts(d, a_lo, b_hi + 2 * k4, idesc, 1u);
            }
          }
          if (round == NCB - 1) umma_commit_pair(norm_full);
        }
        __syncwarp();
      };
      auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {
        const uint32_t ok = mbar_test_wait(bar, parity, true);
        return __shfl_sync(0xffffffffu, ok, 0) != 0;
      };

      mbar_wait(res_full, 0);
      // two in-order work streams, served as their barriers complete:
      //   conv: k-block ck of tile ct (ct < n_it)          gamma: round gr of tile gt (gt < n_it), needs accum_full(gt)
      int ct = 0, ck = 0, gt = 0, gr = 0;
      bool acc_ok = false;             // acc_empty of tile ct's accumulator set has been observed
      long long t0 = clock64();
      // The tensor pipe executes in issue order and the epilogue warps are the critical chain (x^2 -> gamma MMAs ->
      // final pass): once tile gt's accumulator is complete, at most ONE conv k-block of the next tile (the pre-built
      // one, covering the x^2 computation) is issued before tile gt's gamma rounds - otherwise they queue behind up
      // to 3 k clk of convolution MMAs that nobody is waiting for yet (measured: norm wait 3.5 k clk per tile).
      constexpr int conv_credit = 1;
      while (ct < n_it || gt < n_it) {
        bool progressed = false;
        if (ct < n_it && !(gt < ct && ck >= conv_credit)) {
          const uint32_t buf = (uint32_t)ct & 1u;
          if (!acc_ok && ck == 0) acc_ok = ready(&acc_empty[buf], (((uint32_t)ct >> 1) & 1u) ^ 1u);
          if ((acc_ok || ck > 0) && ready(&a_full[ck & 1], (uint32_t)((ck >> 1) & 1))) {
            conv_kb(ck, buf);
            if (++ck == NKB) { ck = 0; ++ct; acc_ok = false; }
            progressed = true;
          }
        }
        // gamma rounds of tile gt come after ALL conv k-blocks of tile gt were issued (gt < ct)
        if (gt < ct && ready(&xs_full[gr], (uint32_t)gt & 1u)) {
          gamma_round(gr, (uint32_t)gt & 1u);
          if (++gr == NCB) { gr = 0; ++gt; }
          progressed = true;
        }
        if (progressed) {
          t0 = clock64();
        } else if (clock64() - t0 > 4000000000LL) {      // ~2 s without progress: a protocol bug, not a hung GPU
          if (lane == 0) printf("b2c: conv1 rs MMA warp stalled (block %d ct %d ck %d gt %d gr %d)\n", blockIdx.x, ct, ck, gt, gr);
          __trap();
        }
      }
    }
  } else if (warp >= BLD_WARP0) {
    // ================================================================ builders (4 warps): thread = output pixel r
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ty = r / TW, tx = r - ty * TW;
    const uint32_t t_ring = tmem_base + ((uint32_t)(q * 32) << 16) + T_A;
    const int pp_off = (4 * ty) * P1_W + 4 * tx;
    const bool mark = warp == BLD_WARP0 && lane == 0;
    griddep_wait();
    uint32_t use = 0;                          // k-blocks built so far: slot = use & 1, its round = use >> 1
    for (int it = 0; it < n_it; ++it) {
      const int ps = it % NP;
      const float* pp = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(patch) + ps * P1_SLOT) + pp_off;
      mbar_wait(&patch_full[ps], (uint32_t)((it / NP) & 1));
      T1_MARK(3, mark);
      auto build = [&](auto kbc) {
        constexpr int KB = decltype(kbc)::value;
        const uint32_t slot = use & 1u;
        mbar_wait(&a_empty[slot], ((use >> 1) & 1u) ^ 1u);     // the MMAs of the slot's previous use have read it
        tcgen05_fence_after();
        uint32_t w[32];
        conv1_build_regs<KB, 0, 32>(pp, w);                    // k 0..31 of the block: w[0..15] hi, w[16..31] lo
        tmem_st16(t_ring + slot * 64u, *reinterpret_cast<uint32_t(*)[16]>(&w[0]));
        if (!TRAIN) tmem_st16(t_ring + slot * 64u + 32u, *reinterpret_cast<uint32_t(*)[16]>(&w[16]));
        conv1_build_regs<KB, 1, 32>(pp, w);                    // k 32..63
        tmem_st16(t_ring + slot * 64u + 16u, *reinterpret_cast<uint32_t(*)[16]>(&w[0]));
        if (!TRAIN) tmem_st16(t_ring + slot * 64u + 48u, *reinterpret_cast<uint32_t(*)[16]>(&w[16]));
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta0(&a_full[slot]);
        ++use;
      };
      build(std::integral_constant<int, 0>{});
      build(std::integral_constant<int, 1>{});
      build(std::integral_constant<int, 2>{});
      build(std::integral_constant<int, 3>{});
      __syncwarp();
      if (lane == 0) mbar_arrive(&patch_empty[ps]);            // this warp is done with the patch
      T1_MARK(3, mark);
    }
  } else {
    // ================================================================ epilogue (8 warps): thread = (pixel r, half)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool mark = threadIdx.x == 64;
    uint8_t* my_stg = stg + (uint32_t)(warp - 2) * STG_WARP;
    const uint32_t st_row = (uint32_t)lane * 64u, st_x = ((uint32_t)lane >> 1) & 3u;
    int stores = 0;                            // lane 0: bulk store groups committed so far
    griddep_wait();

    for (int e = 0; e < n_it; ++e) {
      const uint32_t ebuf = (uint32_t)e & 1u, epar = ((uint32_t)e >> 1) & 1u;
      const uint32_t t_acc = t_lane + T_ACC + ebuf * 128u + (uint32_t)(half * 32), t_nrm = t_lane + T_NORM + (uint32_t)(half * 32);
      const int tile = (unit0 + e * unit_stride) * 2 + (int)cta_rank;
      const int e_img = tile / p.tiles_per_image;
      const int trem = tile - e_img * p.tiles_per_image;
      const int tile_y = trem / p.tiles_x;
      const int e_oy0 = tile_y * TH, e_ox0 = (trem - tile_y * p.tiles_x) * TW;
      uint64_t xk[32];                          // x = acc + bias: xk[g*16 + i] = channels g*64 + half*32 + 2i, +1

      mbar_wait(&accum_full[ebuf], epar);
      tcgen05_fence_after();
      T1_MARK(2, mark);                          // [0] conv accumulator ready
      // ---- x^2 (hi | lo) of both 64-channel rounds, stored over the columns they came from (all 32 columns of a round
      // are in registers before the first store; the stores go out in 8-column pieces to keep the register peak low)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t raw[32];
        tmem_ld32(t_acc + (uint32_t)(g * 64), raw);
        tmem_ld_wait();
        const uint64_t* bias2 = reinterpret_cast<const uint64_t*>(s_bias + g * 64 + half * 32);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t wh[8], wl[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int pi = hh * 8 + i;
            const uint64_t x2 = add2(pk2(raw[2 * pi], raw[2 * pi + 1]), bias2[pi]);
            xk[g * 16 + pi] = x2;
            float v0, v1;
            upk2(mul2(x2, x2), v0, v1);
            wh[i] = cvt_bf16x2(v0, v1);
            if (!TRAIN) wl[i] = lo_of(v0, v1, wh[i]);
          }
          tmem_st8(t_acc + (uint32_t)(g * 64 + hh * 8), wh);
          if (!TRAIN) tmem_st8(t_acc + (uint32_t)(g * 64 + 16 + hh * 8), wl);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta0(&xs_full[g]);
      }
      T1_MARK(2, mark);                          // [1] x^2 written
      // ---- final pass: y = x * rsqrt(beta + norm) -> NHWC bf16 (hi, lo); 16-column chunks, the TMEM load of chunk
      // j+1 in flight while chunk j is computed and staged; per-warp TMA stores per 64-channel block
      mbar_wait(norm_full, (uint32_t)e & 1u);
      tcgen05_fence_after();
      T1_MARK(2, mark);                          // [2] norm ready
      uint32_t rawn[2][16];
      tmem_ld16(t_nrm, rawn[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {              // chunk j: channels (j >> 1) * 64 + half * 32 + (j & 1) * 16 .. +15
        const int blk = j >> 1, sub = j & 1;
        tmem_ld_wait();
        if (j < 3) tmem_ld16(t_nrm + (uint32_t)(((j + 1) >> 1) * 64 + ((j + 1) & 1) * 16), rawn[(j + 1) & 1]);
        const uint64_t* beta2 = reinterpret_cast<const uint64_t*>(s_beta + blk * 64 + half * 32 + sub * 16);
        uint32_t hp[8], lp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float n0, n1, y0, y1;
          upk2(add2(pk2(rawn[j & 1][2 * i], rawn[j & 1][2 * i + 1]), beta2[i]), n0, n1);
          const uint64_t sc = pkf2(rsqrt_fast(n0), rsqrt_fast(n1));
          upk2(mul2(xk[blk * 16 + sub * 8 + i], sc), y0, y1);
          hp[i] = cvt_bf16x2(y0, y1);
          lp[i] = TRAIN ? cvt_bf16x2(n0, n1) : lo_of(y0, y1, hp[i]);      // training: the second plane is the norm
        }
        uint8_t* sb = my_stg + (uint32_t)blk * (2 * STG_PLANE);
        if (sub == 0) {                          // the stores that last read this block's slice (previous tile) are done
          if (lane == 0 && stores >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        const uint32_t k0 = (uint32_t)sub * 2u;
        *reinterpret_cast<uint4*>(sb + st_row + (((k0) ^ st_x) << 4)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(sb + st_row + (((k0 + 1) ^ st_x) << 4)) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
        *reinterpret_cast<uint4*>(sb + STG_PLANE + st_row + (((k0) ^ st_x) << 4)) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        *reinterpret_cast<uint4*>(sb + STG_PLANE + st_row + (((k0 + 1) ^ st_x) << 4)) = make_uint4(lp[4], lp[5], lp[6], lp[7]);
        if (sub == 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int cc = blk * 64 + half * 32;
            tma_store_4d(&tmO_hi, sb, cc, e_ox0, e_oy0 + 2 * q, e_img);
            tma_store_4d(&tmO_lo, sb + STG_PLANE, cc, e_ox0, e_oy0 + 2 * q, e_img);
            tma_store_commit();
            ++stores;
          }
        }
      }
      tcgen05_fence_before();                    // my TMEM reads of this accumulator set / the norm are complete
      __syncwarp();
      if (lane == 0) mbar_arrive_cta0(&acc_empty[ebuf]);
      T1_MARK(2, mark);                          // [3] staged + stores issued
    }
    if (lane == 0 && stores > 0) tma_store_wait_read();      // the staging slices must outlive their stores
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, T_COLS);
  }
#undef T1_MARK
}

// ------------------------------------------------------------------------------------ host side
// Returns 1 when the layer is not the one this kernel covers or another form is selected (the caller then tries the
// others), otherwise the launch status.  B2C_TC_PERSIST_CONV1: 3 (default) = this kernel, 2 = b2c_conv1_ts.cu,
// 1 = shared-memory operands (b2c_conv1_tcp.cu), 0 = per-tile kernel.
template <bool TRAIN>
static int launch_rs(const b2c_tc_layer* L, cudaStream_t stream, long long* trace);

int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  static const int mode = [] {
    const char* e = getenv("B2C_TC_PERSIST_CONV1");
    return e ? atoi(e) : 3;
  }();
  if (mode != 3) return 1;
  const int channels = L->channels ? L->channels : 128;
  if (L->kind != B2C_TC_CONV1 || channels != 128 || L->gdn != 1 || L->out_f32_nchw || !L->out_lo || L->saved_y) return 1;
  if (L->split3 && (!L->w_lo || !L->gamma_lo)) return 1;
  // plain bf16 + a second output plane = the forward launch of the training step (B2C_TRAIN_CONV1_RS=0: per-tile kernel)
  static const int train_rs = [] { const char* e = getenv("B2C_TRAIN_CONV1_RS"); return e ? atoi(e) : 1; }();
  if (!L->split3 && !train_rs) return 1;
  return L->split3 ? launch_rs<false>(L, stream, trace) : launch_rs<true>(L, stream, trace);
}

template <bool TRAIN>
static int launch_rs(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  auto kern = conv1_gdn_rs_kernel<TRAIN>;

  Params p;
  memset(&p, 0, sizeof(p));
  p.B = L->B;
  p.Hg = L->Hin / 4; p.Wg = L->Win / 4;
  p.tiles_x = (p.Wg + TW - 1) / TW;
  p.tiles_per_image = p.tiles_x * ((p.Hg + TH - 1) / TH);
  p.n_tiles = p.B * p.tiles_per_image;
  p.bias = L->bias; p.beta_eff = L->beta_eff;
  p.trace = trace;

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
      int n = 2 * ncl;
      if (n > (sms & ~1)) n = sms & ~1;
      if (const char* env = getenv("B2C_TCP_MAX_CTAS")) { const int v = atoi(env); if (v > 0 && v < n) n = v & ~1; }
      if (n <= 0) { set_error("conv1 rs: no resident CTA pairs"); return B2C_ECUDA; }
      if (dev >= 0 && dev < 64) max_ctas[dev] = n;
      resident = n;
    } else {
      resident = max_ctas[dev];
    }
  }
  const int n_units = (p.n_tiles + 1) / 2;
  int grid = n_units * 2;
  if (grid > resident) grid = resident;

  CUtensorMap mImg, mB_hi, mB_lo, mG_hi, mG_lo, mO_hi, mO_lo;
  int rc;
  if ((rc = make_image_map(&mImg, L->in_hi, L->B, L->Hin, L->Win))) return rc;
  if ((rc = make_mat_map(&mB_hi, L->w_hi, NKB * C, KBLK, C / 2))) return rc;
  if ((rc = make_mat_map(&mB_lo, TRAIN ? L->w_hi : L->w_lo, NKB * C, KBLK, C / 2))) return rc;
  if ((rc = make_mat_map(&mG_hi, L->gamma_hi, C, C, C / 2))) return rc;
  if ((rc = make_mat_map(&mG_lo, TRAIN ? L->gamma_hi : L->gamma_lo, C, C, C / 2))) return rc;
  if ((rc = make_warp_out_map(&mO_hi, L->out_hi, L->B, p.Hg, p.Wg, 32))) return rc;
  if ((rc = make_warp_out_map(&mO_lo, L->out_lo, L->B, p.Hg, p.Wg, 32))) return rc;

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mImg, mB_hi, mB_lo, mG_hi, mG_lo, mO_hi, mO_lo, p);
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(conv1 rs): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  count_launch();
  return check_launch("tc_layer_run(persistent conv1, role-specialised, operands in TMEM)");
}

}  // namespace tcp1rs
}  // namespace b2c
