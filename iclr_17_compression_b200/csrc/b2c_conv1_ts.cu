// Persistent tcgen05 kernel for the first analysis layer, A-operands-in-TMEM form: conv 9x9 / 4, 3 -> 128 (+bias) fused
// with GDN1 (models/analysis_17.py:14,17,33; models/GDN.py:64-94), bf16x3 arithmetic, NHWC bf16 (hi, lo) output.
//
// Same tiling and CTA-pair structure as b2c_conv1_tcp.cu (one cta_group::2 pair per two SMs, a static round-robin list
// of pairs of adjacent 8x16 output tiles, weights and gamma resident in shared memory), but neither A operand ever
// touches shared memory.  ncu of the shared-memory form: LSU wavefronts 43 % + operand reads of the MMAs 27 % of the
// SM's 128 B/clk data path, tensor pipe 37 % - the kernel was bound by shared-memory bandwidth, and an SS-mode
// M128xN128xK16 MMA alone reads 6 KB per 64 clk.  Here
//   * the im2col operand of the convolution is built by the 256 row workers in registers (fp32 patch -> bf16 hi | lo)
//     and stored with tcgen05.st into a two-slot ring of TMEM columns (A of a tcgen05.mma may come from TMEM: row =
//     lane, two bf16 per 32-bit column), 256 B/clk on a path of its own instead of 128 B/clk shared with everything;
//   * the x^2 operand of the gamma product is written back into the very accumulator columns it was computed from
//     (a thread overwrites only columns it has itself read; x stays in registers for the final multiply);
//   * the K order inside a 64-wide k-block is free as long as both operands agree, so a worker (row r, half h) owns
//     the 32 columns [32h, 32h+32) of a slot: words 0-15 = hi of its k pairs, words 16-31 = lo - ONE tcgen05.st.x32.
//   * with no operand or output tile in shared memory the eight worker warps never have to meet: every hand-off to
//     the MMA warp is a per-warp mbarrier arrival (16 per phase over the pair), the output leaves through PER-WARP
//     staging slices and TMA stores (a warp owns 2 tile rows x 16 pixels x 32 channels of each 64-channel block: a
//     {32 ch, 16, 2} box, SWIZZLE_64B), and the warps drift out of phase - shared-memory loads of one warp's build
//     overlap TMEM loads / MUFU work of another's epilogue (the lock-step form spent 4 k of 9.2 k clk per tile in a
//     final pass whose pipes ran one after the other; plain STG.128 from registers, one 16-byte piece per pixel per
//     instruction, was worse still: 32 sectors per warp store, 4.8 k clk).
// TMEM (512 columns): conv accumulator x 2 (double-buffered across tiles) | norm accumulator x 1 | A ring 2 x 64.
// Shared memory per CTA: weights 64 KB + gamma 32 KB (its half of every tile) + 2 patch slots + 8 x 8 KB staging.
// Per-warp order per iteration `it` (B = build tile it, E = epilogue tile it-1; `|` = tcgen05 fence + warp arrival):
//                            B.kb0 B.kb1 | E.x^2[0] | B.kb2 E.x^2[1] | B.kb3 | E.final
// MMA-warp order:            conv kb0  conv kb1  gamma[0]  gamma[1]  conv kb2  conv kb3
#include <mutex>
#include <type_traits>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"
#include "b2c_tc_host.cuh"
#include "b2c_tcp_util.cuh"
#include "b2c_tc_ts.cuh"

namespace b2c {
namespace tcp1ts {

using namespace b2c::tc;

constexpr int C = 128;
constexpr int NCB = 2;                               // gamma k-blocks
constexpr int NKB = 4;                               // conv k-blocks (K = 243 -> 256)
constexpr int NP = 2;                                // patch slots (the producer runs one tile ahead)
constexpr uint32_t B_BYTES = (C / 2) * 128;          // this CTA's half of a 128-row weight / gamma tile
constexpr uint32_t W_RES = NKB * 2 * B_BYTES;        // 64 KB
constexpr uint32_t GAMMA_BYTES = NCB * 2 * B_BYTES;  // 32 KB
// NPART worker warps share a TMEM lane quadrant (32 pixel rows), each owning CHP = 64 / NPART of the 64 channels /
// k values of a block.  NPART = 2 (8 worker warps) is the default; NPART = 4 (16 worker warps, 4 per scheduler) was
// tried for latency hiding and measured slower (0.171 vs 0.158 ms per 24 images).
template <int NPART>
struct Cfg {
  static_assert(NPART == 2 || NPART == 4, "2 or 4 column parts");
  static constexpr int NWORKERS = 128 * NPART;
  static constexpr int NTHREADS = 64 + NWORKERS;
  static constexpr int NWARPS_W = 4 * NPART;                 // worker warps per CTA
  static constexpr int CHP = 64 / NPART;                     // channels (= k values, = TMEM columns) per thread and block
  static constexpr int NW = CHP / 2;                         // packed bf16x2 words per plane
  static constexpr uint32_t STG_PLANE = 32 * CHP * 2;        // per-warp staging: 32 pixels x CHP channels of one plane
  static constexpr uint32_t STG_WARP = 2 /*blocks*/ * 2 /*hi, lo*/ * STG_PLANE;
};
constexpr uint32_t STG_BYTES = 64 * 1024;            // = NWARPS_W * STG_WARP for either NPART
constexpr uint32_t SMEM = W_RES + GAMMA_BYTES + STG_BYTES + NP * P1_SLOT + 1024 /*align*/ + 1024 /*bias, beta*/ +
                          512 /*barriers*/;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
// TMEM columns
constexpr uint32_t T_ACC = 0;                        // 2 x 128: conv accumulators; x^2 (hi | lo) aliases them
constexpr uint32_t T_NORM = 256;                     // 128: beta-less norm = gamma * x^2
constexpr uint32_t T_A = 384;                        // 2 x 64: im2col k-block ring (per half: 16 hi + 16 lo words)
constexpr uint32_t T_COLS = 512;

struct Params {
  int B, Hg, Wg;
  int tiles_x, tiles_per_image, n_tiles;
  const float* bias;
  const float* beta_eff;
  long long* trace;      // debug: clock64 timeline of CTA 0: [3 roles][64 marks]
};


template <int KB, int NPART>
__device__ __forceinline__ void conv1_build_part(const float* __restrict__ pp, int part, uint32_t (&w)[64 / NPART]) {
  if (part == 0) conv1_build_regs<KB, 0, 64 / NPART>(pp, w);
  else if (part == 1) conv1_build_regs<KB, 1, 64 / NPART>(pp, w);
  else if (NPART == 4 && part == 2) conv1_build_regs<KB, 2 % NPART, 64 / NPART>(pp, w);
  else conv1_build_regs<KB, 3 % NPART, 64 / NPART>(pp, w);
}

// (NPART = 4: 18 warps are allocated as 20 - warp allocation granularity 4 - so the register cap is 96 per thread;
// ptxas spills ~150 bytes per thread, all in the cold paths.  112 x 576 threads would fit the file but not the allocator.)
template <int NPART>
__global__ void __launch_bounds__(Cfg<NPART>::NTHREADS, 1)
conv1_gdn_ts_kernel(const __grid_constant__ CUtensorMap tmImg, const __grid_constant__ CUtensorMap tmB_hi,
                    const __grid_constant__ CUtensorMap tmB_lo, const __grid_constant__ CUtensorMap tmG_hi,
                    const __grid_constant__ CUtensorMap tmG_lo, const __grid_constant__ CUtensorMap tmO_hi,
                    const __grid_constant__ CUtensorMap tmO_lo, const Params p) {
  using K = Cfg<NPART>;
  constexpr int NWARPS_W = K::NWARPS_W, CHP = K::CHP, NW = K::NW;
  constexpr uint32_t STG_PLANE = K::STG_PLANE, STG_WARP = K::STG_WARP;
  static_assert(NWARPS_W * STG_WARP == STG_BYTES, "staging budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wres = smem;                                // [hi kb0..3][lo kb0..3], 64 rows each
  uint8_t* gam = wres + W_RES;                         // [hi kb0][hi kb1][lo kb0][lo kb1]
  uint8_t* stg = gam + GAMMA_BYTES;                    // per-warp output staging (1024-byte aligned slices)
  float* patch = reinterpret_cast<float*>(stg + STG_BYTES);            // NP slots of P1_SLOT bytes
  float* s_bias = reinterpret_cast<float*>(stg + STG_BYTES + NP * P1_SLOT);
  float* s_beta = s_bias + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + 128);
  uint64_t* a_full = bars;                     // [2] (leader's: one arrival per worker warp of the pair)
  uint64_t* a_empty = a_full + 2;              // [2] (multicast commit)
  uint64_t* accum_full = a_empty + 2;          // [2] (multicast commit)
  uint64_t* norm_full = accum_full + 2;        // [1] (multicast commit)
  uint64_t* acc_empty = norm_full + 1;         // [2] (leader's)
  uint64_t* xs_full = acc_empty + 2;           // [2] (leader's): x^2 round g of the current epilogue tile is in TMEM
  uint64_t* res_full = xs_full + 2;            // resident weights + gamma landed (leader's)
  uint64_t* patch_full = res_full + 1;         // [NP] local
  uint64_t* patch_empty = patch_full + NP;     // [NP] local
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
      mbar_init(&a_full[i], 2 * NWARPS_W);
      mbar_init(&a_empty[i], 1);
      mbar_init(&accum_full[i], 1);
      mbar_init(&acc_empty[i], 2 * NWARPS_W);
      mbar_init(&xs_full[i], 2 * NWARPS_W);
    }
    mbar_init(norm_full, 1);
    mbar_init(res_full, 1);
    for (int i = 0; i < NP; ++i) { mbar_init(&patch_full[i], 1); mbar_init(&patch_empty[i], NWARPS_W); }
    fence_barrier_init();
    tma_prefetch_desc(&tmImg);
    tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    tma_prefetch_desc(&tmG_hi); tma_prefetch_desc(&tmG_lo);
    tma_prefetch_desc(&tmO_hi); tma_prefetch_desc(&tmO_lo);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, T_COLS);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + C) {   // (NWORKERS >= 128)
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
      if (leader) mbar_arrive_expect_tx(res_full, 2u * (W_RES + GAMMA_BYTES));
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        tma_load_2d_pair(wres + kb * B_BYTES, &tmB_hi, res_full, 0, kb * C + (int)cta_rank * (C / 2));
        tma_load_2d_pair(wres + (NKB + kb) * B_BYTES, &tmB_lo, res_full, 0, kb * C + (int)cta_rank * (C / 2));
      }
#pragma unroll
      for (int g = 0; g < NCB; ++g) {
        tma_load_2d_pair(gam + g * B_BYTES, &tmG_hi, res_full, g * KBLK, (int)cta_rank * (C / 2));
        tma_load_2d_pair(gam + (NCB + g) * B_BYTES, &tmG_lo, res_full, g * KBLK, (int)cta_rank * (C / 2));
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
      uint32_t tile_cnt = 0;                    // epilogue tiles seen so far (phase of the x^2 barriers)

      // conv k-block kb of the tile being accumulated into set `buf`; k16 step k4 lives in the columns of worker part
      // k4 / (4 / NPART): hi words at part * CHP + 8 * (k4 % (4 / NPART)), lo words CHP / 2 columns further
      constexpr int SPP = 4 / NPART;             // k16 steps per part
      auto conv_kb = [&](int kb, uint32_t buf, bool last) {
        const uint32_t d = tb + T_ACC + buf * 128u;
        const int slot = kb & 1;
        mbar_wait_cluster(&a_full[slot], (uint32_t)((kb >> 1) & 1));
        tcgen05_fence_after();
        T1_MARK(1, lane == 0);
        if (elect_one()) {
          const uint32_t a0 = tb + T_A + (uint32_t)slot * 64u;
          const uint64_t b_hi = wres_d + ((uint32_t)(kb * B_BYTES) >> 4);
          const uint64_t b_lo = wres_d + ((uint32_t)((NKB + kb) * B_BYTES) >> 4);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
            const uint32_t a_hi = a0 + (uint32_t)((k4 / SPP) * CHP + (k4 % SPP) * 8), a_lo = a_hi + (uint32_t)NW;
            umma_bf16_pair_ts(d, a_hi, b_hi + 2 * k4, idesc, acc);
            umma_bf16_pair_ts(d, a_hi, b_lo + 2 * k4, idesc, 1u);
            umma_bf16_pair_ts(d, a_lo, b_hi + 2 * k4, idesc, 1u);
          }
          umma_commit_pair(&a_empty[slot]);
          if (last) umma_commit_pair(&accum_full[buf]);
        }
        __syncwarp();
      };
      // gamma k-block `round` of the tile in set `ebuf`: norm (+)= x^2[:, round*64 ..] * gamma[:, round*64 ..]^T with the
      // x^2 operand in the accumulator's own columns [round*64, round*64 + 64)
      auto gamma_round = [&](int round, uint32_t ebuf) {
        mbar_wait_cluster(&xs_full[round], tile_cnt & 1u);
        if (round == NCB - 1) ++tile_cnt;
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
            const uint32_t a_hi = a0 + (uint32_t)((k4 / SPP) * CHP + (k4 % SPP) * 8), a_lo = a_hi + (uint32_t)NW;
            umma_bf16_pair_ts(d, a_hi, b_hi + 2 * k4, idesc, acc);
            umma_bf16_pair_ts(d, a_hi, b_lo + 2 * k4, idesc, 1u);
            umma_bf16_pair_ts(d, a_lo, b_hi + 2 * k4, idesc, 1u);
          }
          if (round == NCB - 1) umma_commit_pair(norm_full);
        }
        __syncwarp();
      };

      mbar_wait(res_full, 0);
      for (int it = 0; it <= n_it; ++it) {
        const bool B_ = it < n_it, E_ = it >= 1;
        const uint32_t buf = (uint32_t)it & 1u, ebuf = buf ^ 1u;
        if (B_) {
          mbar_wait_cluster(&acc_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);   // set drained by the workers (tile it-2)
          tcgen05_fence_after();
          conv_kb(0, buf, false);
          conv_kb(1, buf, false);
        }
        if (E_) gamma_round(0, ebuf);
        if (E_) gamma_round(1, ebuf);
        if (B_) conv_kb(2, buf, false);
        if (B_) conv_kb(3, buf, true);
      }
    }
  } else {
    // ================================================================ row workers: thread = (pixel r, column part)
    // No block-level barrier below: a warp's hand-offs are its own (tcgen05 fence, __syncwarp, one mbarrier arrival).
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ty = r / TW, tx = r - ty * TW;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_ring = t_lane + T_A + (uint32_t)(part * CHP);
    const bool mark = threadIdx.x == 64;
    const int pp_off = (4 * ty) * P1_W + 4 * tx;
    int e_img = 0, e_ox0 = 0, e_oy0 = 0;       // tile whose epilogue runs in the current iteration
    // this warp's staging slice: [block 0 hi | block 0 lo | block 1 hi | block 1 lo], rows of CHP * 2 bytes, 16-byte
    // chunk index XORed as TMA's SWIZZLE_64B (CHP = 32) / SWIZZLE_32B (CHP = 16) does: conflict-free STS.128
    uint8_t* my_stg = stg + (uint32_t)(warp - 2) * STG_WARP;
    const uint32_t st_row = (uint32_t)lane * (uint32_t)(CHP * 2);
    const uint32_t st_x = CHP == 32 ? (((uint32_t)lane >> 1) & 3u) : (((uint32_t)lane >> 2) & 1u);
    int stores = 0;                            // lane 0: bulk store groups committed so far
    griddep_wait();

    for (int it = 0; it <= n_it; ++it) {
      const bool B_ = it < n_it, E_ = it >= 1;
      const uint32_t e = (uint32_t)(it - 1), ebuf = e & 1u, epar = (e >> 1) & 1u;
      const uint32_t t_acc = t_lane + T_ACC + ebuf * 128u + (uint32_t)(part * CHP), t_nrm = t_lane + T_NORM + (uint32_t)(part * CHP);
      const int ps = it % NP;
      const float* pp = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(patch) + ps * P1_SLOT) + pp_off;
      uint64_t xk[2 * NW];                      // x = acc + bias: xk[g*NW + i] = channels g*64 + part*CHP + 2i, +1

      // x^2 (hi | lo) of the CHP channels g*64 + part*CHP .. this thread read, stored over the columns they came from
      auto x2_work = [&](const int g, const uint32_t (&raw)[CHP]) {
        const uint64_t* bias2 = reinterpret_cast<const uint64_t*>(s_bias + g * 64 + part * CHP);
        uint32_t w[CHP];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const uint64_t x2 = add2(pk2(raw[2 * i], raw[2 * i + 1]), bias2[i]);
          xk[g * NW + i] = x2;
          float v0, v1;
          upk2(mul2(x2, x2), v0, v1);
          w[i] = cvt_bf16x2(v0, v1);
          w[NW + i] = lo_of(v0, v1, w[i]);
        }
        tmem_stN(t_acc + (uint32_t)(g * 64), w);
      };
      auto build = [&](auto kbc) {
        constexpr int KB = decltype(kbc)::value;
        uint32_t w[CHP];
        conv1_build_part<KB, NPART>(pp, part, w);
        tmem_stN(t_ring + (uint32_t)((KB & 1) * 64), w);
      };
      // my TMEM stores are complete and ordered before the arrival that follows
      auto publish = [&]() {
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
      };

      // ---- phase A: k-blocks 0, 1 of tile `it`
      if (B_) {
        mbar_wait(&patch_full[ps], (uint32_t)((it / NP) & 1));
        T1_MARK(2, mark);                          // [0] patch landed
        mbar_wait(&a_empty[0], 1u);                // uses 2*it of both slots
        mbar_wait(&a_empty[1], 1u);
        tcgen05_fence_after();
        build(std::integral_constant<int, 0>{});
        build(std::integral_constant<int, 1>{});
        publish();
        if (lane == 0) { mbar_arrive_cta0(&a_full[0]); mbar_arrive_cta0(&a_full[1]); }
      }
      T1_MARK(2, mark);                            // [1] kb0, kb1 built
      // ---- phase B: x^2 round 0 of tile it-1
      if (E_) {
        mbar_wait(&accum_full[ebuf], epar);
        tcgen05_fence_after();
        T1_MARK(2, mark);                          // [2] conv accumulator ready
        uint32_t raw[CHP];
        tmem_ldN(t_acc, raw);
        tmem_ld_wait();
        x2_work(0, raw);
        publish();
        if (lane == 0) mbar_arrive_cta0(&xs_full[0]);
        T1_MARK(2, mark);                          // [3] x^2 round 0 written
      }
      // ---- phase C: k-block 2 of tile `it` (while the gamma MMAs of round 0 run), x^2 round 1
      {
        uint32_t raw[CHP];
        if (E_) tmem_ldN(t_acc + 64u, raw);
        if (B_) {
          mbar_wait(&a_empty[0], 0u);              // use 2*it + 1
          tcgen05_fence_after();
          build(std::integral_constant<int, 2>{});
        }
        if (E_) {
          tmem_ld_wait();
          x2_work(1, raw);
        }
      }
      publish();
      if (lane == 0) {
        if (E_) mbar_arrive_cta0(&xs_full[1]);
        if (B_) mbar_arrive_cta0(&a_full[0]);
      }
      T1_MARK(2, mark);                            // [4] kb2 + x^2 round 1 written
      // ---- phase D: k-block 3 of tile `it` (while the gamma MMAs of round 1 run)
      if (B_) {
        mbar_wait(&a_empty[1], 0u);
        tcgen05_fence_after();
        build(std::integral_constant<int, 3>{});
        publish();
        if (lane == 0) { mbar_arrive_cta0(&a_full[1]); mbar_arrive(&patch_empty[ps]); }   // this warp is done with the patch
        T1_MARK(2, mark);                          // [5] kb3 written
      }
      // ---- phase E: final pass of tile it-1: y = x * rsqrt(beta + norm) -> NHWC bf16 (hi, lo).
      // 16-column chunks, the TMEM load of chunk j+1 in flight while chunk j is computed (MUFU) and staged; after the
      // last chunk of a 64-channel block the warp's lane 0 issues the block's two TMA stores.
      if (E_) {
        mbar_wait(norm_full, e & 1u);
        tcgen05_fence_after();
        T1_MARK(2, mark);                          // [6] norm ready
        constexpr int CPB = CHP / 16;              // chunks per 64-channel block
        uint32_t rawn[2][16];
        tmem_ld16(t_nrm, rawn[0]);
#pragma unroll
        for (int j = 0; j < 2 * CPB; ++j) {      // chunk j: channels (j / CPB) * 64 + part * CHP + (j % CPB) * 16 .. +15
          const int blk = j / CPB, sub = j % CPB;
          tmem_ld_wait();
          if (j + 1 < 2 * CPB)
            tmem_ld16(t_nrm + (uint32_t)(((j + 1) / CPB) * 64 + ((j + 1) % CPB) * 16), rawn[(j + 1) & 1]);
          const uint64_t* beta2 = reinterpret_cast<const uint64_t*>(s_beta + blk * 64 + part * CHP + sub * 16);
          uint32_t hp[8], lp[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float n0, n1, y0, y1;
            upk2(add2(pk2(rawn[j & 1][2 * i], rawn[j & 1][2 * i + 1]), beta2[i]), n0, n1);
            const uint64_t sc = pkf2(rsqrt_fast(n0), rsqrt_fast(n1));
            upk2(mul2(xk[blk * NW + sub * 8 + i], sc), y0, y1);
            hp[i] = cvt_bf16x2(y0, y1);
            lp[i] = lo_of(y0, y1, hp[i]);
          }
          uint8_t* sb = my_stg + (uint32_t)blk * (2 * STG_PLANE);
          if (sub == 0) {                        // the stores that last read this block's slice (previous tile) are done
            if (lane == 0 && stores >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
          const uint32_t k0 = (uint32_t)sub * 2u;                // 16-byte chunks k0, k0 + 1 of the row
          *reinterpret_cast<uint4*>(sb + st_row + (((k0) ^ st_x) << 4)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          *reinterpret_cast<uint4*>(sb + st_row + (((k0 + 1) ^ st_x) << 4)) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
          *reinterpret_cast<uint4*>(sb + STG_PLANE + st_row + (((k0) ^ st_x) << 4)) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
          *reinterpret_cast<uint4*>(sb + STG_PLANE + st_row + (((k0 + 1) ^ st_x) << 4)) = make_uint4(lp[4], lp[5], lp[6], lp[7]);
          if (sub == CPB - 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int cc = blk * 64 + part * CHP;
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
        T1_MARK(2, mark);                          // [7] staged + stores issued
      }
      if (B_) {                                  // the tile just built is the next iteration's epilogue tile
        const int tile = (unit0 + it * unit_stride) * 2 + (int)cta_rank;
        e_img = tile / p.tiles_per_image;
        const int trem = tile - e_img * p.tiles_per_image;
        const int tile_y = trem / p.tiles_x;
        e_oy0 = tile_y * TH;
        e_ox0 = (trem - tile_y * p.tiles_x) * TW;
      }
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
template <int NPART>
static int launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace);

int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  static const int mode = [] {
    const char* e = getenv("B2C_TC_PERSIST_CONV1");
    return e ? atoi(e) : 3;
  }();
  if (mode != 2) return 1;
  static const int nparts = [] {               // B2C_CONV1_PARTS=4: 16 worker warps per CTA instead of 8 (measured
                                               // slower: 0.171 vs 0.158 ms - 96-register cap, twice the arrivals)
    const char* e = getenv("B2C_CONV1_PARTS");
    return e && atoi(e) == 4 ? 4 : 2;
  }();
  return nparts == 2 ? launch<2>(L, stream, trace) : launch<4>(L, stream, trace);
}

template <int NPART>
static int launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  constexpr int NTHREADS = Cfg<NPART>::NTHREADS;
  const int channels = L->channels ? L->channels : 128;
  if (L->kind != B2C_TC_CONV1 || channels != 128 || !L->split3 || L->gdn != 1 || L->out_f32_nchw || !L->out_lo ||
      L->saved_y || !L->w_lo || !L->gamma_lo)
    return 1;
  auto kern = conv1_gdn_ts_kernel<NPART>;

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
      if (n <= 0) { set_error("conv1 ts: no resident CTA pairs"); return B2C_ECUDA; }
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
  if ((rc = make_mat_map(&mB_lo, L->w_lo, NKB * C, KBLK, C / 2))) return rc;
  if ((rc = make_mat_map(&mG_hi, L->gamma_hi, C, C, C / 2))) return rc;
  if ((rc = make_mat_map(&mG_lo, L->gamma_lo, C, C, C / 2))) return rc;
  if ((rc = make_warp_out_map(&mO_hi, L->out_hi, L->B, p.Hg, p.Wg, Cfg<NPART>::CHP))) return rc;
  if ((rc = make_warp_out_map(&mO_lo, L->out_lo, L->B, p.Hg, p.Wg, Cfg<NPART>::CHP))) return rc;

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
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(conv1 ts): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  count_launch();
  return check_launch("tc_layer_run(persistent conv1, operands in TMEM)");
}

}  // namespace tcp1ts
}  // namespace b2c
