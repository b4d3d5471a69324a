// Persistent tcgen05 kernel for the first analysis layer: conv 9x9 / 4, 3 -> 128 (+bias) fused with GDN1
// (models/analysis_17.py:14,17,33; models/GDN.py:64-94), bf16x3 arithmetic, NHWC bf16 (hi, lo) output.
//
// One cta_group::2 CTA pair per two SMs walks a static round-robin list of pairs of adjacent 8x16 output tiles.
// Everything the tensor pipe reads more than once stays in shared memory for the whole kernel:
//   * the conv weights (4 k-blocks x (hi, lo), K = 243 -> 256) and gamma (2 k-blocks x (hi, lo)); each CTA of the
//     pair holds HALF of the rows of every tile (cta_group::2 MMAs read B from both CTAs), 96 KB per CTA;
//   * the im2col A operand is built in software from the TMA-staged fp32 patch (3 x 37 x 72) into a two-slot ring
//     of (hi, lo) k-block tiles.
// The per-tile chain patch -> build -> conv MMA -> x^2 -> gamma MMA -> rsqrt -> store of the one-tile-per-CTA
// kernel (b2c_conv_tc.cu) is software-pipelined across tiles: the 256 row workers build the k-blocks of tile
// i+1 in the gaps of tile i's epilogue (where they would otherwise wait for the gamma MMAs), two TMEM
// accumulator sets let the conv MMAs of tile i+1 run under the epilogue of tile i, and the next patch is
// fetched while the current tile's epilogue finishes.  Worker order per iteration `it` (B = build tile it,
// E = epilogue tile it-1; `|` = fence + worker barrier + signals):
//                            B.kb0 B.kb1 | E.x^2[0] | B.kb2 E.x^2[1] | B.kb3 | E.final
// MMA-warp order (static, deadlock-free against the order above):
//                            conv kb0  conv kb1  gamma[0]  gamma[1]  conv kb2  conv kb3
#include <mutex>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"
#include "b2c_tc_host.cuh"
#include "b2c_tcp_util.cuh"

namespace b2c {
namespace tcp1 {

using namespace b2c::tc;

constexpr int C = 128;
constexpr int NCB = 2;                               // gamma k-blocks
constexpr int NKB = 4;                               // conv k-blocks (K = 256)
constexpr int NWORKERS = 256;
constexpr int NTHREADS = 64 + NWORKERS;
constexpr uint32_t A_PLANE = SLOT_BYTES;             // 16 KB: 128 rows x 64 bf16
constexpr uint32_t A_SLOT = 2 * A_PLANE;             // (hi, lo)
constexpr int NA = 2;
constexpr uint32_t B_BYTES = (C / 2) * 128;          // this CTA's half of a 128-row weight / gamma tile
constexpr uint32_t W_RES = NKB * 2 * B_BYTES;        // 64 KB
constexpr uint32_t GAMMA_BYTES = NCB * 2 * B_BYTES;  // 32 KB
constexpr uint32_t XS_BYTES = 2 * SLOT_BYTES;        // x^2 (hi, lo) tile of one round / output staging
constexpr uint32_t SMEM = NA * A_SLOT + W_RES + GAMMA_BYTES + XS_BYTES + P1_SLOT + 1024 /*align*/ + 1024 /*bias, beta*/ +
                          512 /*barriers*/;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t ACC_STRIDE = 256;                 // TMEM columns per accumulator set: conv acc + norm

struct Params {
  int B, Hg, Wg;
  int tiles_x, tiles_per_image, n_tiles;
  const float* bias;
  const float* beta_eff;
  long long* trace;      // debug: clock64 timeline of CTA 0: [3 roles][64 marks]
};

#define WORKER_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(NWORKERS) : "memory")

// im2col k-block KB of this CTA's tile -> A slot (KB & 1); the caller waits for the slot and signals a_full
template <int KB>
__device__ __forceinline__ void build_rows(const float* pp, uint8_t* smem, int r, int half) {
  uint8_t* st = smem + (KB & 1) * A_SLOT;
  if (half == 0) conv1_build_row<KB, 0, true>(pp, st, st + A_PLANE, r);
  else conv1_build_row<KB, 1, true>(pp, st, st + A_PLANE, r);
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv1_gdn_tcp_kernel(const __grid_constant__ CUtensorMap tmImg, const __grid_constant__ CUtensorMap tmB_hi,
                     const __grid_constant__ CUtensorMap tmB_lo, const __grid_constant__ CUtensorMap tmG_hi,
                     const __grid_constant__ CUtensorMap tmG_lo, const __grid_constant__ CUtensorMap tmO_hi,
                     const __grid_constant__ CUtensorMap tmO_lo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wres = smem + NA * A_SLOT;                  // [hi kb0..3][lo kb0..3], 64 rows each
  uint8_t* gam = wres + W_RES;                         // [hi kb0][hi kb1][lo kb0][lo kb1]
  uint8_t* xs = gam + GAMMA_BYTES;
  float* patch = reinterpret_cast<float*>(xs + XS_BYTES);
  float* s_bias = reinterpret_cast<float*>(xs + XS_BYTES + P1_SLOT);
  float* s_beta = s_bias + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + 128);
  uint64_t* a_full = bars;                     // [2] (leader's: one elected worker per CTA)
  uint64_t* a_empty = a_full + NA;             // [2] (multicast commit)
  uint64_t* accum_full = a_empty + NA;         // [2]
  uint64_t* norm_full = accum_full + 2;        // [2]
  uint64_t* acc_empty = norm_full + 2;         // [2] (leader's)
  uint64_t* xs_full = acc_empty + 2;           // (leader's)
  uint64_t* xs_empty = xs_full + 1;
  uint64_t* res_full = xs_empty + 1;           // resident weights + gamma landed (leader's)
  uint64_t* patch_full = res_full + 1;         // local
  uint64_t* patch_empty = patch_full + 1;      // local
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(patch_empty + 1);

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
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 2); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&accum_full[i], 1);
      mbar_init(&norm_full[i], 1);
      mbar_init(&acc_empty[i], 2);
    }
    mbar_init(xs_full, 2);
    mbar_init(xs_empty, 1);
    mbar_init(res_full, 1);
    mbar_init(patch_full, 1);
    mbar_init(patch_empty, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmImg);
    tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    tma_prefetch_desc(&tmG_hi); tma_prefetch_desc(&tmG_lo);
    tma_prefetch_desc(&tmO_hi); tma_prefetch_desc(&tmO_lo);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 2 * ACC_STRIDE);
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
        mbar_wait(patch_empty, (it & 1) ^ 1);
        T1_MARK(0, true);
        mbar_arrive_expect_tx(patch_full, (uint32_t)P1_BYTES);
        tma_load_4d(patch, &tmImg, patch_full, tile_x * TW * 4 - 4, tile_y * TH * 4 - 4, 0, n_img);
        if (it + 1 < n_it) {    // next tile's patch -> L2 now, so the load issued when the buffer frees skips the HBM latency
          const int ntile = (unit0 + (it + 1) * unit_stride) * 2 + (int)cta_rank;
          const int ni = ntile / p.tiles_per_image;
          const int nrem = ntile - ni * p.tiles_per_image;
          const int nty = nrem / p.tiles_x;
          if (ni < p.B) tma_prefetch_4d(&tmImg, (nrem - nty * p.tiles_x) * TW * 4 - 4, nty * TH * 4 - 4, 0, ni);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (whole warp, one elected lane issues)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16_m256(C);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint64_t desc0 = umma_desc_sw128(sbase);
      const uint64_t wres_d = desc0 + ((uint32_t)(NA * A_SLOT) >> 4);
      const uint64_t gam_d = wres_d + (W_RES >> 4);
      const uint64_t xs_d = gam_d + (GAMMA_BYTES >> 4);
      uint32_t xs_cnt = 0;

      // conv k-block KB of the tile being accumulated into set `d`
      auto conv_kb = [&](int kb, uint32_t buf, bool last) {
        const uint32_t d = tb + buf * ACC_STRIDE;
        const int slot = kb & 1;
        mbar_wait_cluster(&a_full[slot], (uint32_t)((kb >> 1) & 1));
        tcgen05_fence_after();
        T1_MARK(1, lane == 0);
        if (elect_one()) {
          const uint64_t a_hi = desc0 + ((uint32_t)(slot * A_SLOT) >> 4), a_lo = a_hi + (A_PLANE >> 4);
          const uint64_t b_hi = wres_d + ((uint32_t)(kb * B_BYTES) >> 4);
          const uint64_t b_lo = wres_d + ((uint32_t)((NKB + kb) * B_BYTES) >> 4);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
            umma_bf16_pair(d, a_hi + 2 * k4, b_hi + 2 * k4, idesc, acc);
            umma_bf16_pair(d, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1u);
            umma_bf16_pair(d, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1u);
          }
          umma_commit_pair(&a_empty[slot]);
          if (last) umma_commit_pair(&accum_full[buf]);
        }
        __syncwarp();
      };
      // gamma k-block `round` of the tile in set `ebuf`: norm += x^2[:, round*64 ..] * gamma[:, round*64 ..]^T
      auto gamma_round = [&](int round, uint32_t ebuf) {
        mbar_wait_cluster(xs_full, xs_cnt & 1u);
        ++xs_cnt;
        tcgen05_fence_after();
        T1_MARK(1, lane == 0);
        if (elect_one()) {
          const uint32_t d = tb + ebuf * ACC_STRIDE + 128u;
          const uint64_t a_hi = xs_d, a_lo = xs_d + (SLOT_BYTES >> 4);
          const uint64_t b_hi = gam_d + ((uint32_t)(round * B_BYTES) >> 4);
          const uint64_t b_lo = gam_d + ((uint32_t)((NCB + round) * B_BYTES) >> 4);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acc = (round > 0 || k4 > 0) ? 1u : 0u;
            umma_bf16_pair(d, a_hi + 2 * k4, b_hi + 2 * k4, idesc, acc);
            umma_bf16_pair(d, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1u);
            umma_bf16_pair(d, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1u);
          }
          if (round == NCB - 1) umma_commit_pair(&norm_full[ebuf]);
          else umma_commit_pair(xs_empty);
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
        if (E_) gamma_round(1, ebuf);           // posted together with k-block 2: the workers need the norm first
        if (B_) conv_kb(2, buf, false);
        if (B_) conv_kb(3, buf, true);
      }
    }
  } else {
    // ================================================================ row workers (8 warps): thread = (pixel r, half)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ty = r / TW, tx = r - ty * TW;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool elected = warp == 2 && lane == 0;
    const float* pp = patch + (4 * ty) * P1_W + 4 * tx;
    bool store_pending = false;
    int e_img = 0, e_ox0 = 0, e_oy0 = 0;       // tile whose epilogue runs in the current iteration
    griddep_wait();

    for (int it = 0; it <= n_it; ++it) {
      const bool B_ = it < n_it, E_ = it >= 1;
      const uint32_t e = (uint32_t)(it - 1), ebuf = e & 1u, epar = (e >> 1) & 1u;
      const uint32_t t_acc = t_lane + ebuf * ACC_STRIDE, t_nrm = t_acc + 128u;
      uint64_t xk[32];                          // x = acc + bias: xk[g*16 + i] = channels g*64 + half*32 + 2i, +1

      // x^2 (hi, lo) of k-block g as the K-major swizzled A tile of the gamma product
      auto x2_work = [&](const int g, const uint32_t (&raw)[32]) {
        const uint64_t* bias2 = reinterpret_cast<const uint64_t*>(s_bias + g * 64 + half * 32);
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
            l[j] = lo_of(v0, v1, h[j]);
          }
          sts128(xs, r, half * 4 + ci, h[0], h[1], h[2], h[3]);
          sts128(xs + SLOT_BYTES, r, half * 4 + ci, l[0], l[1], l[2], l[3]);
        }
      };

      // Each phase ends in ONE fence + worker barrier, after which the elected thread posts all of its signals
      // (every fence / barrier / arrival exposes ~200 clk with two warps per scheduler).
      // ---- phase A: k-blocks 0, 1 of tile `it`
      if (B_) {
        mbar_wait(patch_full, (uint32_t)it & 1u);
        T1_MARK(2, threadIdx.x == 64);             // [0] patch landed
        mbar_wait(&a_empty[0], 1u);                // uses 2*it of both slots
        mbar_wait(&a_empty[1], 1u);
        build_rows<0>(pp, smem, r, half);
        build_rows<1>(pp, smem, r, half);
        fence_proxy_async_smem();
      }
      // the previous tile's last store must have left the staging buffer before phase B writes x^2 into it
      if (elected && store_pending) { tma_store_wait_read(); store_pending = false; }
      WORKER_SYNC();
      if (B_ && elected) { mbar_arrive_cta0(&a_full[0]); mbar_arrive_cta0(&a_full[1]); }
      T1_MARK(2, threadIdx.x == 64);               // [1] kb0, kb1 built
      // ---- phase B: x^2 round 0 of tile it-1
      if (E_) {
        mbar_wait(&accum_full[ebuf], epar);
        tcgen05_fence_after();
        T1_MARK(2, threadIdx.x == 64);             // [2] conv accumulator ready
        uint32_t raw[32];
        tmem_ld32(t_acc + (uint32_t)(half * 32), raw);
        tmem_ld_wait();
        x2_work(0, raw);
        fence_proxy_async_smem();
        WORKER_SYNC();
        if (elected) mbar_arrive_cta0(xs_full);
        T1_MARK(2, threadIdx.x == 64);             // [3] x^2 round 0 written
      }
      // ---- phase C: k-block 2 of tile `it` (while the gamma MMAs of round 0 read the x^2 tile), x^2 round 1
      {
        uint32_t raw[32];
        if (E_) tmem_ld32(t_acc + (uint32_t)(64 + half * 32), raw);
        if (B_) {
          mbar_wait(&a_empty[0], 0u);              // use 2*it + 1
          build_rows<2>(pp, smem, r, half);
        }
        if (E_) {
          tmem_ld_wait();
          mbar_wait(xs_empty, e & 1u);             // the gamma MMAs of round 0 have read the x^2 tile
          x2_work(1, raw);
        }
      }
      fence_proxy_async_smem();
      WORKER_SYNC();
      if (elected) {
        if (E_) mbar_arrive_cta0(xs_full);
        if (B_) mbar_arrive_cta0(&a_full[0]);
      }
      T1_MARK(2, threadIdx.x == 64);               // [4] kb2 + x^2 round 1 written
      // ---- phase D: k-block 3 of tile `it` (while the gamma MMAs of round 1 run)
      if (B_) {
        mbar_wait(&a_empty[1], 0u);
        build_rows<3>(pp, smem, r, half);
        fence_proxy_async_smem();
        WORKER_SYNC();
        if (elected) { mbar_arrive_cta0(&a_full[1]); mbar_arrive(patch_empty); }   // every worker is done with the patch
        T1_MARK(2, threadIdx.x == 64);             // [5] kb3 written
      }
      // ---- phase E: final pass of tile it-1
      if (E_) {
        mbar_wait(&norm_full[ebuf], epar);         // also: every gamma MMA has finished reading the x^2 tile
        tcgen05_fence_after();
        T1_MARK(2, threadIdx.x == 64);             // [6] norm ready
        // ---- y = x * rsqrt(beta + norm) -> NHWC bf16 through the x^2 buffer + TMA stores, one 64-channel block
        // (hi plane | lo plane) at a time: block 1 is computed while the TMA unit still reads block 0
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = h * 64 + half * 32;
          uint32_t rawn[32];
          tmem_ld32(t_nrm + (uint32_t)c0, rawn);
          tmem_ld_wait();
          if (h == 1) tcgen05_fence_before();
          const uint64_t* beta2 = reinterpret_cast<const uint64_t*>(s_beta + c0);
          uint32_t hp[16], lp[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float n0, n1, y0, y1;
            upk2(add2(pk2(rawn[2 * i], rawn[2 * i + 1]), beta2[i]), n0, n1);
            const uint64_t sc = pkf2(rsqrt_fast(n0), rsqrt_fast(n1));
            upk2(mul2(xk[h * 16 + i], sc), y0, y1);
            hp[i] = cvt_bf16x2(y0, y1);
            lp[i] = lo_of(y0, y1, hp[i]);
          }
          if (h == 1) {                          // block 0's stores must have left the staging buffer
            if (elected) tma_store_wait_read();
            WORKER_SYNC();
          }
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            sts128(xs, r, half * 4 + ci, hp[ci * 4 + 0], hp[ci * 4 + 1], hp[ci * 4 + 2], hp[ci * 4 + 3]);
            sts128(xs + SLOT_BYTES, r, half * 4 + ci, lp[ci * 4 + 0], lp[ci * 4 + 1], lp[ci * 4 + 2], lp[ci * 4 + 3]);
          }
          fence_proxy_async_smem();
          WORKER_SYNC();                        // staging complete (h = 1: every worker has finished its TMEM reads)
          if (elected) {
            if (h == 1) mbar_arrive_cta0(&acc_empty[ebuf]);
            tma_store_4d(&tmO_hi, xs, h * KBLK, e_ox0, e_oy0, e_img);
            tma_store_4d(&tmO_lo, xs + SLOT_BYTES, h * KBLK, e_ox0, e_oy0, e_img);
            tma_store_commit();
            if (h == 1) store_pending = true;
          }
        }
        T1_MARK(2, threadIdx.x == 64);
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
    if (elected && store_pending) tma_store_wait_read();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * ACC_STRIDE);
  }
#undef T1_MARK
}
#undef WORKER_SYNC

// ------------------------------------------------------------------------------------ host side
// Returns 1 when the layer is not the one this kernel covers (the caller then uses the per-tile kernel),
// otherwise the launch status.
int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace) {
  static const int mode = [] {
    const char* e = getenv("B2C_TC_PERSIST_CONV1");
    return e ? atoi(e) : 1;
  }();
  if (!mode) return 1;
  const int channels = L->channels ? L->channels : 128;
  if (L->kind != B2C_TC_CONV1 || channels != 128 || !L->split3 || L->gdn != 1 || L->out_f32_nchw || !L->out_lo ||
      L->saved_y || !L->w_lo || !L->gamma_lo)
    return 1;
  auto kern = conv1_gdn_tcp_kernel;

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
      if (n <= 0) { set_error("conv1 tcp: no resident CTA pairs"); return B2C_ECUDA; }
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
  if ((rc = make_out_map(&mO_hi, L->out_hi, L->B, p.Hg, p.Wg, p.Hg, p.Wg, 1, 0, 0, C))) return rc;
  if ((rc = make_out_map(&mO_lo, L->out_lo, L->B, p.Hg, p.Wg, p.Hg, p.Wg, 1, 0, 0, C))) return rc;

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
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(conv1 tcp): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  count_launch();
  return check_launch("tc_layer_run(persistent conv1)");
}

}  // namespace tcp1
}  // namespace b2c
