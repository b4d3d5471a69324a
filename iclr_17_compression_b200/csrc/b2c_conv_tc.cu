// tcgen05 implicit-GEMM convolution family for the inference path (out_channel_N = 128).
//
//   CONV1      9x9 s4  3->128   A tile built in shared memory from the fp32 NCHW image (K 243 -> 256)
//   CONV5S2    5x5 s2 128->128  A tiles = TMA boxes of the NHWC bf16 input with traversal stride 2
//   DECONV5S2  5x5 s2 transposed, 4 sub-pixel phases, each a stride-1 tap subset (no zero insertion)
//   DECONV3    9x9 s4 transposed 128->3, as a 3x3 stride-1 conv to 48 = 16 phases x 3 channels
//
// One CTA = one 8x16-pixel output tile (M = 128 rows = TMEM lanes) x all output channels.
//   warp 0    TMA producer  (A boxes + tap-major weight panels, SWIZZLE_128B, mbarrier ring)
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer (fp32 accumulators in TMEM)
//   warps 2-5 row workers: software A tiles (CONV1, x^2 for GDN) and the epilogue (tcgen05.ld)
// GDN / IGDN is a second MMA in the same CTA: the workers square the biased accumulator, write it
// as a K-major swizzled bf16 tile, the MMA warp multiplies it by gamma (TMA) into a second TMEM
// accumulator, and the workers finish y = x / sqrt(beta + norm)  (or x * sqrt).
// SPLIT3: every operand is a (hi, lo) bf16 pair and each product is hi*hi + hi*lo + lo*hi
// (fp32 accumulate) - the error-compensated mode the analysis transform needs for round().
#include <mutex>

#include "b2c_common.cuh"
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"
#include "b2c_tc_host.cuh"
#include "b2c_deconv3x_pack.cuh"

namespace b2c {
namespace tc {

#ifndef B2C_TC_PAIRS_DEFAULT
#define B2C_TC_PAIRS_DEFAULT 1
#endif
#ifndef B2C_GDNBWD_PREFETCH
#define B2C_GDNBWD_PREFETCH 1
#endif
#ifndef B2C_TC_DEEP_STAGES
#define B2C_TC_DEEP_STAGES 3
#endif
#ifndef B2C_CONV1_STAGES
#define B2C_CONV1_STAGES 1
#endif
// channels C in {128, 192, 256}: C/64 k-blocks per tap, C/64 row workers per tile row (64 channels each)
constexpr int nworkers(int C) { return 128 * (C / 64); }
constexpr int nthreads(int C) { return 64 + nworkers(C); }   // + TMA producer warp + MMA warp

enum { OUT_BF16 = 0, OUT_F32_NCHW = 1, OUT_F32_NCHW_PS4 = 2 };

struct TcParams {
  int B, Hin, Win;      // input grid
  int Hg, Wg;           // tile grid (pixels): conv -> output dims; deconv -> input dims
  int tiles_x;          // tiles per row
  int Hfull, Wfull;     // full output dims
  int ps;               // output lattice stride (1 conv, 2 deconv5s2, 4 deconv3)
  const float* bias;
  const float* beta_eff;
  void* out_hi;
  void* out_lo;
  const __nv_bfloat16* saved_y;  // GDN backward epilogue: forward output y and norm n = beta + gamma x^2
  const __nv_bfloat16* saved_n;  //   (NHWC bf16, indexed like the output of this launch)
  const float* ref_image;  // DECONV3 fused distortion: input image x (fp32 NCHW) or null
  float* sse_partial;      // [B][tiles][2] per-CTA {sum (r-x)^2, sum (clamp(r)-x)^2}
  int tiles_per_image;
  int ksplit;           // CONV5S2, fp32 NCHW output, no GDN: the 25 taps are split over gridDim.y CTAs that add their
                        // partial sums into the zero-filled output with red.global (1 = off)
  long long* trace;     // debug: clock64 timeline of CTA (0,0,0): [3 roles][64 marks]
};

template <int KIND, int NCB>
__device__ __forceinline__ int num_steps(int phase) {
  if (KIND == B2C_TC_CONV1) return 4;
  if (KIND == B2C_TC_CONV5S2) return 25 * NCB;
  if (KIND == B2C_TC_DECONV3) return 9 * NCB;
  const int ry = phase >> 1, rx = phase & 1;  // DECONV5S2: (3-ry) x (3-rx) taps x NCB channel blocks
  return (3 - ry) * (3 - rx) * NCB;
}

// step -> A box offset (dy, dx in input pixels), channel block, weight-panel group
template <int KIND, int NCB>
__device__ __forceinline__ void decode_step(int s, int phase, int& dy, int& dx, int& cb, int& group) {
  if (KIND == B2C_TC_CONV1) {
    dy = dx = 0; cb = 0; group = s;
  } else if (KIND == B2C_TC_CONV5S2) {
    cb = s % NCB;
    const int t = s / NCB, ky = t / 5, kx = t - ky * 5;
    dy = ky - 2; dx = kx - 2; group = t;                 // iy = 2*oy - 2 + ky
  } else if (KIND == B2C_TC_DECONV3) {
    cb = s % NCB;
    const int t = s / NCB, sy = t / 3, sx = t - sy * 3;
    dy = sy - 1; dx = sx - 1; group = t;                 // iy = my + dy
  } else {                                               // DECONV5S2: ky = ry + 2 - 2*dy, dy = 1 - jy
    const int ry = phase >> 1, rx = phase & 1;
    cb = s % NCB;
    const int t = s / NCB, ntx = 3 - rx, jy = t / ntx, jx = t - jy * ntx;
    dy = 1 - jy; dx = 1 - jx;
    group = (ry + 2 * jy) * 5 + (rx + 2 * jx);
  }
}

// pipeline depth / CTAs per SM: the plain-bf16 kernels keep <= 113 KB of shared memory so two CTAs share an
// SM and one CTA's epilogue overlaps the other's MMA main loop; the split (hi, lo) kernels use the whole SM.
// NHWC bf16 output tiles leave through TMA stores: conv kinds use [0] = hi plane, [1] = lo plane; DECONV5S2
// uses one strided map per sub-pixel phase (the phase's output lattice as a dense 4-D tensor).
struct TcOutMaps {
  CUtensorMap m[8];   // [0..3] primary plane (conv: [0]; deconv: per phase), [4..7] second plane (lo / norm / dnorm)
};

template <int C, bool SPLIT3, int KIND, int GDN>
struct TcCfg {
  static constexpr int NCB = C / KBLK;
  // weight / gamma tile: C rows of 128 B (DECONV3: 48 rows = 6 KB in an 8 KB slot)
  static constexpr int B_SLOT = (KIND == B2C_TC_DECONV3) ? 8192 : C * 128;
  static constexpr int STAGE_BYTES = (SPLIT3 ? 2 : 1) * (SLOT_BYTES + B_SLOT);
  static constexpr int PATCH = (KIND == B2C_TC_CONV1) ? P1_SLOT : 0;
  static constexpr int CONSTS = (C <= 128) ? 128 : 256;       // floats of bias and of beta kept in smem
  static constexpr int EXTRA = 1024 /*align*/ + 2 * CONSTS * 4 /*bias, beta*/ + 256 /*barriers*/;
  static constexpr int N_TILE = (KIND == B2C_TC_DECONV3) ? 48 : C;
  static constexpr int TMEM_COLS = GDN ? (2 * C <= 256 ? 256 : 512) : (N_TILE <= 64 ? 64 : (N_TILE <= 128 ? 128 : 256));
  static constexpr int MAX_SMEM = 227 * 1024;
  // B2C_TC_DEEP_STAGES (compile-time experiment, default 3 = off): ring depth of the plain-bf16 128-channel 5x5
  // kinds.  6 stages (one CTA per SM) measured only -1 % on the training step: the main loop is bound by shared
  // memory bandwidth (MMA operand reads + TMA fills), not by TMA latency, and 2 CTAs / SM keep 256-CTA grids in one wave.
  static constexpr bool DEEP = !SPLIT3 && C == 128 && (KIND == B2C_TC_CONV5S2 || KIND == B2C_TC_DECONV5S2);
  static constexpr int STAGES =
      (KIND == B2C_TC_CONV1 && SPLIT3 && C == 128) ? B2C_CONV1_STAGES
      : (DEEP && B2C_TC_DEEP_STAGES * STAGE_BYTES + PATCH + EXTRA <= MAX_SMEM) ? B2C_TC_DEEP_STAGES
                                                   : ((3 * STAGE_BYTES + PATCH + EXTRA <= MAX_SMEM)
                                                          ? 3
                                                          : ((2 * STAGE_BYTES + PATCH + EXTRA <= MAX_SMEM) ? 2 : 1));
  static constexpr int SMEM = STAGES * STAGE_BYTES + PATCH + EXTRA;
  static_assert(SMEM <= MAX_SMEM, "shared memory budget");
  static constexpr int BY_SMEM = (SMEM <= 75 * 1024) ? 3 : (SMEM <= 113 * 1024) ? 2 : 1;
  static constexpr int BY_TMEM = 512 / TMEM_COLS;
  // GDN backward epilogues (GDN 3 / 4) keep this thread's saved (y, n) rows in 64 registers, prefetched under the
  // main loop: one CTA per SM (register budget), see B2C_GDNBWD_PREFETCH
  static constexpr int MIN_CTAS = (GDN >= 3 && B2C_GDNBWD_PREFETCH) ? 1 : (BY_SMEM < BY_TMEM ? BY_SMEM : BY_TMEM);
};

template <int C, bool SPLIT3, int KIND, int GDN, int OUT_MODE, bool PAIR>
__global__ void __launch_bounds__(nthreads(C), TcCfg<C, SPLIT3, KIND, GDN>::MIN_CTAS)
igemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ CUtensorMap tmG_hi, const __grid_constant__ CUtensorMap tmG_lo,
                const __grid_constant__ TcOutMaps tmO, const TcParams p) {
  using Cfg = TcCfg<C, SPLIT3, KIND, GDN>;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NCB = Cfg::NCB;                // 64-channel k-blocks (and row workers per tile row)
  constexpr int N_TILE = Cfg::N_TILE;
  constexpr int B_SLOT = Cfg::B_SLOT;
  constexpr int NWORKERS = nworkers(C);
  constexpr int IN_STRIDE = (KIND == B2C_TC_CONV5S2) ? 2 : 1;
  constexpr bool SOFT_A = (KIND == B2C_TC_CONV1);
  constexpr uint32_t TMEM_COLS = Cfg::TMEM_COLS;
  // PAIR: two CTAs (adjacent tiles) form a cta_group::2 pair: one M = 256 MMA per step issued by the leader,
  // each CTA stages its own A tile and HALF of the weight tile, halving B fill and B operand reads per SM.
  static_assert(!PAIR || KIND != B2C_TC_DECONV3, "CTA pairs: C-channel output kinds only");
  // DEC (CONV1 as a CTA pair): the stage is re-cut into ONE software A slot and a ring of TWO half-size weight
  // slots (each CTA of the pair stages half of the weight rows), so the next k-block's weights are prefetched
  // while the current one is built and multiplied - same shared-memory footprint, two CTAs per SM.
  constexpr bool DEC = SOFT_A && PAIR;
  static_assert(!DEC || (STAGES == 1 && SPLIT3), "decoupled rings are the single-stage bf16x3 CONV1 layout");
  constexpr int NA = DEC ? 1 : STAGES;                  // A ring depth
  constexpr int NB = DEC ? 2 : STAGES;                  // B ring depth
  constexpr int NBAR = NA > NB ? NA : NB;
  constexpr uint32_t B_HALF = B_SLOT / 2;
  constexpr uint32_t B_LO = DEC ? B_HALF : (uint32_t)B_SLOT;   // offset of the lo weight tile from the hi one
  auto a_off = [](int s) -> uint32_t { return DEC ? 0u : (uint32_t)((s % STAGES) * STAGE_BYTES); };
  auto b_off = [](int s) -> uint32_t {
    return DEC ? (uint32_t)(2 * SLOT_BYTES) + (uint32_t)(s % 2) * 2u * B_HALF
               : (uint32_t)((s % STAGES) * STAGE_BYTES) + (uint32_t)((SPLIT3 ? 2 : 1) * SLOT_BYTES);
  };
  constexpr uint32_t B_BYTES = N_TILE * KBLK * 2 / (PAIR ? 2 : 1);
  static_assert(GDN == 0 || N_TILE == C, "fused GDN needs all channels in the tile");
  // GDN: 0 none, 1 GDN forward, 2 IGDN forward, 3 GDN backward, 4 IGDN backward (epilogue of the dgrad conv)
  constexpr bool GDN_BWD = (GDN == 3 || GDN == 4);
  static_assert(!GDN_BWD || !SPLIT3, "the GDN backward epilogue is bf16 only");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* patch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);   // CONV1 only
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + Cfg::PATCH);
  float* s_beta = s_bias + Cfg::CONSTS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 2 * Cfg::CONSTS);
  uint64_t* full = bars;                    // [NB] TMA bytes landed
  uint64_t* empty = bars + NBAR;            // [NB] MMAs that read the (B or whole) stage completed
  uint64_t* afull = bars + 2 * NBAR;        // [NA] software-written A tile ready (NWORKERS arrivals per CTA)
  uint64_t* aempty = DEC ? bars + 3 * NBAR : empty;   // [NA] A slot released (its own barrier only when decoupled)
  uint64_t* accum_full = bars + 4 * NBAR;
  uint64_t* norm_full = bars + 4 * NBAR + 1;
  uint64_t* patch_full = bars + 4 * NBAR + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * NBAR + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int phase = (KIND == B2C_TC_CONV5S2) ? 0 : blockIdx.y;   // DECONV5S2: sub-pixel phase
  const int ks = (KIND == B2C_TC_CONV5S2) ? blockIdx.y : 0;       // CONV5S2: split-K slice (taps)
  const int n_img = blockIdx.z;
  const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x - tile_y * p.tiles_x;
  const int oy0 = tile_y * TH, ox0 = tile_x * TW;
  int nsteps = num_steps<KIND, NCB>(phase);
  int s_off = 0;                                                  // first global step of this CTA (split-K)
  if (KIND == B2C_TC_CONV5S2 && !GDN && p.ksplit > 1) {
    const int t0 = ks * 25 / p.ksplit, t1 = (ks + 1) * 25 / p.ksplit;
    s_off = t0 * NCB;
    nsteps = (t1 - t0) * NCB;
  }
  const int total_steps = nsteps + (GDN ? NCB : 0);
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int tmark = 0;
#define TC_MARK(role)                                                       \
  do {                                                                      \
    if (tracing && tmark < 64) p.trace[(role) * 64 + tmark++] = clock64(); \
  } while (0)

  griddep_launch_dependents();       // the next kernel's launch + prologue may overlap this kernel (no-op without PDL)
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBAR; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&afull[i], NWORKERS * (PAIR ? 2 : 1));
      if (DEC) mbar_init(&aempty[i], 1);
    }
    mbar_init(accum_full, 1);
    mbar_init(norm_full, 1);
    mbar_init(patch_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA_hi);
    if (!SOFT_A && SPLIT3) tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    if (SPLIT3) tma_prefetch_desc(&tmB_lo);
    if (GDN) { tma_prefetch_desc(&tmG_hi); if (SPLIT3) tma_prefetch_desc(&tmG_lo); }
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  // Programmatic dependent launch: everything above (barrier init, descriptor prefetch, TMEM allocation) overlapped
  // the predecessor's tail; every global read / write of this kernel happens after this point.
  griddep_wait();
  if (threadIdx.x >= 64 && threadIdx.x < 64 + C) {   // per-channel epilogue constants
    const int c = threadIdx.x - 64;
    s_bias[c] = (p.bias && c < (KIND == B2C_TC_DECONV3 ? 3 : C)) ? __ldg(p.bias + c) : 0.f;
    s_beta[c] = (GDN && p.beta_eff) ? __ldg(p.beta_eff + c) : 1.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer's barriers are initialised before any remote arrive / TMA signal
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      if (SOFT_A) {  // fp32 image patch (tmA_hi is the [B][3][H][W] fp32 map); OOB rows / columns are zero-filled
        mbar_arrive_expect_tx(patch_full, (uint32_t)P1_BYTES);
        tma_load_4d(patch, &tmA_hi, patch_full, ox0 * 4 - 4, oy0 * 4 - 4, 0, n_img);
      }
      TC_MARK(0);
      for (int s = 0; s < total_steps; ++s) {
        const int stage = s % NB, it = s / NB;
        mbar_wait(&empty[stage], (it & 1) ^ 1);
        TC_MARK(0);
        uint8_t* a_hi = smem + a_off(s);
        uint8_t* a_lo = a_hi + SLOT_BYTES;
        uint8_t* b_hi = smem + b_off(s);
        uint8_t* b_lo = b_hi + B_LO;
        if (s < nsteps) {
          int dy, dx, cb, group;
          decode_step<KIND, NCB>(s + s_off, phase, dy, dx, cb, group);
          const uint32_t a_bytes = SOFT_A ? 0u : (uint32_t)SLOT_BYTES;
          if (PAIR) {
            // bytes of BOTH CTAs are counted on the leader's barrier
            if (leader) mbar_arrive_expect_tx(&full[stage], 2u * (a_bytes + B_BYTES) * (SPLIT3 ? 2u : 1u));
            const int cx = ox0 * IN_STRIDE + dx, cy = oy0 * IN_STRIDE + dy;
            const int brow = group * N_TILE + (int)cta_rank * (N_TILE / 2);
            if (!SOFT_A) {
              tma_load_4d_pair(a_hi, &tmA_hi, &full[stage], cb * KBLK, cx, cy, n_img);
              if (SPLIT3) tma_load_4d_pair(a_lo, &tmA_lo, &full[stage], cb * KBLK, cx, cy, n_img);
            }
            tma_load_2d_pair(b_hi, &tmB_hi, &full[stage], cb * KBLK, brow);
            if (SPLIT3) tma_load_2d_pair(b_lo, &tmB_lo, &full[stage], cb * KBLK, brow);
          } else {
            mbar_arrive_expect_tx(&full[stage], (a_bytes + B_BYTES) * (SPLIT3 ? 2u : 1u));
            if (!SOFT_A) {
              const int cx = ox0 * IN_STRIDE + dx, cy = oy0 * IN_STRIDE + dy;
              tma_load_4d(a_hi, &tmA_hi, &full[stage], cb * KBLK, cx, cy, n_img);
              if (SPLIT3) tma_load_4d(a_lo, &tmA_lo, &full[stage], cb * KBLK, cx, cy, n_img);
            }
            tma_load_2d(b_hi, &tmB_hi, &full[stage], cb * KBLK, group * N_TILE);
            if (SPLIT3) tma_load_2d(b_lo, &tmB_lo, &full[stage], cb * KBLK, group * N_TILE);
          }
        } else {
          const int g = s - nsteps;  // GDN k-block: gamma[:, g*64 .. g*64+63]
          if (PAIR) {
            if (leader) mbar_arrive_expect_tx(&full[stage], (uint32_t)(C * KBLK * 2) * (SPLIT3 ? 2u : 1u));
            tma_load_2d_pair(b_hi, &tmG_hi, &full[stage], g * KBLK, (int)cta_rank * (C / 2));
            if (SPLIT3) tma_load_2d_pair(b_lo, &tmG_lo, &full[stage], g * KBLK, (int)cta_rank * (C / 2));
          } else {
            mbar_arrive_expect_tx(&full[stage], (uint32_t)(C * KBLK * 2) * (SPLIT3 ? 2u : 1u));
            tma_load_2d(b_hi, &tmG_hi, &full[stage], g * KBLK, 0);
            if (SPLIT3) tma_load_2d(b_lo, &tmG_lo, &full[stage], g * KBLK, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    // The whole warp runs the (warp-uniform) loop and one elected lane issues: descriptors, TMEM addresses and
    // loop state then live in uniform registers.  (A single-lane region forced a R2UR/ELECT round trip in front
    // of every UTCHMMA and made the kernel MMA-issue bound at ~100 clk per instruction.)
    if (leader) {
      constexpr uint32_t idesc_main = PAIR ? umma_idesc_bf16_m256(N_TILE) : umma_idesc_bf16(N_TILE);
      constexpr uint32_t idesc_gdn = PAIR ? umma_idesc_bf16_m256(C) : umma_idesc_bf16(C);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint64_t desc0 = umma_desc_sw128(sbase);     // + (byte offset >> 4) selects a tile / K slice
      uint32_t soft_phase = 0;                            // bit per stage: parity of the next software-A fill
      for (int s = 0; s < total_steps; ++s) {
        const int stage = s % NB, it = s / NB, sa = s % NA;
        const bool is_gdn = s >= nsteps;
        mbar_wait(&full[stage], it & 1);
        if (SOFT_A || is_gdn) {
          if (PAIR) mbar_wait_cluster(&afull[sa], (soft_phase >> sa) & 1u);
          else mbar_wait(&afull[sa], (soft_phase >> sa) & 1u);
          soft_phase ^= 1u << sa;
        }
        tcgen05_fence_after();
        if (lane == 0) TC_MARK(1);
        const uint64_t a_hi = desc0 + (a_off(s) >> 4), a_lo = a_hi + (SLOT_BYTES >> 4);
        const uint64_t b_hi = desc0 + (b_off(s) >> 4), b_lo = b_hi + (B_LO >> 4);
        const uint32_t d = tb + (is_gdn ? (uint32_t)C : 0u);
        const uint32_t idesc = is_gdn ? idesc_gdn : idesc_main;
        const uint32_t acc0 = (s > (is_gdn ? nsteps : 0)) ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {  // 4 x (K = 16) per 64-wide k-block: +32 B inside the swizzle atom
            const uint32_t acc = k4 > 0 ? 1u : acc0;
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
          if (PAIR) {
            umma_commit_pair(&empty[stage]);
            if (DEC) umma_commit_pair(&aempty[sa]);
            if (s == nsteps - 1) umma_commit_pair(accum_full);
            if (GDN && s == total_steps - 1) umma_commit_pair(norm_full);
          } else {
            umma_commit(&empty[stage]);
            if (s == nsteps - 1) umma_commit(accum_full);
            if (GDN && s == total_steps - 1) umma_commit(norm_full);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================ row workers (8 warps, 256 threads)
    // NCB threads per tile row: `half` (0 .. NCB-1) selects which 64 channels (or which k-chunks) it owns
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;            // tile row = pixel
    const int ty = r / TW, tx = r - ty * TW;
    const int oy = oy0 + ty, ox = ox0 + tx;
    const bool valid = oy < p.Hg && ox < p.Wg;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool wtrace = tracing && r == 0 && half == 0;
#define TC_WMARK()                                                     \
  do {                                                                 \
    if (wtrace && tmark < 64) p.trace[2 * 64 + tmark++] = clock64();   \
  } while (0)
    TC_WMARK();

    // DECONV3 fused distortion: fetch this thread's 2 rows x 3 channels x 4 pixels of the reference image now, so
    // the HBM latency hides behind the main loop instead of serialising the epilogue
    float4 xref[6];
    if (OUT_MODE == OUT_F32_NCHW_PS4) {
#pragma unroll
      for (int i = 0; i < 6; ++i) xref[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.ref_image && valid && half < 2) {
#pragma unroll
        for (int o = 0; o < 3; ++o)
#pragma unroll
          for (int ryl = 0; ryl < 2; ++ryl)
            xref[o * 2 + ryl] = __ldg(reinterpret_cast<const float4*>(
                p.ref_image + (((size_t)n_img * 3 + o) * p.Hfull + (oy * 4 + half * 2 + ryl)) * p.Wfull + ox * 4));
      }
    }

    if (SOFT_A) {
      // CONV1: im2col rows from the TMA-staged fp32 patch; pixel (ty,tx) starts at patch[4*ty][4*tx]
      const float* pp = patch + (4 * ty) * P1_W + 4 * tx;
      mbar_wait(patch_full, 0);
      TC_WMARK();
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int stage = s % NA, it = s / NA;
        mbar_wait(&aempty[stage], (it & 1) ^ 1);
        uint8_t* st = smem + a_off(s);
        if (half >= 2) {
          // C > 128: the 8 chunks of a k-block are built by the first two workers of the row
        } else if (half == 0) {
          if (s == 0) conv1_build_row<0, 0, SPLIT3>(pp, st, st + SLOT_BYTES, r);
          else if (s == 1) conv1_build_row<1, 0, SPLIT3>(pp, st, st + SLOT_BYTES, r);
          else if (s == 2) conv1_build_row<2, 0, SPLIT3>(pp, st, st + SLOT_BYTES, r);
          else conv1_build_row<3, 0, SPLIT3>(pp, st, st + SLOT_BYTES, r);
        } else {
          if (s == 0) conv1_build_row<0, 1, SPLIT3>(pp, st, st + SLOT_BYTES, r);
          else if (s == 1) conv1_build_row<1, 1, SPLIT3>(pp, st, st + SLOT_BYTES, r);
          else if (s == 2) conv1_build_row<2, 1, SPLIT3>(pp, st, st + SLOT_BYTES, r);
          else conv1_build_row<3, 1, SPLIT3>(pp, st, st + SLOT_BYTES, r);
        }
        fence_proxy_async_smem();
        if (PAIR) mbar_arrive_leader(&afull[stage]);
        else mbar_arrive(&afull[stage]);
        TC_WMARK();
      }
    }

    // GDN backward epilogue: this thread's 64 channels of the saved forward output y and norm n (2 x 128 B) are
    // fetched NOW, while the main loop runs, and stay in registers for both epilogue passes.
    uint4 sv_y[GDN_BWD ? 8 : 1], sv_n[GDN_BWD ? 8 : 1];
    {
      const size_t pixs = (((size_t)n_img * p.Hfull + (oy * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase >> 1) : 0))) *
                               p.Wfull + (ox * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase & 1) : 0))) * C + half * 64;
      (void)pixs;
#if B2C_GDNBWD_PREFETCH
      if (GDN_BWD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sv_y[i] = make_uint4(0, 0, 0, 0);
          sv_n[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
          if (valid) {
            sv_y[i] = __ldg(reinterpret_cast<const uint4*>(p.saved_y + pixs) + i);
            sv_n[i] = __ldg(reinterpret_cast<const uint4*>(p.saved_n + pixs) + i);
          }
        }
      }
#endif
    }
    const size_t pix_saved = (((size_t)n_img * p.Hfull + (oy * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase >> 1) : 0))) *
                                  p.Wfull + (ox * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase & 1) : 0))) * C + half * 64;
    auto saved_chunk = [&](int i, uint4& qy, uint4& qn) {      // 8 channels (i = 0..7) of this thread's 64
#if B2C_GDNBWD_PREFETCH
      qy = sv_y[i]; qn = sv_n[i];
#else
      qy = make_uint4(0, 0, 0, 0);
      qn = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
      if (valid) {
        qy = __ldg(reinterpret_cast<const uint4*>(p.saved_y + pix_saved) + i);
        qn = __ldg(reinterpret_cast<const uint4*>(p.saved_n + pix_saved) + i);
      }
#endif
    };

    mbar_wait(accum_full, 0);
    tcgen05_fence_after();
    TC_WMARK();

    if (GDN) {
      // ---- pass 1: (acc + bias)^2 -> K-major swizzled bf16 A tile for the gamma product.
      // thread (r, half) writes k-block g = half (its 64 channels) and only arrives on the other one.
#pragma unroll 1
      for (int g = 0; g < NCB; ++g) {
        const int s = nsteps + g;
        const int stage = s % NA, it = s / NA;
        // every worker waits for the slot (with one A slot, k-block 1 reuses k-block 0's slot and barrier:
        // arriving early would be counted in the previous phase)
        mbar_wait(&aempty[stage], (it & 1) ^ 1);
        if (g == half) {
          uint8_t* st = smem + a_off(s);
          uint32_t raw[64];
          tmem_ld32(t_row + (uint32_t)(g * 64), *reinterpret_cast<uint32_t(*)[32]>(&raw[0]));
          tmem_ld32(t_row + (uint32_t)(g * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(&raw[32]));
          tmem_ld_wait();
          if (GDN_BWD) {
            // acc = dL/dy.  dL/dnorm = -+ 1/2 * g * y / n  (GDN: y = x n^-1/2, IGDN: y = x n^1/2)
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx) {
              uint4 qy, qn;
              saved_chunk(cidx, qy, qn);
              const uint32_t wy[4] = {qy.x, qy.y, qy.z, qy.w}, wn[4] = {qn.x, qn.y, qn.z, qn.w};
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float yv = __uint_as_float((j & 1) ? (wy[j >> 1] & 0xffff0000u) : (wy[j >> 1] << 16));
                const float nv = __uint_as_float((j & 1) ? (wn[j >> 1] & 0xffff0000u) : (wn[j >> 1] << 16));
                const float gv = __uint_as_float(raw[cidx * 8 + j]);
                float rn;                                        // n is a bf16 value >= beta_min > 0: one MUFU op
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"(nv));
                v[j] = (GDN == 3 ? -0.5f : 0.5f) * gv * yv * rn;
              }
              store_row_chunk<false>(st, st + SLOT_BYTES, r, cidx, v);
            }
          } else {
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float xv = __uint_as_float(raw[cidx * 8 + j]) + s_bias[g * 64 + cidx * 8 + j];
                v[j] = xv * xv;
              }
              store_row_chunk<SPLIT3>(st, st + SLOT_BYTES, r, cidx, v);
            }
          }
          fence_proxy_async_smem();
        }
        if (PAIR) mbar_arrive_leader(&afull[stage]);
        else mbar_arrive(&afull[stage]);
      }
      TC_WMARK();
      mbar_wait(norm_full, 0);
      tcgen05_fence_after();
      TC_WMARK();
    }

    // ---- final pass: TMEM -> registers -> global
    const int oyf = oy * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase >> 1) : 0);
    const int oxf = ox * p.ps + (KIND == B2C_TC_DECONV5S2 ? (phase & 1) : 0);
    if (OUT_MODE == OUT_F32_NCHW_PS4) {
      // DECONV3: column = (ry*4+rx)*3 + o  ->  out[n, o, 4*oy+ry, 4*ox+rx]; this thread owns ry = 2*half, 2*half+1.
      // With a reference image the distortion is fused: out = clamp(r, 0, 1) (model.py:59) and the CTA's
      // partial sums of (r - x)^2 (model.py:61) and (clamp(r) - x)^2 (train.py:172) go to sse_partial.
      uint32_t raw[24];
      const int hh = half < 2 ? half : 0;      // warp-uniform; workers beyond the first two per row only reduce
      tmem_ld8(t_row + (uint32_t)(hh * 24), *reinterpret_cast<uint32_t(*)[8]>(&raw[0]));
      tmem_ld8(t_row + (uint32_t)(hh * 24 + 8), *reinterpret_cast<uint32_t(*)[8]>(&raw[8]));
      tmem_ld8(t_row + (uint32_t)(hh * 24 + 16), *reinterpret_cast<uint32_t(*)[8]>(&raw[16]));
      tmem_ld_wait();
      float s0 = 0.f, s1 = 0.f;
      if (valid && half < 2) {
        float* out = reinterpret_cast<float*>(p.out_hi);
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
              float d;
              d = v4.x - x4.x; s0 = fmaf(d, d, s0); d = v4.y - x4.y; s0 = fmaf(d, d, s0);
              d = v4.z - x4.z; s0 = fmaf(d, d, s0); d = v4.w - x4.w; s0 = fmaf(d, d, s0);
              d = c4.x - x4.x; s1 = fmaf(d, d, s1); d = c4.y - x4.y; s1 = fmaf(d, d, s1);
              d = c4.z - x4.z; s1 = fmaf(d, d, s1); d = c4.w - x4.w; s1 = fmaf(d, d, s1);
              v4 = c4;
            }
            *reinterpret_cast<float4*>(out + off) = v4;
          }
        }
      }
      if (p.ref_image) {   // fixed-order CTA reduction: warp shuffle, then the 8 worker warps through smem
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        float* red = s_beta;   // unused by DECONV3 (no GDN)
        if (lane == 0) { red[(warp - 2) * 2] = s0; red[(warp - 2) * 2 + 1] = s1; }
        asm volatile("bar.sync 1, %0;" ::"n"(NWORKERS) : "memory");
        if (warp == 2 && lane == 0) {
          float t0 = 0.f, t1 = 0.f;
#pragma unroll
          for (int w = 0; w < NWORKERS / 32; ++w) { t0 += red[w * 2]; t1 += red[w * 2 + 1]; }
          float* dst = p.sse_partial + ((size_t)n_img * p.tiles_per_image + blockIdx.x) * 2;
          dst[0] = t0;
          dst[1] = t1;
        }
      }
    } else if (OUT_MODE == OUT_F32_NCHW) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {            // 2 x 32 channels of this thread's 64
        const int c0 = half * 64 + h * 32;
        uint32_t raw[32], rawn[32];
        tmem_ld32(t_row + (uint32_t)c0, raw);
        if (GDN) tmem_ld32(t_row + (uint32_t)(C + c0), rawn);
        tmem_ld_wait();
        if (valid) {
          float* out = reinterpret_cast<float*>(p.out_hi);
          const bool split_k = KIND == B2C_TC_CONV5S2 && !GDN && p.ksplit > 1;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float xv = __uint_as_float(raw[j]) + ((split_k && ks != 0) ? 0.f : s_bias[c0 + j]);
            if (GDN) {
              const float nb = __uint_as_float(rawn[j]) + s_beta[c0 + j];
              const float rs = rsqrtf(nb);
              xv = (GDN == 2) ? xv * (nb * rs) : xv * rs;
            }
            float* dst = out + (((size_t)n_img * N_TILE + c0 + j) * p.Hfull + oyf) * p.Wfull + oxf;
            if (split_k) atomicAdd(dst, xv);      // lanes = consecutive pixels: coalesced red.global.add
            else *dst = xv;
          }
        }
      }
    } else {
      // NHWC bf16: stage the 128 x 64-channel half tiles in idle ring slots in the same swizzled K-major
      // layout the loads use, then one elected thread issues TMA stores (coalesced 128-byte rows,
      // image-border clipping done by the TMA unit).
      //   SPLIT3: stage 0 = {hi k-blocks 0,1 ; lo k-blocks 0,1}.
      //   bf16:   primary plane -> both slots of the stage the GDN steps did not touch; second plane (forward:
      //           norm n, for training; backward: dL/dnorm, which already sits in the GDN steps' A slots).
      //   C != 128 (inference only): all planes staged from the start of the ring.
      constexpr bool SIMPLE_STAGING = SPLIT3 || C != 128;
      const int st_c = SIMPLE_STAGING ? 0 : (nsteps + 2) % STAGES;
      uint8_t* stg_hi = smem + st_c * STAGE_BYTES + half * SLOT_BYTES;
      uint8_t* stg_lo = smem + (NCB + half) * SLOT_BYTES;                                 // SPLIT3 only
      const bool second_plane = !SPLIT3 && C == 128 && GDN != 0 && p.out_lo != nullptr;
      uint8_t* stg_n = smem + ((nsteps + half) % STAGES) * STAGE_BYTES + SLOT_BYTES;     // forward norm: B slot
      const size_t pixb = (((size_t)n_img * p.Hfull + oyf) * p.Wfull + oxf) * C;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {            // 2 x 32 channels of this thread's 64
        const int c0 = half * 64 + h * 32;
        uint32_t raw[32], rawn[32];
        tmem_ld32(t_row + (uint32_t)c0, raw);
        if (GDN) tmem_ld32(t_row + (uint32_t)(C + c0), rawn);
        tmem_ld_wait();
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          float v[8], vn[8];
          uint32_t wy[4] = {0, 0, 0, 0}, wn[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
          if (GDN_BWD) {
            // registers prefetched before the main loop; the h loop is rolled, so select with constant indices
            uint4 qy, qn;
#if B2C_GDNBWD_PREFETCH
            uint4 qy1, qn1;              // the h loop is rolled: select between constant register indices
            saved_chunk(ci, qy, qn);
            saved_chunk(4 + ci, qy1, qn1);
            if (h) { qy = qy1; qn = qn1; }
#else
            saved_chunk(h * 4 + ci, qy, qn);
#endif
            wy[0] = qy.x; wy[1] = qy.y; wy[2] = qy.z; wy[3] = qy.w;
            wn[0] = qn.x; wn[1] = qn.y; wn[2] = qn.z; wn[3] = qn.w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float xv = __uint_as_float(raw[ci * 8 + j]) + s_bias[c0 + ci * 8 + j];
            if (GDN == 1 || GDN == 2) {
              const float nb = __uint_as_float(rawn[ci * 8 + j]) + s_beta[c0 + ci * 8 + j];
              const float rs = rsqrtf(nb);
              xv = (GDN == 2) ? xv * (nb * rs) : xv * rs;
              vn[j] = nb;
            } else if (GDN_BWD) {
              // dL/dx = g * n^-+1/2 + 2 x * (gamma^T dn),  x = y n^+-1/2
              const float yv = __uint_as_float((j & 1) ? (wy[j >> 1] & 0xffff0000u) : (wy[j >> 1] << 16));
              const float nv = __uint_as_float((j & 1) ? (wn[j >> 1] & 0xffff0000u) : (wn[j >> 1] << 16));
              const float rs = rsqrtf(nv), sq = nv * rs;
              const float t = __uint_as_float(rawn[ci * 8 + j]);
              xv = (GDN == 3) ? fmaf(2.f * yv * sq, t, xv * rs) : fmaf(2.f * yv * rs, t, xv * sq);
            }
            v[j] = xv;
          }
          store_row_chunk<SPLIT3>(stg_hi, stg_lo, r, h * 4 + ci, v);
          if (second_plane && !GDN_BWD) store_row_chunk<false>(stg_n, stg_n, r, h * 4 + ci, vn);
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(NWORKERS) : "memory");
      if (warp == 2 && lane == 0) {
        const int mi = (KIND == B2C_TC_DECONV5S2) ? phase : 0;
        uint8_t* c_base = smem + st_c * STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < NCB; ++j) tma_store_4d(&tmO.m[mi], c_base + j * SLOT_BYTES, j * KBLK, ox0, oy0, n_img);
        if (SPLIT3) {
#pragma unroll
          for (int j = 0; j < NCB; ++j)
            tma_store_4d(&tmO.m[4 + mi], smem + (NCB + j) * SLOT_BYTES, j * KBLK, ox0, oy0, n_img);
        } else if (second_plane) {
          // forward: norm tiles in the B slots; backward: dnorm tiles still in the A slots of the GDN steps
          const int off = GDN_BWD ? 0 : SLOT_BYTES;
          tma_store_4d(&tmO.m[4 + mi], smem + ((nsteps + 0) % STAGES) * STAGE_BYTES + off, 0, ox0, oy0, n_img);
          tma_store_4d(&tmO.m[4 + mi], smem + ((nsteps + 1) % STAGES) * STAGE_BYTES + off, KBLK, ox0, oy0, n_img);
        }
        tma_store_commit_and_wait();
      }
    }
    TC_WMARK();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
#undef TC_MARK
#undef TC_WMARK
}

// ------------------------------------------------------------------------------------ packing kernels
// weights -> bf16 (hi, lo) panels, row = (group, out channel), 128-byte K-major rows
__device__ __forceinline__ void pack_weight_elem(int kind, const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                                 __nv_bfloat16* __restrict__ lo, int idx, int C) {
  float v = 0.f;
  if (kind == B2C_TC_CONV1) {            // [kb 4][o C][j 64], k = kb*64 + j, w OIHW [C][3][9][9]
    const int j = idx & 63, o = (idx >> 6) % C, kb = idx / (64 * C);
    const int k = kb * 64 + j;
    if (k < 243) v = w[o * 243 + k];
  } else if (kind == B2C_TC_CONV5S2) {   // [tap 25][o C][c C], w OIHW [C][C][5][5]
    const int c = idx % C, o = (idx / C) % C, t = idx / (C * C);
    v = w[(o * C + c) * 25 + t];
  } else if (kind == B2C_TC_DECONV5S2) { // [tap 25][o C][c C], w IOHW [C c][C o][5][5]
    const int c = idx % C, o = (idx / C) % C, t = idx / (C * C);
    v = w[(c * C + o) * 25 + t];
  } else {                               // DECONV3: [shift 9][n 48][c C], n = (ry*4+rx)*3+o, w IOHW [C][3][9][9]
    const int c = idx % C, n = (idx / C) % 48, sh = idx / (C * 48);
    const int o = n % 3, ph = n / 3, ry = ph >> 2, rx = ph & 3;
    const int dy = sh / 3 - 1, dx = sh % 3 - 1;
    const int ky = ry + 4 - 4 * dy, kx = rx + 4 - 4 * dx;
    if (ky >= 0 && ky < 9 && kx >= 0 && kx < 9) v = w[((c * 3 + o) * 9 + ky) * 9 + kx];
  }
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  hi[idx] = h;
  if (lo) lo[idx] = l;
}

__global__ void pack_weights_kernel(int kind, const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int total, int C) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  pack_weight_elem(kind, w, hi, lo, idx, C);
}

__device__ __forceinline__ void pack_gdn_elem(const float* __restrict__ beta_p, const float* __restrict__ gamma_p,
                                              __nv_bfloat16* __restrict__ ghi, __nv_bfloat16* __restrict__ glo,
                                              float* __restrict__ beta_eff, int transpose, int C, int idx) {
  if (idx < C * C) {
    const int src = transpose ? ((idx % C) * C + (idx / C)) : idx;
    const float b = fmaxf(gamma_p[src], 3.814697265625e-06f);
    const float g = b * b - 1.4551915228366852e-11f;
    __nv_bfloat16 h, l;
    split_bf16(g, h, l);
    ghi[idx] = h;
    if (glo) glo[idx] = l;
  }
  if (idx < C) {
    const float b = fmaxf(beta_p[idx], (float)0.0010000072759311445);
    beta_eff[idx] = b * b - 1.4551915228366852e-11f;
  }
}

__global__ void pack_gdn_kernel(const float* __restrict__ beta_p, const float* __restrict__ gamma_p,
                                __nv_bfloat16* __restrict__ ghi, __nv_bfloat16* __restrict__ glo,
                                float* __restrict__ beta_eff, int transpose, int C) {
  pdl_trigger();
  pdl_wait();
  pack_gdn_elem(beta_p, gamma_p, ghi, glo, beta_eff, transpose, C, blockIdx.x * blockDim.x + threadIdx.x);
}

// several packing jobs in one launch: blockIdx.y = job, blockIdx.x strides over the job's elements
constexpr int PACK_MAX_JOBS = 16;
struct PackJobs {
  b2c_tc_pack_job j[PACK_MAX_JOBS];
  int total[PACK_MAX_JOBS];
};
__global__ void __launch_bounds__(256) pack_many_kernel(const __grid_constant__ PackJobs J) {
  pdl_trigger();
  pdl_wait();
  const b2c_tc_pack_job& job = J.j[blockIdx.y];
  const int total = J.total[blockIdx.y];
  if ((job.kind == B2C_TC_CONV5S2 || job.kind == B2C_TC_DECONV5S2) && job.channels == 128) {
    // 5x5 panels through shared memory: block = output channel o.  The element-wise form reads w with a stride of 25
    // floats (one 32-byte sector per element: 17 us for the eight forward operands of a training step); here the 3200
    // source floats of one o are read in 100-byte runs (conv: one contiguous 12.8 KB run) and every tap row
    // out[t][o][0..127] leaves as one 256-byte write.
    __shared__ float s_w[128 * 25];
    const int o = blockIdx.x;
    if (o >= 128) return;
    const float* __restrict__ w = job.src;
    for (int e = threadIdx.x; e < 128 * 25; e += 256) {
      const int c = e / 25, t = e - c * 25;
      s_w[e] = (job.kind == B2C_TC_CONV5S2) ? __ldg(w + (size_t)o * 3200 + e) : __ldg(w + ((size_t)c * 128 + o) * 25 + t);
    }
    __syncthreads();
    uint32_t* hi = reinterpret_cast<uint32_t*>(job.dst_hi);
    uint32_t* lo = reinterpret_cast<uint32_t*>(job.dst_lo);
    for (int i = threadIdx.x; i < 25 * 64; i += 256) {
      const int t = i >> 6, cp = i & 63;
      const float v0 = s_w[(2 * cp) * 25 + t], v1 = s_w[(2 * cp + 1) * 25 + t];
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v0, h0, l0);
      split_bf16(v1, h1, l1);
      const size_t dst = ((size_t)t * 128 + o) * 64 + cp;
      hi[dst] = pack_bf16x2(h0, h1);
      if (lo) lo[dst] = pack_bf16x2(l0, l1);
    }
    return;
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    if (job.kind == B2C_TC_PACK_GDN || job.kind == B2C_TC_PACK_GDN_T)
      pack_gdn_elem(job.src, job.src2, (__nv_bfloat16*)job.dst_hi, (__nv_bfloat16*)job.dst_lo, job.beta_eff,
                    job.kind == B2C_TC_PACK_GDN_T, job.channels, idx);
    else if (job.kind == B2C_TC_DECONV3X)
      ((__nv_bfloat16*)job.dst_hi)[idx] = __float2bfloat16_rn(tcp3x::deconv3x_weight(job.src, idx, job.channels));
    else
      pack_weight_elem(job.kind, job.src, (__nv_bfloat16*)job.dst_hi, (__nv_bfloat16*)job.dst_lo, idx, job.channels);
  }
}

// fp32 NCHW -> NHWC bf16 (hi[,lo]) through a 32x32 shared-memory transpose
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int C, int HW) {
  pdl_trigger();
  pdl_wait();
  __shared__ float t[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pix = p0 + threadIdx.x;
    t[i][threadIdx.x] = (c < C && pix < HW) ? in[((size_t)n * C + c) * HW + pix] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pix = p0 + i, c = c0 + threadIdx.x;
    if (c < C && pix < HW) {
      __nv_bfloat16 h, l;
      split_bf16(t[threadIdx.x][i], h, l);
      hi[((size_t)n * HW + pix) * C + c] = h;
      if (lo) lo[((size_t)n * HW + pix) * C + c] = l;
    }
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                    float* __restrict__ out, int C, int HW) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pix = p0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (c < C && pix < HW) {
      const size_t a = ((size_t)n * HW + pix) * C + c;
      v = __bfloat162float(hi[a]) + (lo ? __bfloat162float(lo[a]) : 0.f);
    }
    t[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pix = p0 + threadIdx.x;
    if (c < C && pix < HW) out[((size_t)n * C + c) * HW + pix] = t[threadIdx.x][i];
  }
}

template <int C, bool SPLIT3, int KIND, int GDN, int OUT_MODE, bool PAIR = false>
static int launch(const b2c_tc_layer* L, const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int N_TILE = TcCfg<C, SPLIT3, KIND, GDN>::N_TILE;
  constexpr int NTHREADS = nthreads(C);
  constexpr size_t SMEM = TcCfg<C, SPLIT3, KIND, GDN>::SMEM;
  auto kern = igemm_tc_kernel<C, SPLIT3, KIND, GDN, OUT_MODE, PAIR>;
  {  // opt in to > 48 KB dynamic shared memory once per (instantiation, device); never during graph capture replays
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
  }

  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mG_hi, mG_lo;
  memset(&mA_hi, 0, sizeof(mA_hi)); memset(&mA_lo, 0, sizeof(mA_lo));
  memset(&mG_hi, 0, sizeof(mG_hi)); memset(&mG_lo, 0, sizeof(mG_lo));
  memset(&mB_lo, 0, sizeof(mB_lo));
  int rc;
  constexpr int S = (KIND == B2C_TC_CONV5S2) ? 2 : 1;
  if (KIND == B2C_TC_CONV1) {
    if ((rc = make_image_map(&mA_hi, L->in_hi, L->B, L->Hin, L->Win))) return rc;
  } else {
    if ((rc = make_act_map(&mA_hi, L->in_hi, L->B, L->Hin, L->Win, S, C))) return rc;
    if (SPLIT3 && (rc = make_act_map(&mA_lo, L->in_lo, L->B, L->Hin, L->Win, S, C))) return rc;
  }
  const int groups = KIND == B2C_TC_CONV1 ? 4 : (KIND == B2C_TC_DECONV3 ? 9 : 25);
  const int wcols = KIND == B2C_TC_CONV1 ? 64 : C;
  constexpr int BROWS = PAIR ? N_TILE / 2 : N_TILE;     // a CTA pair stages half of the weight rows per CTA
  if ((rc = make_mat_map(&mB_hi, L->w_hi, groups * N_TILE, wcols, BROWS))) return rc;
  if (SPLIT3 && (rc = make_mat_map(&mB_lo, L->w_lo, groups * N_TILE, wcols, BROWS))) return rc;
  if (GDN) {
    constexpr int GROWS = PAIR ? C / 2 : C;
    if ((rc = make_mat_map(&mG_hi, L->gamma_hi, C, C, GROWS))) return rc;
    if (SPLIT3 && (rc = make_mat_map(&mG_lo, L->gamma_lo, C, C, GROWS))) return rc;
  }
  TcOutMaps mO;
  memset(&mO, 0, sizeof(mO));
  if (OUT_MODE == OUT_BF16) {
    const int nph = (KIND == B2C_TC_DECONV5S2) ? 4 : 1;
    const int ps = (KIND == B2C_TC_DECONV5S2) ? 2 : 1;
    for (int ph = 0; ph < nph; ++ph) {
      const int ry = nph == 4 ? (ph >> 1) : 0, rx = nph == 4 ? (ph & 1) : 0;
      if ((rc = make_out_map(&mO.m[ph], L->out_hi, L->B, p.Hg, p.Wg, p.Hfull, p.Wfull, ps, ry, rx, C))) return rc;
      if (L->out_lo && (rc = make_out_map(&mO.m[4 + ph], L->out_lo, L->B, p.Hg, p.Wg, p.Hfull, p.Wfull, ps, ry, rx, C)))
        return rc;
    }
  }
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((grid.x + 1) & ~1u, grid.y, grid.z);     // an odd tile count gets one fully masked tile
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mA_hi, mA_lo, mB_hi, mB_lo, mG_hi, mG_lo, mO, p);
    if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(pair): %s", cudaGetErrorString(e)); return B2C_ECUDA; }
    count_launch();
    return check_launch("tc_layer_run(pair)");
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mA_hi, mA_lo, mB_hi, mB_lo, mG_hi, mG_lo, mO, p);
    if (e != cudaSuccess) { set_error("cudaLaunchKernelEx: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  }
  count_launch();
  return check_launch("tc_layer_run");
}

}  // namespace tc
}  // namespace b2c

namespace b2c {
namespace tcp {
int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace);   // b2c_conv_tcp.cu; 1 = not covered
}
namespace tcp1rs {
int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace);   // b2c_conv1_rs.cu; 1 = not covered
}
namespace tcp1ts {
int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace);   // b2c_conv1_ts.cu; 1 = not covered
}
namespace tcp1 {
int try_launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace);   // b2c_conv1_tcp.cu; 1 = not covered
}
namespace tcp3x {                                                                // b2c_deconv3x.cu
int launch(const b2c_tc_layer* L, cudaStream_t stream, long long* trace);
int pack(const float* w, void* out, int channels, cudaStream_t stream);
size_t packed_elems(int channels);
int tiles_per_image(int Hin, int Win);
}
}  // namespace b2c

using namespace b2c;
using namespace b2c::tc;

static long long* g_tc_trace = nullptr;

extern "C" {

/* debug hook (not part of the stable ABI): device buffer of 3*64 int64 receiving a clock64 timeline */
void b2c_tc_debug_set_trace(void* device_buffer) { g_tc_trace = (long long*)device_buffer; }

static bool tc_channels_ok(int C) { return C == 128 || C == 192 || C == 256; }

size_t b2c_tc_packed_weight_elems(int kind, int channels) {
  const size_t C = (size_t)channels;
  if (!tc_channels_ok(channels)) return 0;
  switch (kind) {
    case B2C_TC_CONV1: return 4 * C * 64;
    case B2C_TC_CONV5S2:
    case B2C_TC_DECONV5S2: return 25 * C * C;
    case B2C_TC_DECONV3: return 9 * 48 * C;
    case B2C_TC_DECONV3X: return b2c::tcp3x::packed_elems(channels);
    default: return 0;
  }
}

int b2c_tc_pack_weights(int kind, int channels, const float* w, void* w_hi, void* w_lo, void* stream) {
  const size_t n = b2c_tc_packed_weight_elems(kind, channels);
  B2C_REQUIRE(n > 0 && w && w_hi, "tc_pack_weights: bad arguments (channels must be 128, 192 or 256)");
  if (kind == B2C_TC_DECONV3X) return b2c::tcp3x::pack(w, w_hi, channels, (cudaStream_t)stream);   // plain bf16 only
  pack_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      kind, w, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo, (int)n, channels);
  count_launch();
  return check_launch("tc_pack_weights");
}

int b2c_tc_pack_many(const b2c_tc_pack_job* jobs, int n_jobs, void* stream) {
  B2C_REQUIRE(jobs && n_jobs > 0 && n_jobs <= PACK_MAX_JOBS, "tc_pack_many: 1..%d jobs", PACK_MAX_JOBS);
  PackJobs J;
  memset(&J, 0, sizeof(J));
  int max_total = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2c_tc_pack_job& j = jobs[i];
    B2C_REQUIRE(j.channels == 128 && j.src && j.dst_hi, "tc_pack_many: job %d: 128 channels, src and dst_hi required", i);
    int total;
    if (j.kind == B2C_TC_PACK_GDN || j.kind == B2C_TC_PACK_GDN_T) {
      B2C_REQUIRE(j.src2 && j.beta_eff, "tc_pack_many: job %d: GDN jobs need gamma_p (src2) and beta_eff", i);
      total = j.channels * j.channels;
    } else {
      total = (int)b2c_tc_packed_weight_elems(j.kind, j.channels);
      B2C_REQUIRE(total > 0, "tc_pack_many: job %d: unknown kind %d", i, j.kind);
    }
    J.j[i] = j;
    J.total[i] = total;
    if (total > max_total) max_total = total;
  }
  int bx = (max_total + 255) / 256;
  if (bx > 400) bx = 400;
  pack_many_kernel<<<dim3(bx, n_jobs), 256, 0, (cudaStream_t)stream>>>(J);
  count_launch();
  return check_launch("tc_pack_many");
}

int b2c_tc_pack_gdn(int channels, const float* beta_p, const float* gamma_p, void* gamma_hi, void* gamma_lo,
                    float* beta_eff, void* stream) {
  B2C_REQUIRE(tc_channels_ok(channels) && beta_p && gamma_p && gamma_hi && beta_eff, "tc_pack_gdn: bad arguments");
  pack_gdn_kernel<<<(channels * channels + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      beta_p, gamma_p, (__nv_bfloat16*)gamma_hi, (__nv_bfloat16*)gamma_lo, beta_eff, 0, channels);
  count_launch();
  return check_launch("tc_pack_gdn");
}

int b2c_tc_pack_gdn_t(int channels, const float* beta_p, const float* gamma_p, void* gamma_t, float* beta_eff,
                      void* stream) {
  B2C_REQUIRE(tc_channels_ok(channels) && beta_p && gamma_p && gamma_t && beta_eff, "tc_pack_gdn_t: bad arguments");
  pack_gdn_kernel<<<(channels * channels + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      beta_p, gamma_p, (__nv_bfloat16*)gamma_t, nullptr, beta_eff, 1, channels);
  count_launch();
  return check_launch("tc_pack_gdn_t");
}

int b2c_tc_nchw_to_nhwc(const float* in, void* out_hi, void* out_lo, int B, int C, int HW, void* stream) {
  B2C_REQUIRE(in && out_hi && B > 0 && C > 0 && HW > 0, "tc_nchw_to_nhwc: bad arguments");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_pdl(PDL_LAYOUT, nchw_to_nhwc_kernel, grid, block, 0, (cudaStream_t)stream, in, (__nv_bfloat16*)out_hi,
             (__nv_bfloat16*)out_lo, C, HW);
  count_launch();
  return check_launch("tc_nchw_to_nhwc");
}

int b2c_tc_nhwc_to_nchw(const void* in_hi, const void* in_lo, float* out, int B, int C, int HW, void* stream) {
  B2C_REQUIRE(in_hi && out && B > 0 && C > 0 && HW > 0, "tc_nhwc_to_nchw: bad arguments");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in_hi,
                                                                (const __nv_bfloat16*)in_lo, out, C, HW);
  count_launch();
  return check_launch("tc_nhwc_to_nchw");
}

int b2c_tc_tiles_per_image(int kind, int Hin, int Win) {
  int Hg, Wg;
  if (kind == B2C_TC_DECONV3X) return b2c::tcp3x::tiles_per_image(Hin, Win);
  switch (kind) {
    case B2C_TC_CONV1: Hg = Hin / 4; Wg = Win / 4; break;
    case B2C_TC_CONV5S2: Hg = (Hin - 1) / 2 + 1; Wg = (Win - 1) / 2 + 1; break;
    case B2C_TC_DECONV5S2:
    case B2C_TC_DECONV3: Hg = Hin; Wg = Win; break;
    default: return 0;
  }
  return ((Wg + TW - 1) / TW) * ((Hg + TH - 1) / TH);
}

int b2c_tc_layer_run(const b2c_tc_layer* L, void* stream_) {
  B2C_REQUIRE(L, "tc_layer_run: null layer");
  B2C_REQUIRE(L->in_hi && L->w_hi && L->out_hi, "tc_layer_run: null tensor pointer");
  B2C_REQUIRE(L->B > 0 && L->B <= 65535 && L->Hin > 0 && L->Win > 0, "tc_layer_run: bad geometry");
  B2C_REQUIRE(!L->split3 || (L->w_lo && (L->kind == B2C_TC_CONV1 || L->in_lo)), "tc_layer_run: split3 needs lo planes");
  B2C_REQUIRE(!L->gdn || (L->gamma_hi && L->beta_eff && (!L->split3 || L->gamma_lo)), "tc_layer_run: gdn needs gamma/beta");
  B2C_REQUIRE(L->out_f32_nchw || !L->split3 || L->out_lo, "tc_layer_run: split3 bf16 output needs out_lo");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (L->kind == B2C_TC_DECONV3X) {
    B2C_REQUIRE((L->channels == 0 || tc_channels_ok(L->channels)) && !L->split3 && !L->gdn && L->out_f32_nchw && !L->saved_y,
                "tc_layer_run: DECONV3X is the plain-bf16 last synthesis layer (128 / 192 / 256 channels), fp32 NCHW output");
    B2C_REQUIRE(!L->ref_image || L->sse_partial, "tc_layer_run: ref_image and sse_partial go together");
    return b2c::tcp3x::launch(L, stream, g_tc_trace);
  }
  {   // persistent kernels for the TMA-fed 128-channel inference layers
    int rc = b2c::tcp::try_launch(L, stream, g_tc_trace);
    if (rc != 1) return rc;
    rc = b2c::tcp1rs::try_launch(L, stream, g_tc_trace);      // CONV1 + GDN1, operands in TMEM, builder / epilogue warps
    if (rc != 1) return rc;
    rc = b2c::tcp1ts::try_launch(L, stream, g_tc_trace);      // CONV1 + GDN1 with both A operands in TMEM
    if (rc != 1) return rc;
    rc = b2c::tcp1::try_launch(L, stream, g_tc_trace);
    if (rc != 1) return rc;
  }
  B2C_REQUIRE(!L->lut, "tc_layer_run: the fused quantizer / entropy epilogue (lut) is implemented by the persistent 128-channel "
                       "CONV5S2 kernel only (split3, gdn 0, out_f32_nchw, B2C_TC_PERSIST on)");
  B2C_REQUIRE(!L->grad_out, "tc_layer_run: grad_out needs the persistent DECONV3 kernel (ref_image set, 128 channels, "
                            "B2C_TC_PERSIST on)");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.B = L->B; p.Hin = L->Hin; p.Win = L->Win;
  p.bias = L->bias; p.beta_eff = L->beta_eff; p.out_hi = L->out_hi; p.out_lo = L->out_lo;
  p.trace = g_tc_trace;
  p.ref_image = L->ref_image; p.sse_partial = L->sse_partial;
  p.saved_y = (const __nv_bfloat16*)L->saved_y; p.saved_n = (const __nv_bfloat16*)L->saved_n;
  B2C_REQUIRE(L->gdn < 3 || (L->saved_y && L->saved_n && !L->split3 && !L->out_f32_nchw),
              "tc_layer_run: GDN backward epilogue needs saved_y/saved_n, bf16 operands and NHWC output");
  B2C_REQUIRE(!L->ref_image || (L->kind == B2C_TC_DECONV3 && L->sse_partial),
              "tc_layer_run: ref_image/sse_partial are for DECONV3 only and go together");
  int phases = 1;
  switch (L->kind) {
    case B2C_TC_CONV1:
      B2C_REQUIRE(L->Hin % 4 == 0 && L->Win % 4 == 0, "tc conv1: H, W must be multiples of 4");
      p.Hg = L->Hin / 4; p.Wg = L->Win / 4; p.ps = 1;
      break;
    case B2C_TC_CONV5S2:
      p.Hg = (L->Hin - 1) / 2 + 1; p.Wg = (L->Win - 1) / 2 + 1; p.ps = 1;
      break;
    case B2C_TC_DECONV5S2:
      p.Hg = L->Hin; p.Wg = L->Win; p.ps = 2; phases = 4;
      break;
    case B2C_TC_DECONV3:
      p.Hg = L->Hin; p.Wg = L->Win; p.ps = 4;
      break;
    default:
      set_error("tc_layer_run: unknown kind %d", L->kind);
      return B2C_EINVAL;
  }
  p.Hfull = p.Hg * p.ps; p.Wfull = p.Wg * p.ps;
  p.tiles_x = (p.Wg + TW - 1) / TW;
  const int tiles_y = (p.Hg + TH - 1) / TH;
  p.tiles_per_image = p.tiles_x * tiles_y;
  p.ksplit = 1;
  if (L->ksplit > 1) {
    B2C_REQUIRE(L->kind == B2C_TC_CONV5S2 && !L->gdn && L->out_f32_nchw && L->ksplit <= 25,
                "tc_layer_run: ksplit is for CONV5S2 with fp32 NCHW output and no GDN (2..25)");
    p.ksplit = L->ksplit;
    phases = L->ksplit;
  }
  dim3 grid(p.tiles_x * tiles_y, phases, L->B);

  const int key = L->kind * 1000 + (L->split3 ? 100 : 0) + L->gdn * 10 + (L->out_f32_nchw ? 1 : 0);
  const int channels = L->channels ? L->channels : 128;
  B2C_REQUIRE(tc_channels_ok(channels), "tc_layer_run: channels must be 128, 192 or 256 (got %d)", channels);
  if (channels != 128) {   // wider latents (config 5): inference kinds only
#define B2C_WIDE(CC)                                                                                                    \
  switch (key) {                                                                                                       \
    case B2C_TC_CONV1 * 1000 + 100 + 10 + 0: return launch<CC, true, B2C_TC_CONV1, 1, OUT_BF16>(L, p, grid, stream);   \
    case B2C_TC_CONV1 * 1000 + 100 + 0 + 1: return launch<CC, true, B2C_TC_CONV1, 0, OUT_F32_NCHW>(L, p, grid, stream); \
    case B2C_TC_CONV5S2 * 1000 + 100 + 10 + 0: return launch<CC, true, B2C_TC_CONV5S2, 1, OUT_BF16, true>(L, p, grid, stream); \
    case B2C_TC_CONV5S2 * 1000 + 100 + 0 + 1: return launch<CC, true, B2C_TC_CONV5S2, 0, OUT_F32_NCHW, true>(L, p, grid, stream); \
    case B2C_TC_DECONV5S2 * 1000 + 0 + 20 + 0: return launch<CC, false, B2C_TC_DECONV5S2, 2, OUT_BF16>(L, p, grid, stream); \
    case B2C_TC_DECONV5S2 * 1000 + 0 + 0 + 1: return launch<CC, false, B2C_TC_DECONV5S2, 0, OUT_F32_NCHW>(L, p, grid, stream); \
    case B2C_TC_DECONV3 * 1000 + 0 + 0 + 1: return launch<CC, false, B2C_TC_DECONV3, 0, OUT_F32_NCHW_PS4>(L, p, grid, stream); \
    default: break;                                                                                                    \
  }
    if (channels == 192) { B2C_WIDE(192) } else { B2C_WIDE(256) }
#undef B2C_WIDE
    set_error("tc_layer_run: unsupported combination for %d channels: kind=%d split3=%d gdn=%d out_f32_nchw=%d", channels,
              L->kind, L->split3, L->gdn, L->out_f32_nchw);
    return B2C_EINVAL;
  }
  static const int use_pairs = [] {   // 0: off, 1: bf16x3 5x5 convs (default), 2: + bf16 transposed convs, 3: + conv1
    const char* e = getenv("B2C_TC_PAIRS");
    return e ? atoi(e) : B2C_TC_PAIRS_DEFAULT;
  }();
  if (use_pairs) {
    switch (key) {
      case B2C_TC_CONV1 * 1000 + 100 + 10 + 0:   // measured slower (0.32 vs 0.28 ms at cfg2): both CTAs of a pair build their
        if (use_pairs > 2 && B2C_CONV1_STAGES == 1)   // A tiles in lock-step and contend for shared memory; opt-in (=3)
          return launch<128, true, B2C_TC_CONV1, 1, OUT_BF16, true>(L, p, grid, stream);
        break;
      case B2C_TC_CONV5S2 * 1000 + 100 + 10 + 0: return launch<128, true, B2C_TC_CONV5S2, 1, OUT_BF16, true>(L, p, grid, stream);
      case B2C_TC_CONV5S2 * 1000 + 100 + 0 + 1: return launch<128, true, B2C_TC_CONV5S2, 0, OUT_F32_NCHW, true>(L, p, grid, stream);
      case B2C_TC_DECONV5S2 * 1000 + 0 + 20 + 0:
        if (use_pairs > 1) return launch<128, false, B2C_TC_DECONV5S2, 2, OUT_BF16, true>(L, p, grid, stream);
        break;   // measured slower than two independent CTAs per SM for the bf16 kinds: opt-in only (B2C_TC_PAIRS=2)
      default: break;
    }
  }
  switch (key) {
    // analysis transform (error-compensated)
    case B2C_TC_CONV1 * 1000 + 100 + 10 + 0: return launch<128, true, B2C_TC_CONV1, 1, OUT_BF16>(L, p, grid, stream);
    case B2C_TC_CONV1 * 1000 + 100 + 0 + 1: return launch<128, true, B2C_TC_CONV1, 0, OUT_F32_NCHW>(L, p, grid, stream);
    case B2C_TC_CONV5S2 * 1000 + 100 + 10 + 0: return launch<128, true, B2C_TC_CONV5S2, 1, OUT_BF16>(L, p, grid, stream);
    case B2C_TC_CONV5S2 * 1000 + 100 + 0 + 1: return launch<128, true, B2C_TC_CONV5S2, 0, OUT_F32_NCHW>(L, p, grid, stream);
    case B2C_TC_CONV5S2 * 1000 + 100 + 10 + 1: return launch<128, true, B2C_TC_CONV5S2, 1, OUT_F32_NCHW>(L, p, grid, stream);
    // bf16 training: forward (saves the norm plane when out_lo is given) and dgrad + GDN-backward epilogues
    case B2C_TC_CONV1 * 1000 + 0 + 10 + 0: return launch<128, false, B2C_TC_CONV1, 1, OUT_BF16>(L, p, grid, stream);
    case B2C_TC_CONV5S2 * 1000 + 0 + 10 + 0: return launch<128, false, B2C_TC_CONV5S2, 1, OUT_BF16>(L, p, grid, stream);
    case B2C_TC_CONV5S2 * 1000 + 0 + 0 + 1: return launch<128, false, B2C_TC_CONV5S2, 0, OUT_F32_NCHW>(L, p, grid, stream);
    case B2C_TC_CONV1 * 1000 + 0 + 40 + 0: return launch<128, false, B2C_TC_CONV1, 4, OUT_BF16>(L, p, grid, stream);
    case B2C_TC_CONV5S2 * 1000 + 0 + 40 + 0: return launch<128, false, B2C_TC_CONV5S2, 4, OUT_BF16>(L, p, grid, stream);
    case B2C_TC_DECONV5S2 * 1000 + 0 + 30 + 0: return launch<128, false, B2C_TC_DECONV5S2, 3, OUT_BF16>(L, p, grid, stream);
    // synthesis transform (plain bf16)
    case B2C_TC_DECONV5S2 * 1000 + 0 + 20 + 0: return launch<128, false, B2C_TC_DECONV5S2, 2, OUT_BF16>(L, p, grid, stream);
    case B2C_TC_DECONV5S2 * 1000 + 0 + 0 + 1: return launch<128, false, B2C_TC_DECONV5S2, 0, OUT_F32_NCHW>(L, p, grid, stream);
    case B2C_TC_DECONV5S2 * 1000 + 0 + 20 + 1: return launch<128, false, B2C_TC_DECONV5S2, 2, OUT_F32_NCHW>(L, p, grid, stream);
    case B2C_TC_DECONV3 * 1000 + 0 + 0 + 1: return launch<128, false, B2C_TC_DECONV3, 0, OUT_F32_NCHW_PS4>(L, p, grid, stream);
    default:
      set_error("tc_layer_run: unsupported combination kind=%d split3=%d gdn=%d out_f32_nchw=%d", L->kind, L->split3,
                L->gdn, L->out_f32_nchw);
      return B2C_EINVAL;
  }
}

}  // extern "C"
