// Real bitstream for the integer latents: a range-ANS coder driven by the BitEstimator's per-channel
// CDF.  The reference only ESTIMATES the size (model.py:71-74: bits = sum -log2(CDF(z+.5) - CDF(z-.5))); this turns the
// same probabilities into bytes and back.  No reference bitstream exists, so the format is this library's own:
//
//   * per channel c an alphabet of A = 2L+2 symbols: z in [-L, L] -> z + L, and one ESCAPE symbol (index 2L+1) followed
//     by the value as 16 raw bits (z + 32768), for the tails;
//   * frequencies f[c][s] >= 1 with sum 2^prec, f = 1 + floor(p * (2^prec - A)), the remainder given to the most
//     probable symbol; p(z) = CDF(z+.5) - CDF(z-.5) with the Bitparm chain of bitEstimator.py:20-25,38-42 in exact fp32;
//   * 32-bit rANS states, 16-bit renormalisation words (state in [2^16, 2^32)), 32 INTERLEAVED states = the 32 lanes of
//     one warp, all writing ONE shared word stream: in a step every lane codes the symbol at index base + 32*step + lane
//     (coalesced 128-byte reads of z) and the lanes that renormalise place their words in lane order (ballot + popc).
//     The encoder runs the steps backwards and fills its buffer from the end, so the decoder reads forwards;
//   * one warp codes one BLOCK of `block_symbols` consecutive latents (NCHW order) of one image; a block's bytes are the
//     32 final states (lane order, little endian) followed by the words.  Blocks are independent -> the grid is
//     images x blocks; per block the cost is 128 bytes of states, so block_symbols trades parallelism for size.
//
// Kernels: table build (one CTA per channel), encode (one warp per block, into a worst-case-sized scratch slot),
// exclusive scan of the block sizes, gather into the contiguous output; decode (one warp per block).  Integer work:
// bit-exact against oracle/rans_oracle.c (tests/test_gpu_rans.py).
#include "b2c_common.cuh"

namespace b2c {

constexpr uint32_t RANS_LOW = 1u << 16;

struct BitPtrs11 {
  const float* p[B2C_NUM_BITPARAMS];
};

__device__ __forceinline__ float rs_softplus(float h) { return h > 20.f ? h : log1pf(expf(h)); }
__device__ __forceinline__ float rs_sigmoid(float t) { return 1.f / (1.f + expf(-t)); }

// CDF of bitEstimator.py:38-42 with k = {sp1,b1,ta1, sp2,b2,ta2, sp3,b3,ta3, sp4,b4}
__device__ __forceinline__ float rs_cdf(const float* k, float v) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float t = fmaf(v, k[3 * i], k[3 * i + 1]);
    v = fmaf(tanhf(t), k[3 * i + 2], t);
  }
  return rs_sigmoid(fmaf(v, k[9], k[10]));
}

// one CTA per channel; freq_cum[c][s] = f | (start << 16) ; slot_sym[c][slot] = s
__global__ void __launch_bounds__(256) rans_tables_kernel(BitPtrs11 bp, int L, int prec, uint32_t* __restrict__ freq_cum,
                                                          uint8_t* __restrict__ slot_sym) {
  __shared__ float k[12];
  __shared__ uint32_t f[256];
  __shared__ uint32_t start[257];
  const int c = blockIdx.x, A = 2 * L + 2;
  const uint32_t M = 1u << prec;
  if (threadIdx.x < 11) {
    const int i = threadIdx.x;
    const float raw = __ldg(bp.p[i] + c);
    const bool is_h = (i < 9 && i % 3 == 0) || i == 9, is_a = (i < 9 && i % 3 == 2);
    k[i] = is_h ? rs_softplus(raw) : (is_a ? tanhf(raw) : raw);
  }
  __syncthreads();
  const float scale = (float)(M - A);
  for (int s = threadIdx.x; s < A; s += blockDim.x) {
    float p;
    if (s < A - 1) {
      const float zv = (float)(s - L);
      p = rs_cdf(k, zv + 0.5f) - rs_cdf(k, zv - 0.5f);
    } else {                                                  // escape: both tails
      p = rs_cdf(k, -(float)L - 0.5f) + (1.f - rs_cdf(k, (float)L + 0.5f));
    }
    p = fminf(fmaxf(p, 0.f), 1.f);
    f[s] = 1u + (uint32_t)floorf(p * scale);
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                     // A <= 254: a serial fix-up is nothing
    uint32_t total = 0, best = 0;
    for (int s = 0; s < A; ++s) {
      total += f[s];
      if (f[s] > f[best]) best = s;
    }
    if (total <= M) f[best] += M - total;
    else f[best] -= total - M;                                // only from fp32 rounding of sum p > 1: a few counts
    uint32_t acc = 0;
    for (int s = 0; s < A; ++s) {
      start[s] = acc;
      acc += f[s];
    }
    start[A] = acc;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < A; s += blockDim.x) freq_cum[(size_t)c * A + s] = f[s] | (start[s] << 16);
  for (int s = 0; s < A; ++s)
    for (uint32_t slot = start[s] + threadIdx.x; slot < start[s + 1]; slot += blockDim.x)
      slot_sym[(size_t)c * M + slot] = (uint8_t)s;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// status bits
constexpr int RANS_U = 4;        // encoder: steps per software-pipeline group
constexpr int RS_RANGE = 1;      // |z| > 32767 (clamped) or non-finite
constexpr int RS_CAPACITY = 2;   // output buffer too small
constexpr int RS_CORRUPT = 4;    // decoder: final state / length mismatch or read past the block

__global__ void __launch_bounds__(32) rans_encode_kernel(const float* __restrict__ z, const uint32_t* __restrict__ freq_cum,
                                                         int L, int prec, int C, int HW, int BS, int nblk,
                                                         uint16_t* __restrict__ scratch, uint32_t cap_words,
                                                         uint32_t* __restrict__ block_bytes, int* __restrict__ status) {
  const unsigned blk = blockIdx.x, n = blk / nblk, b = blk - n * nblk, lane = threadIdx.x;
  const unsigned T = (unsigned)C * HW, begin = b * BS, end = min(begin + (unsigned)BS, T);
  const int steps = (int)((end - begin + 31) / 32), A = 2 * L + 2;
  uint16_t* out = scratch + (size_t)blk * cap_words;
  const float* zp = z + (size_t)n * T;
  uint32_t ptr = cap_words, x = RANS_LOW;
  const uint32_t lt = lanemask_lt();
  bool bad = false;
  // Software pipeline over groups of RANS_U steps (the steps run downwards).  Nothing a step LOADS depends on the coder
  // state: the latent, and through its symbol the table entry, are known ahead.  While group g is coded, the table
  // entries of group g+1 and the latents of group g+2 are in flight, so the serial chain per step is ALU only (ncu of
  // the unpipelined loop: 56 % of all stall samples waiting on the table entry, one L2 latency per step).
  auto load_v = [&](int j) -> float {
    if (j < 0) return 0.f;
    const unsigned i = begin + (unsigned)j * 32 + lane;
    return i < end ? __ldg(zp + i) : 0.f;
  };
  // latent of step j -> (value, flags bit 0 valid, bit 1 escape), table entry load issued
  auto classify = [&](int j, float v, int& zi, uint32_t& flags, uint32_t& fc) {
    zi = 0; flags = 0; fc = 1;
    if (j < 0) return;
    const unsigned i = begin + (unsigned)j * 32 + lane;
    const bool valid = i < end;
    if (valid) {
      const float r = rintf(v);
      if (!(fabsf(r) <= 32767.f)) { bad = true; zi = (r > 0.f) ? 32767 : -32767; if (r != r) zi = 0; }
      else zi = (int)r;
    }
    const bool esc = valid && (zi > L || zi < -L);
    const int sym = esc ? A - 1 : zi + L;
    const unsigned c = valid ? i / (unsigned)HW : 0;
    fc = __ldg(freq_cum + (size_t)c * A + sym);
    flags = (valid ? 1u : 0u) | (esc ? 2u : 0u);
  };
  float vn[RANS_U];
  int zi_c[RANS_U], zi_n[RANS_U];
  uint32_t fl_c[RANS_U], fl_n[RANS_U], fc_c[RANS_U], fc_n[RANS_U];
  int jg = steps - 1;
#pragma unroll
  for (int u = 0; u < RANS_U; ++u) vn[u] = load_v(jg - u);
#pragma unroll
  for (int u = 0; u < RANS_U; ++u) classify(jg - u, vn[u], zi_c[u], fl_c[u], fc_c[u]);
#pragma unroll
  for (int u = 0; u < RANS_U; ++u) vn[u] = load_v(jg - RANS_U - u);
  for (; jg >= 0; jg -= RANS_U) {
#pragma unroll
    for (int u = 0; u < RANS_U; ++u) classify(jg - RANS_U - u, vn[u], zi_n[u], fl_n[u], fc_n[u]);
#pragma unroll
    for (int u = 0; u < RANS_U; ++u) vn[u] = load_v(jg - 2 * RANS_U - u);
#pragma unroll
    for (int u = 0; u < RANS_U; ++u) {
      if (jg - u < 0) break;                                  // warp-uniform
      const bool valid = fl_c[u] & 1u, esc = fl_c[u] & 2u;
      const int zi = zi_c[u];
      const uint32_t fc = fc_c[u];
      // (B') raw 16 bits of an escaped value: always exactly one word out, then x = (x >> 16) << 16 | raw
      uint32_t m = __ballot_sync(0xffffffffu, esc);
      if (m) {
        const uint32_t cnt = __popc(m);
        if (esc) {
          out[ptr - cnt + __popc(m & lt)] = (uint16_t)x;
          x = (x & 0xffff0000u) | (uint32_t)(zi + 32768);
        }
        ptr -= cnt;
      }
      // (A') renormalise, then code the symbol
      const uint32_t f = fc & 0xffffu, st = fc >> 16;
      const uint32_t xmax = ((RANS_LOW >> prec) << 16) * f;   // f <= 2^prec - A + 1: no overflow
      const bool emit = valid && x >= xmax;
      m = __ballot_sync(0xffffffffu, emit);
      if (m) {
        const uint32_t cnt = __popc(m);
        if (emit) {
          out[ptr - cnt + __popc(m & lt)] = (uint16_t)x;
          x >>= 16;
        }
        ptr -= cnt;
      }
      if (valid) {
        const uint32_t q = x / f;
        x = (q << prec) + (x - q * f) + st;
      }
    }
#pragma unroll
    for (int u = 0; u < RANS_U; ++u) { zi_c[u] = zi_n[u]; fl_c[u] = fl_n[u]; fc_c[u] = fc_n[u]; }
  }
  ptr -= 64;
  out[ptr + 2 * lane] = (uint16_t)x;
  out[ptr + 2 * lane + 1] = (uint16_t)(x >> 16);
  if (lane == 0) block_bytes[blk] = (cap_words - ptr) * 2u;
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, RS_RANGE);
}

// offsets[0..n] = exclusive scan of block_bytes[0..n) (bytes), one CTA
__global__ void __launch_bounds__(1024) rans_scan_kernel(const uint32_t* __restrict__ block_bytes, int n,
                                                         unsigned long long* __restrict__ offsets,
                                                         unsigned long long capacity, int* __restrict__ status) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x, per = (n + 1023) / 1024;
  const int lo = min(t * per, n), hi = min(lo + per, n);
  unsigned long long s = 0;
  for (int i = lo; i < hi; ++i) s += block_bytes[i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    unsigned long long acc = 0;
    for (int i = 0; i < 1024; ++i) {
      const unsigned long long v = part[i];
      part[i] = acc;
      acc += v;
    }
    offsets[n] = acc;
    if (acc > capacity) atomicOr(status, RS_CAPACITY);
  }
  __syncthreads();
  unsigned long long acc = part[t];
  for (int i = lo; i < hi; ++i) {
    offsets[i] = acc;
    acc += block_bytes[i];
  }
}

__global__ void __launch_bounds__(128) rans_gather_kernel(const uint16_t* __restrict__ scratch, uint32_t cap_words,
                                                          const uint32_t* __restrict__ block_bytes,
                                                          const unsigned long long* __restrict__ offsets,
                                                          uint16_t* __restrict__ out, unsigned long long capacity) {
  const unsigned blk = blockIdx.x;
  const uint32_t words = block_bytes[blk] / 2;
  const unsigned long long off = offsets[blk];
  if (off + 2ull * words > capacity) return;                  // RS_CAPACITY already flagged by the scan
  const uint16_t* src = scratch + (size_t)blk * cap_words + (cap_words - words);
  uint16_t* dst = out + off / 2;
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(32) rans_decode_kernel(const uint16_t* __restrict__ in,
                                                         const unsigned long long* __restrict__ offsets,
                                                         const uint32_t* __restrict__ freq_cum,
                                                         const uint8_t* __restrict__ slot_sym, int L, int prec, int C, int HW,
                                                         int BS, int nblk, float* __restrict__ z, int* __restrict__ status) {
  const unsigned blk = blockIdx.x, n = blk / nblk, b = blk - n * nblk, lane = threadIdx.x;
  const unsigned T = (unsigned)C * HW, begin = b * BS, end = min(begin + (unsigned)BS, T);
  const int steps = (int)((end - begin + 31) / 32), A = 2 * L + 2;
  const uint32_t M = 1u << prec, lt = lanemask_lt();
  const uint16_t* w = in + offsets[blk] / 2;
  const uint32_t len = (uint32_t)((offsets[blk + 1] - offsets[blk]) / 2);
  float* zp = z + (size_t)n * T;
  bool bad = len < 64;
  uint32_t x = bad ? RANS_LOW : ((uint32_t)w[2 * lane] | ((uint32_t)w[2 * lane + 1] << 16));
  uint32_t rp = 64;
  // Register window over the word stream: wa = w[base + lane], wb = w[base + 32 + lane] with base <= rp < base + 32, so
  // a renormalising lane takes its word with a shuffle instead of a dependent global load; the refill (one coalesced
  // 64-byte load per 32 consumed words) is issued many steps before its words are needed.
  uint32_t base = 64;
  auto load_w = [&](uint32_t at) -> uint32_t { return at < len ? (uint32_t)w[at] : 0u; };
  uint32_t wa = load_w(base + lane), wb = load_w(base + 32 + lane);
  auto fetch = [&](uint32_t at) -> uint32_t {                 // called by all lanes; base <= at < base + 64
    const uint32_t idx = at - base;
    const uint32_t a = __shfl_sync(0xffffffffu, wa, idx & 31), b2 = __shfl_sync(0xffffffffu, wb, idx & 31);
    return idx < 32 ? a : b2;
  };
  auto consume = [&](uint32_t cnt) {                          // warp-uniform; cnt <= 32
    rp += cnt;
    if (rp - base >= 32) {
      wa = wb;
      base += 32;
      wb = load_w(base + 32 + lane);
    }
  };
  for (int j = 0; j < steps; ++j) {
    const unsigned i = begin + (unsigned)j * 32 + lane;
    const bool valid = i < end;
    const unsigned c = valid ? i / (unsigned)HW : 0;
    int s = 0;
    if (valid) {
      const uint32_t slot = x & (M - 1);
      s = __ldg(slot_sym + (size_t)c * M + slot);
      const uint32_t fc = __ldg(freq_cum + (size_t)c * A + s);
      x = (fc & 0xffffu) * (x >> prec) + slot - (fc >> 16);
    }
    const bool need = valid && x < RANS_LOW;
    uint32_t m = __ballot_sync(0xffffffffu, need);
    if (m) {
      const uint32_t at = rp + __popc(m & lt);
      const uint32_t word = fetch(at);
      if (need) {
        if (at >= len) bad = true;
        x = (x << 16) | word;
      }
      consume(__popc(m));
    }
    const bool esc = valid && s == A - 1;
    int zi = s - L;
    m = __ballot_sync(0xffffffffu, esc);
    if (m) {
      const uint32_t at = rp + __popc(m & lt);
      const uint32_t word = fetch(at);
      if (esc) {
        zi = (int)(x & 0xffffu) - 32768;
        if (at >= len) bad = true;
        x = (x & 0xffff0000u) | word;
      }
      consume(__popc(m));
    }
    if (valid) zp[i] = (float)zi;
  }
  bad = bad || x != RANS_LOW || rp != len;
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, RS_CORRUPT);
}

static bool rans_args_ok(int half_range, int prec, int B, int C, int HW, int block_symbols) {
  return half_range >= 1 && half_range <= 126 && prec >= 10 && prec <= 15 && (1 << prec) >= 4 * (2 * half_range + 2) &&
         B > 0 && C > 0 && HW > 0 && block_symbols >= 32 && block_symbols % 32 == 0 &&
         (unsigned long long)C * HW < (1ull << 31) && block_symbols <= (1 << 24);
}
static int rans_blocks(int C, int HW, int block_symbols) {
  return (int)(((long long)C * HW + block_symbols - 1) / block_symbols);
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_rans_build_tables(const float* const* host_params, int C, int half_range, int prec,
                                     uint32_t* freq_cum, uint8_t* slot_sym, void* stream) {
  B2C_REQUIRE(host_params && freq_cum && slot_sym && rans_args_ok(half_range, prec, 1, C, 1, 32),
              "rans_build_tables: bad arguments (1 <= half_range <= 126, 10 <= prec <= 15, 2^prec >= 4 * (2 * half_range + 2))");
  BitPtrs11 bp;
  for (int i = 0; i < B2C_NUM_BITPARAMS; ++i) {
    B2C_REQUIRE(host_params[i], "rans_build_tables: null parameter pointer %d", i);
    bp.p[i] = host_params[i];
  }
  rans_tables_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(bp, half_range, prec, freq_cum, slot_sym);
  count_launch();
  return check_launch("rans_build_tables");
}

extern "C" int b2c_rans_num_blocks(int C, int HW, int block_symbols) {
  if (C <= 0 || HW <= 0 || block_symbols < 32) return 0;
  return rans_blocks(C, HW, block_symbols);
}

extern "C" size_t b2c_rans_scratch_bytes(int B, int C, int HW, int block_symbols) {
  if (!rans_args_ok(1, 10, B, C, HW, block_symbols)) return 0;
  return (size_t)B * rans_blocks(C, HW, block_symbols) * (2 * (size_t)block_symbols + 64) * sizeof(uint16_t);
}

extern "C" int b2c_rans_encode(const float* z, const uint32_t* freq_cum, int half_range, int prec, int B, int C, int HW,
                               int block_symbols, void* scratch, size_t scratch_bytes, uint32_t* block_bytes,
                               uint64_t* offsets, uint8_t* out, size_t out_capacity, int* status, void* stream) {
  B2C_REQUIRE(z && freq_cum && scratch && block_bytes && offsets && out && status &&
                  rans_args_ok(half_range, prec, B, C, HW, block_symbols),
              "rans_encode: bad arguments");
  B2C_REQUIRE(((uintptr_t)out & 1) == 0 && ((uintptr_t)scratch & 1) == 0, "rans_encode: buffers must be 2-byte aligned");
  if (scratch_bytes < b2c_rans_scratch_bytes(B, C, HW, block_symbols)) {
    set_error("rans_encode: scratch too small (%zu < %zu)", scratch_bytes, b2c_rans_scratch_bytes(B, C, HW, block_symbols));
    return B2C_EWORKSPACE;
  }
  const int nblk = rans_blocks(C, HW, block_symbols), total = B * nblk;
  const uint32_t cap_words = 2u * block_symbols + 64u;
  cudaStream_t s = (cudaStream_t)stream;
  rans_encode_kernel<<<total, 32, 0, s>>>(z, freq_cum, half_range, prec, C, HW, block_symbols, nblk, (uint16_t*)scratch,
                                          cap_words, block_bytes, status);
  rans_scan_kernel<<<1, 1024, 0, s>>>(block_bytes, total, (unsigned long long*)offsets, out_capacity, status);
  rans_gather_kernel<<<total, 128, 0, s>>>((const uint16_t*)scratch, cap_words, block_bytes,
                                           (const unsigned long long*)offsets, (uint16_t*)out, out_capacity);
  count_launch(3);
  return check_launch("rans_encode");
}

extern "C" int b2c_rans_decode(const uint8_t* in, const uint64_t* offsets, const uint32_t* freq_cum, const uint8_t* slot_sym,
                               int half_range, int prec, int B, int C, int HW, int block_symbols, float* z, int* status,
                               void* stream) {
  B2C_REQUIRE(in && offsets && freq_cum && slot_sym && z && status && rans_args_ok(half_range, prec, B, C, HW, block_symbols),
              "rans_decode: bad arguments");
  B2C_REQUIRE(((uintptr_t)in & 1) == 0, "rans_decode: input must be 2-byte aligned");
  const int nblk = rans_blocks(C, HW, block_symbols);
  rans_decode_kernel<<<B * nblk, 32, 0, (cudaStream_t)stream>>>((const uint16_t*)in, (const unsigned long long*)offsets,
                                                                freq_cum, slot_sym, half_range, prec, C, HW, block_symbols,
                                                                nblk, z, status);
  count_launch();
  return check_launch("rans_decode");
}
