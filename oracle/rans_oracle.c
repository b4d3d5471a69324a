/* TEST INFRASTRUCTURE - CPU oracle of libb2c's latent bitstream (include/b2c.h: b2c_rans_encode / b2c_rans_decode).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * PARITY UNPINNED: the reference has no bitstream (it estimates bits, model.py:71-74, and approximates file sizes with
 * gzip, NewTests/testModelRecon.py:80), so there is no reference output to pin against.  This file restates the
 * published algorithm - range asymmetric numeral systems (J. Duda, "Asymmetric numeral systems: entropy coding
 * combining speed of Huffman coding with compression rate of arithmetic coding", 2013), in the word-renormalised,
 * interleaved form of F. Giesen, "Interleaved entropy coders" (2014): state x in [2^16, 2^32), 16-bit words,
 *      encode(s):  while x >= 2^(32-prec) * f[s]: emit x & 0xffff, x >>= 16 ;  x = (x / f[s]) << prec | + x % f[s] + start[s]
 *      decode:     slot = x & (2^prec - 1) ; s: start[s] <= slot < start[s+1] ; x = f[s] * (x >> prec) + slot - start[s] ;
 *                  if x < 2^16: x = x << 16 | next word
 * and the container of csrc/b2c_rans.cu, written here as plain sequential loops over 32 lane states: lane l of a block
 * codes the symbols at block_begin + 32 * step + l; within a step the words of the renormalising lanes are stored in
 * lane order; an escaped value (|z| > L) is the escape symbol followed by 16 raw bits (one forced word); a block =
 * 32 final states (little endian) + words.  What pins it: encode -> decode round trips, a hand-checked one-symbol
 * known answer (tests/test_oracle_rans.py) and the size against the reference's own estimate of the same latents.
 *
 *   gcc -O2 -shared -fPIC oracle/rans_oracle.c -o oracle/_build/librans_oracle.so      (done by __graft_entry__.build) */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define LOW (1u << 16)
enum { ST_RANGE = 1, ST_CAPACITY = 2, ST_CORRUPT = 4 };

static int to_int(float v, int* status) {
  float r = nearbyintf(v); /* round half to even, like torch.round (model.py:56) */
  if (!(fabsf(r) <= 32767.f)) {
    *status |= ST_RANGE;
    if (r != r) return 0;
    return r > 0 ? 32767 : -32767;
  }
  return (int)r;
}

/* one block; words are produced in REVERSE decode order into tmp (so tmp is reversed at the end) */
static size_t encode_block(const float* z, size_t begin, size_t end, const uint32_t* freq_cum, int L, int prec, int HW,
                           uint16_t* tmp, int* status) {
  const int A = 2 * L + 2;
  uint32_t x[32];
  size_t nw = 0;
  for (int l = 0; l < 32; ++l) x[l] = LOW;
  const long steps = (long)((end - begin + 31) / 32);
  for (long j = steps - 1; j >= 0; --j) {
    int zi[32], valid[32];
    for (int l = 0; l < 32; ++l) {
      size_t i = begin + (size_t)j * 32 + l;
      valid[l] = i < end;
      zi[l] = valid[l] ? to_int(z[i], status) : 0;
    }
    /* raw bits of escapes: the decoder reads these AFTER the step's renormalisation words, lanes ascending, so the
       reversed writer emits them first, lanes descending */
    for (int l = 31; l >= 0; --l)
      if (valid[l] && (zi[l] > L || zi[l] < -L)) {
        tmp[nw++] = (uint16_t)x[l];
        x[l] = ((x[l] >> 16) << 16) | (uint32_t)(zi[l] + 32768);
      }
    for (int l = 31; l >= 0; --l)
      if (valid[l]) {
        size_t i = begin + (size_t)j * 32 + l;
        int c = (int)(i / (size_t)HW);
        int s = (zi[l] > L || zi[l] < -L) ? A - 1 : zi[l] + L;
        uint32_t fc = freq_cum[(size_t)c * A + s], f = fc & 0xffff, start = fc >> 16;
        if ((uint64_t)x[l] >= ((uint64_t)f << (32 - prec))) {
          tmp[nw++] = (uint16_t)x[l];
          x[l] >>= 16;
        }
        x[l] = ((x[l] / f) << prec) + (x[l] % f) + start;
      }
  }
  for (int l = 31; l >= 0; --l) { /* states: lane 0 first in the stream, low half first */
    tmp[nw++] = (uint16_t)(x[l] >> 16);
    tmp[nw++] = (uint16_t)x[l];
  }
  return nw;
}

int rans_oracle_encode(const float* z, const uint32_t* freq_cum, int L, int prec, int B, int C, int HW, int BS,
                       uint32_t* block_bytes, uint64_t* offsets, uint8_t* out, size_t capacity, int* status) {
  const size_t T = (size_t)C * HW;
  const int nblk = (int)((T + BS - 1) / BS);
  uint16_t* tmp = (uint16_t*)malloc((2 * (size_t)BS + 64) * sizeof(uint16_t));
  if (!tmp) return -1;
  uint64_t off = 0;
  for (int n = 0; n < B; ++n)
    for (int b = 0; b < nblk; ++b) {
      size_t begin = (size_t)b * BS, end = begin + BS < T ? begin + BS : T;
      size_t nw = encode_block(z + (size_t)n * T, begin, end, freq_cum, L, prec, HW, tmp, status);
      offsets[n * nblk + b] = off;
      block_bytes[n * nblk + b] = (uint32_t)(2 * nw);
      if (off + 2 * nw <= capacity) {
        for (size_t k = 0; k < nw; ++k) { /* reverse into decode order, little-endian words */
          uint16_t w = tmp[nw - 1 - k];
          out[off + 2 * k] = (uint8_t)(w & 0xff);
          out[off + 2 * k + 1] = (uint8_t)(w >> 8);
        }
      } else {
        *status |= ST_CAPACITY;
      }
      off += 2 * nw;
    }
  offsets[(size_t)B * nblk] = off;
  free(tmp);
  return 0;
}

int rans_oracle_decode(const uint8_t* in, const uint64_t* offsets, const uint32_t* freq_cum, int L, int prec, int B, int C,
                       int HW, int BS, float* z, int* status) {
  const size_t T = (size_t)C * HW;
  const int nblk = (int)((T + BS - 1) / BS), A = 2 * L + 2;
  const uint32_t M = 1u << prec;
  for (int n = 0; n < B; ++n)
    for (int b = 0; b < nblk; ++b) {
      const uint8_t* w = in + offsets[n * nblk + b];
      const size_t len = (size_t)(offsets[n * nblk + b + 1] - offsets[n * nblk + b]) / 2;
      size_t begin = (size_t)b * BS, end = begin + BS < T ? begin + BS : T, rp = 64;
      uint32_t x[32];
      if (len < 64) { *status |= ST_CORRUPT; continue; }
#define WORD(k) ((uint32_t)w[2 * (k)] | ((uint32_t)w[2 * (k) + 1] << 8))
      for (int l = 0; l < 32; ++l) x[l] = WORD(2 * l) | (WORD(2 * l + 1) << 16);
      const long steps = (long)((end - begin + 31) / 32);
      for (long j = 0; j < steps; ++j) {
        int sym[32], valid[32];
        for (int l = 0; l < 32; ++l) {
          size_t i = begin + (size_t)j * 32 + l;
          valid[l] = i < end;
          if (!valid[l]) continue;
          int c = (int)(i / (size_t)HW);
          const uint32_t* row = freq_cum + (size_t)c * A;
          uint32_t slot = x[l] & (M - 1);
          int s = 0;
          while (s + 1 < A && (row[s + 1] >> 16) <= slot) ++s; /* start[s] <= slot < start[s+1] */
          sym[l] = s;
          x[l] = (row[s] & 0xffff) * (x[l] >> prec) + slot - (row[s] >> 16);
          if (x[l] < LOW) {
            if (rp < len) x[l] = (x[l] << 16) | WORD(rp); else *status |= ST_CORRUPT;
            ++rp;
          }
        }
        for (int l = 0; l < 32; ++l) {
          if (!valid[l]) continue;
          size_t i = begin + (size_t)j * 32 + l;
          int zi = sym[l] - L;
          if (sym[l] == A - 1) {
            zi = (int)(x[l] & 0xffff) - 32768;
            uint32_t word = 0;
            if (rp < len) word = WORD(rp); else *status |= ST_CORRUPT;
            ++rp;
            x[l] = ((x[l] >> 16) << 16) | word;
          }
          z[(size_t)n * T + i] = (float)zi;
        }
      }
#undef WORD
      for (int l = 0; l < 32; ++l)
        if (x[l] != LOW) *status |= ST_CORRUPT;
      if (rp != len) *status |= ST_CORRUPT;
    }
  return 0;
}
