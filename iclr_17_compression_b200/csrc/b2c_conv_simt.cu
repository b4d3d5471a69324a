// Exact-fp32 convolution family (NCHW, the reference's own layout):
//   conv2d forward            models/analysis_17.py:14,18,22  (nn.Conv2d)
//   conv_transpose2d forward  models/synthesis_17.py:15,19,23 (nn.ConvTranspose2d), sub-pixel phase form
//   weight gradient           autograd of both (train.py:105)
//   channel sum               bias gradients
// All are implicit GEMMs on the shared 128x128x8 fp32 tile core (b2c_common.cuh); operands are
// gathered on the fly, nothing is materialised.  These kernels are the bit-faithful path
// (fp32 FMA accumulation); the bf16 tcgen05 kernels in b2c_conv_tc.cu are the fast path.
#include "b2c_common.cuh"

namespace b2c {

// ----------------------------------------------------------------------------- conv2d fwd
struct ConvGeom {
  int B, Cin, H, W, Cout, Ho, Wo, ks, stride, pad, M, K;
};

struct ConvA {  // A[m][k] = x[n, c, oy*s-p+ky, ox*s-p+kx],  m=(n,oy,ox), k=(c,ky,kx)
  const float* x;
  ConvGeom g;
  const float* base;
  int iy0, ix0, c, ky, kx, k;
  bool valid;
  __device__ __forceinline__ void init(int m, int k_first) {
    valid = m < g.M;
    const int hw = g.Ho * g.Wo;
    const int n = valid ? m / hw : 0;
    const int r = valid ? m - n * hw : 0;
    const int oy = r / g.Wo, ox = r - oy * g.Wo;
    iy0 = oy * g.stride - g.pad;
    ix0 = ox * g.stride - g.pad;
    base = x + (size_t)n * g.Cin * g.H * g.W;
    k = k_first;
    const int kk = g.ks * g.ks;
    c = k / kk;
    const int r2 = k - c * kk;
    ky = r2 / g.ks;
    kx = r2 - ky * g.ks;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int iy = iy0 + ky, ix = ix0 + kx;
      const bool ok = valid && k < g.K && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      v[j] = ok ? __ldg(base + ((size_t)c * g.H + iy) * g.W + ix) : 0.f;
      ++k;
      if (++kx == g.ks) {
        kx = 0;
        if (++ky == g.ks) { ky = 0; ++c; }
      }
    }
    k += 4;
    kx += 4;
    while (kx >= g.ks) { kx -= g.ks; ++ky; }
    while (ky >= g.ks) { ky -= g.ks; ++c; }
  }
};

struct ConvB {  // B[k][o] = w[o][k]  (OIHW is K-contiguous per output channel)
  const float* w;
  int Cout, K;
  const float* p;
  int k;
  bool valid;
  __device__ __forceinline__ void init(int o, int k_first) {
    valid = o < Cout;
    p = w + (size_t)(valid ? o : 0) * K;
    k = k_first;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (valid && k + j < K) ? __ldg(p + k + j) : 0.f;
    k += 8;
  }
};

__global__ void __launch_bounds__(TG_THREADS) conv2d_fwd_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ y, ConvGeom g) {
  ConvA la{x, g};
  ConvB lb{w, g.Cout, g.K};
  TileAcc acc;
  tile_gemm(la, lb, blockIdx.x * TG_BM, blockIdx.y * TG_BN, 0, g.K, acc);
  const int hw = g.Ho * g.Wo;
  const bool vec = (hw & 3) == 0;
#pragma unroll
  for (int ih = 0; ih < 2; ++ih) {
    const int m = acc.m(ih * 4);
    if (m >= g.M) continue;
    const int n = m / hw, r = m - n * hw;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = acc.n(j);
      if (o >= g.Cout) continue;
      const float bv = bias ? __ldg(bias + o) : 0.f;
      float* dst = y + ((size_t)n * g.Cout + o) * hw + r;
      if (vec) {  // 4 consecutive m share n and are contiguous
        float4 o4 = make_float4(acc.v[ih * 4 + 0][j] + bv, acc.v[ih * 4 + 1][j] + bv,
                                acc.v[ih * 4 + 2][j] + bv, acc.v[ih * 4 + 3][j] + bv);
        *reinterpret_cast<float4*>(dst) = o4;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int mi = m + i;
          if (mi >= g.M) break;
          const int ni = mi / hw, ri = mi - ni * hw;
          y[((size_t)ni * g.Cout + o) * hw + ri] = acc.v[ih * 4 + i][j] + bv;
        }
      }
    }
  }
}

// ------------------------------------------------------------------- conv_transpose2d fwd
// out[n,o,oy,ox] = b[o] + sum_{i,ky,kx: oy = s*iy - p + ky} x[n,i,iy,ix] * W[i,o,ky,kx]
// Phase (ry,rx) = (oy % s, ox % s): ky = ky0 + s*jy with ky0 = (ry+p) % s, iy = my + dy0 - jy,
// dy0 = (ry+p)/s, oy = s*my + ry.  Each phase is a dense stride-1 correlation.
struct ConvTGeom {
  int B, Cin, H, W, Cout, Ho, Wo, ks, stride, pad;
};
struct PhaseGeom {
  int ry, rx, ky0, kx0, nty, ntx, dy0, dx0, Hm, Wm, M, K;
};
__device__ __forceinline__ PhaseGeom make_phase(const ConvTGeom& g, int phase) {
  PhaseGeom p;
  p.ry = phase / g.stride;
  p.rx = phase - p.ry * g.stride;
  p.ky0 = (p.ry + g.pad) % g.stride;
  p.kx0 = (p.rx + g.pad) % g.stride;
  p.nty = (g.ks - p.ky0 + g.stride - 1) / g.stride;
  p.ntx = (g.ks - p.kx0 + g.stride - 1) / g.stride;
  p.dy0 = (p.ry + g.pad) / g.stride;
  p.dx0 = (p.rx + g.pad) / g.stride;
  p.Hm = (g.Ho - p.ry + g.stride - 1) / g.stride;
  p.Wm = (g.Wo - p.rx + g.stride - 1) / g.stride;
  p.M = g.B * p.Hm * p.Wm;
  p.K = g.Cin * p.nty * p.ntx;
  return p;
}

struct ConvTA {  // A[m][k] = x[n, i, my+dy0-jy, mx+dx0-jx], m=(n,my,mx), k=(i,jy,jx)
  const float* x;
  ConvTGeom g;
  PhaseGeom p;
  const float* base;
  int y0, x0, i, jy, jx, k;
  bool valid;
  __device__ __forceinline__ void init(int m, int k_first) {
    valid = m < p.M;
    const int hw = p.Hm * p.Wm;
    const int n = valid ? m / hw : 0;
    const int r = valid ? m - n * hw : 0;
    const int my = r / p.Wm, mx = r - my * p.Wm;
    y0 = my + p.dy0;
    x0 = mx + p.dx0;
    base = x + (size_t)n * g.Cin * g.H * g.W;
    k = k_first;
    const int tt = p.nty * p.ntx;
    i = k / tt;
    const int r2 = k - i * tt;
    jy = r2 / p.ntx;
    jx = r2 - jy * p.ntx;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int iy = y0 - jy, ix = x0 - jx;
      const bool ok = valid && k < p.K && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      v[j] = ok ? __ldg(base + ((size_t)i * g.H + iy) * g.W + ix) : 0.f;
      ++k;
      if (++jx == p.ntx) {
        jx = 0;
        if (++jy == p.nty) { jy = 0; ++i; }
      }
    }
    k += 4;
    jx += 4;
    while (jx >= p.ntx) { jx -= p.ntx; ++jy; }
    while (jy >= p.nty) { jy -= p.nty; ++i; }
  }
};

struct ConvTB {  // B[k][o] = W[i, o, ky0+s*jy, kx0+s*jx]
  const float* w;
  ConvTGeom g;
  PhaseGeom p;
  int o, i, jy, jx, k;
  bool valid;
  __device__ __forceinline__ void init(int o_, int k_first) {
    o = o_;
    valid = o < g.Cout;
    k = k_first;
    const int tt = p.nty * p.ntx;
    i = k / tt;
    const int r2 = k - i * tt;
    jy = r2 / p.ntx;
    jx = r2 - jy * p.ntx;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = valid && k < p.K;
      v[j] = ok ? __ldg(w + (((size_t)i * g.Cout + o) * g.ks + p.ky0 + g.stride * jy) * g.ks + p.kx0 + g.stride * jx)
                : 0.f;
      ++k;
      if (++jx == p.ntx) {
        jx = 0;
        if (++jy == p.nty) { jy = 0; ++i; }
      }
    }
    k += 4;
    jx += 4;
    while (jx >= p.ntx) { jx -= p.ntx; ++jy; }
    while (jy >= p.nty) { jy -= p.nty; ++i; }
  }
};

__global__ void __launch_bounds__(TG_THREADS) convT_fwd_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               float* __restrict__ y, ConvTGeom g) {
  const PhaseGeom p = make_phase(g, blockIdx.z);
  if ((int)blockIdx.x * TG_BM >= p.M) return;  // phases may differ in size (uniform per block)
  ConvTA la{x, g, p};
  ConvTB lb{w, g, p};
  TileAcc acc;
  tile_gemm(la, lb, blockIdx.x * TG_BM, blockIdx.y * TG_BN, 0, p.K, acc);
  const int hw = p.Hm * p.Wm;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = acc.m(i);
    if (m >= p.M) continue;
    const int n = m / hw, r = m - n * hw;
    const int my = r / p.Wm, mx = r - my * p.Wm;
    const int oy = my * g.stride + p.ry, ox = mx * g.stride + p.rx;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = acc.n(j);
      if (o >= g.Cout) continue;
      const float bv = bias ? __ldg(bias + o) : 0.f;
      y[(((size_t)n * g.Cout + o) * g.Ho + oy) * g.Wo + ox] = acc.v[i][j] + bv;
    }
  }
}

// Small-Cout (<=4) transposed conv (deconv3: N -> 3 channels): one thread per output pixel of
// a phase, all Cout channels; the phase's weights sit in shared memory as float4 per tap.
__global__ void __launch_bounds__(256) convT_fwd_smallN_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               float* __restrict__ y, ConvTGeom g) {
  extern __shared__ float4 wsm[];  // [Cin][nty][ntx]
  const PhaseGeom p = make_phase(g, blockIdx.z);
  if ((int)blockIdx.x * 256 >= p.M) return;
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    const int tt = p.nty * p.ntx;
    const int i = k / tt, r2 = k - i * tt, jy = r2 / p.ntx, jx = r2 - jy * p.ntx;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int o = 0; o < g.Cout; ++o)
      t[o] = __ldg(w + (((size_t)i * g.Cout + o) * g.ks + p.ky0 + g.stride * jy) * g.ks + p.kx0 + g.stride * jx);
    wsm[k] = make_float4(t[0], t[1], t[2], t[3]);
  }
  __syncthreads();
  const int m = blockIdx.x * 256 + threadIdx.x;
  if (m >= p.M) return;
  const int hw = p.Hm * p.Wm;
  const int n = m / hw, r = m - n * hw;
  const int my = r / p.Wm, mx = r - my * p.Wm;
  const int y0 = my + p.dy0, x0 = mx + p.dx0;
  const float* base = x + (size_t)n * g.Cin * g.H * g.W;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const size_t plane = (size_t)g.H * g.W;
  for (int jy = 0; jy < p.nty; ++jy) {
    const int iy = y0 - jy;
    if (iy < 0 || iy >= g.H) continue;
    for (int jx = 0; jx < p.ntx; ++jx) {
      const int ix = x0 - jx;
      if (ix < 0 || ix >= g.W) continue;
      const float* px = base + (size_t)iy * g.W + ix;
      const float4* pw = wsm + jy * p.ntx + jx;
      const int tt = p.nty * p.ntx;
#pragma unroll 4
      for (int i = 0; i < g.Cin; ++i) {
        const float xv = __ldg(px + i * plane);
        const float4 wv = pw[i * tt];
        a0 = fmaf(xv, wv.x, a0);
        a1 = fmaf(xv, wv.y, a1);
        a2 = fmaf(xv, wv.z, a2);
        a3 = fmaf(xv, wv.w, a3);
      }
    }
  }
  const int oy = my * g.stride + p.ry, ox = mx * g.stride + p.rx;
  const float acc[4] = {a0, a1, a2, a3};
  for (int o = 0; o < g.Cout; ++o)
    y[(((size_t)n * g.Cout + o) * g.Ho + oy) * g.Wo + ox] = acc[o] + (bias ? __ldg(bias + o) : 0.f);
}

// ------------------------------------------------------------------------------- wgrad
// dW[cs][cb,ky,kx] += sum_{m=(n,ys,xs)} small[n,cs,ys,xs] * big[n,cb,s*ys-p+ky,s*xs-p+kx]
struct WgGeom {
  int B, Cs, hs, ws, Cb, hb, wb, ks, stride, pad, M, NK;  // M = B*hs*ws (reduction), NK = Cb*ks*ks
};

struct WgSmall {  // "A[row=cs][red=m]"
  const float* t;
  WgGeom g;
  int cs, m, m_end;
  bool valid;
  __device__ __forceinline__ void init(int cs_, int m_first) {
    cs = cs_;
    valid = cs < g.Cs;
    m = m_first;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
    const int hw = g.hs * g.ws;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mj = m + j;
      if (valid && mj < m_end) {
        const int n = mj / hw, r = mj - n * hw;
        v[j] = __ldg(t + ((size_t)n * g.Cs + cs) * hw + r);
      } else {
        v[j] = 0.f;
      }
    }
    m += 8;
  }
};

struct WgBig {  // "B[red=m][col=(cb,ky,kx)]"
  const float* t;
  WgGeom g;
  int m, m_end, n, ys, xs, ky, kx;
  const float* chan;
  bool valid;
  __device__ __forceinline__ void init(int col, int m_first) {
    valid = col < g.NK;
    const int kk = g.ks * g.ks;
    const int cb = valid ? col / kk : 0;
    const int r2 = valid ? col - cb * kk : 0;
    ky = r2 / g.ks - g.pad;
    kx = r2 - (r2 / g.ks) * g.ks - g.pad;
    chan = t + (size_t)cb * g.hb * g.wb;
    m = m_first;
    const int hw = g.hs * g.ws;
    n = m / hw;
    const int r = m - n * hw;
    ys = r / g.ws;
    xs = r - ys * g.ws;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int yb = ys * g.stride + ky, xb = xs * g.stride + kx;
      const bool ok = valid && m < m_end && yb >= 0 && yb < g.hb && xb >= 0 && xb < g.wb;
      v[j] = ok ? __ldg(chan + ((size_t)n * g.Cb * g.hb + yb) * g.wb + xb) : 0.f;
      ++m;
      if (++xs == g.ws) {
        xs = 0;
        if (++ys == g.hs) { ys = 0; ++n; }
      }
    }
    m += 4;
    xs += 4;
    while (xs >= g.ws) { xs -= g.ws; ++ys; }
    while (ys >= g.hs) { ys -= g.hs; ++n; }
  }
};

__global__ void __launch_bounds__(TG_THREADS) conv_wgrad_kernel(const float* __restrict__ small_t,
                                                                const float* __restrict__ big_t,
                                                                float* __restrict__ dW, WgGeom g, int chunk) {
  const int m_begin = blockIdx.z * chunk;
  const int m_end = min(g.M, m_begin + chunk);
  if (m_begin >= m_end) return;
  WgSmall la{small_t, g};
  la.m_end = m_end;
  WgBig lb{big_t, g};
  lb.m_end = m_end;
  TileAcc acc;
  tile_gemm(la, lb, blockIdx.x * TG_BM, blockIdx.y * TG_BN, m_begin, m_end, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cs = acc.m(i);
    if (cs >= g.Cs) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = acc.n(j);
      if (col >= g.NK) continue;
      atomicAdd(dW + (size_t)cs * g.NK + col, acc.v[i][j]);
    }
  }
}

// ------------------------------------------------------------------------- channel sum
// grid (C, splits): each block reduces a slice of the B*HW elements of one channel; with few channels
// (deconv3 bias: C = 3) a single block per channel would serialise hundreds of thousands of loads.
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ t, float* __restrict__ out,
                                                          int B, int C, int HW) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const size_t total = (size_t)B * HW;
  const size_t per = (total + gridDim.y - 1) / gridDim.y;
  const size_t i0 = blockIdx.y * per, i1 = min(total, i0 + per);
  float s = 0.f;
  for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const size_t n = i / HW, r = i - n * HW;
    s += __ldg(t + (n * C + c) * HW + r);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    if (gridDim.y == 1) out[c] = s;
    else atomicAdd(out + c, s);
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" {

int b2c_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W,
                   int Cout, int k, int stride, int pad, void* stream) {
  B2C_REQUIRE(x && w && y, "conv2d_fwd: null pointer");
  B2C_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0 && k > 0 && stride > 0 && pad >= 0,
              "conv2d_fwd: bad geometry");
  ConvGeom g;
  g.B = B; g.Cin = Cin; g.H = H; g.W = W; g.Cout = Cout; g.ks = k; g.stride = stride; g.pad = pad;
  g.Ho = (H + 2 * pad - k) / stride + 1;
  g.Wo = (W + 2 * pad - k) / stride + 1;
  B2C_REQUIRE(g.Ho > 0 && g.Wo > 0, "conv2d_fwd: empty output");
  g.M = B * g.Ho * g.Wo;
  g.K = Cin * k * k;
  dim3 grid((g.M + TG_BM - 1) / TG_BM, (Cout + TG_BN - 1) / TG_BN);
  conv2d_fwd_kernel<<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(x, w, bias, y, g);
  count_launch();
  return check_launch("conv2d_fwd");
}

int b2c_conv_transpose2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H,
                             int W, int Cout, int k, int stride, int pad, int out_pad, void* stream) {
  B2C_REQUIRE(x && w && y, "conv_transpose2d_fwd: null pointer");
  B2C_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0 && k > 0 && stride > 0 && pad >= 0 && out_pad >= 0,
              "conv_transpose2d_fwd: bad geometry");
  ConvTGeom g;
  g.B = B; g.Cin = Cin; g.H = H; g.W = W; g.Cout = Cout; g.ks = k; g.stride = stride; g.pad = pad;
  g.Ho = (H - 1) * stride - 2 * pad + k + out_pad;
  g.Wo = (W - 1) * stride - 2 * pad + k + out_pad;
  B2C_REQUIRE(g.Ho > 0 && g.Wo > 0, "conv_transpose2d_fwd: empty output");
  const int Hm = (g.Ho + stride - 1) / stride, Wm = (g.Wo + stride - 1) / stride;
  const int Mmax = B * Hm * Wm;
  const int phases = stride * stride;
  if (Cout <= 4) {
    const int maxt = (k + stride - 1) / stride;
    const size_t smem = (size_t)Cin * maxt * maxt * sizeof(float4);
    B2C_REQUIRE(smem <= 200 * 1024, "conv_transpose2d_fwd: Cin too large for the small-Cout kernel");
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(convT_fwd_smallN_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
      if (e != cudaSuccess) {
        set_error("conv_transpose2d_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return B2C_ECUDA;
      }
    }
    dim3 grid((Mmax + 255) / 256, 1, phases);
    convT_fwd_smallN_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, bias, y, g);
  } else {
    dim3 grid((Mmax + TG_BM - 1) / TG_BM, (Cout + TG_BN - 1) / TG_BN, phases);
    convT_fwd_kernel<<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(x, w, bias, y, g);
  }
  count_launch();
  return check_launch("conv_transpose2d_fwd");
}

int b2c_conv_wgrad(const float* small_t, const float* big_t, float* dW, int B, int Cs, int hs, int ws, int Cb,
                   int hb, int wb, int k, int stride, int pad, void* stream) {
  B2C_REQUIRE(small_t && big_t && dW, "conv_wgrad: null pointer");
  B2C_REQUIRE(B > 0 && Cs > 0 && hs > 0 && ws > 0 && Cb > 0 && hb > 0 && wb > 0 && k > 0 && stride > 0 && pad >= 0,
              "conv_wgrad: bad geometry");
  WgGeom g;
  g.B = B; g.Cs = Cs; g.hs = hs; g.ws = ws; g.Cb = Cb; g.hb = hb; g.wb = wb; g.ks = k; g.stride = stride;
  g.pad = pad;
  g.M = B * hs * ws;
  g.NK = Cb * k * k;
  const int tiles = ((Cs + TG_BM - 1) / TG_BM) * ((g.NK + TG_BN - 1) / TG_BN);
  int splits = (148 * 4 + tiles - 1) / tiles;
  const int max_splits = (g.M + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int chunk = (g.M + splits - 1) / splits;
  chunk = (chunk + 7) & ~7;
  splits = (g.M + chunk - 1) / chunk;
  dim3 grid((Cs + TG_BM - 1) / TG_BM, (g.NK + TG_BN - 1) / TG_BN, splits);
  conv_wgrad_kernel<<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(small_t, big_t, dW, g, chunk);
  count_launch();
  return check_launch("conv_wgrad");
}

int b2c_channel_sum(const float* t, float* out, int B, int C, int HW, void* stream) {
  B2C_REQUIRE(t && out && B > 0 && C > 0 && HW > 0, "channel_sum: bad arguments");
  const size_t total = (size_t)B * HW;
  int splits = (int)((total + 8191) / 8192);                 // >= 8192 elements per block
  const int max_splits = (148 * 4 + C - 1) / C;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 1) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("channel_sum: memset: %s", cudaGetErrorString(e)); return B2C_ECUDA; }
  }
  channel_sum_kernel<<<dim3(C, splits), 256, 0, (cudaStream_t)stream>>>(t, out, B, C, HW);
  count_launch();
  return check_launch("channel_sum");
}

}  // extern "C"
