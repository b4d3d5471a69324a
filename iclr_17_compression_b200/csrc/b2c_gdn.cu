// GDN / IGDN, exact fp32 (NCHW).  Reference: models/GDN.py:64-94 (forward), :10-24 (LowerBound).
//   beta  = max(beta_p,  beta_bound )^2 - pedestal          (GDN.py:74-75)
//   gamma = max(gamma_p, gamma_bound)^2 - pedestal          (GDN.py:78-79)
//   norm_i = beta_i + sum_j gamma[i][j] x_j^2 ; y = x / sqrt(norm)  or  x * sqrt(norm)   (GDN.py:83-90)
// The channel-mixing product is a [pixels x C] x [C x C] GEMM on the shared fp32 tile core.
#include "b2c_common.cuh"

namespace b2c {

constexpr float kPedestal = 1.4551915228366852e-11f;   // 2^-36            (GDN.py:47)
constexpr float kBetaBound = (float)0.0010000072759311445;  // sqrt(1e-6 + 2^-36) rounded to fp32 as torch does (GDN.py:48)
constexpr float kGammaBound = 3.814697265625e-06f;     // 2^-18            (GDN.py:49)

__device__ __forceinline__ float gamma_eff(float gp) {
  const float b = fmaxf(gp, kGammaBound);
  return b * b - kPedestal;
}
__device__ __forceinline__ float beta_eff(float bp) {
  const float b = fmaxf(bp, kBetaBound);
  return b * b - kPedestal;
}

struct PixGeom {
  int B, C, HW, M;
};

// A[m][k=channel] = f(t[n, k, pix]),  m = (n, pix)
template <bool SQUARE>
struct PixA {
  const float* t;
  PixGeom g;
  const float* p;
  int k;
  bool valid;
  __device__ __forceinline__ void init(int m, int k_first) {
    valid = m < g.M;
    const int n = valid ? m / g.HW : 0;
    const int r = valid ? m - n * g.HW : 0;
    p = t + (size_t)n * g.C * g.HW + r;
    k = k_first;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = (valid && k + j < g.C) ? __ldg(p + (size_t)(k + j) * g.HW) : 0.f;
      v[j] = SQUARE ? a * a : a;
    }
    k += 8;
  }
};

// B[k=j][col=i] = gamma_eff[i][j]   (TRANSPOSED=false: forward product, reduce over gamma's 2nd index)
// B[k=i][col=j] = gamma_eff[i][j]   (TRANSPOSED=true : backward product, reduce over gamma's 1st index)
template <bool TRANSPOSED>
struct GammaB {
  const float* gamma_p;
  int C;
  int col, k;
  bool valid;
  __device__ __forceinline__ void init(int col_, int k_first) {
    col = col_;
    valid = col < C;
    k = k_first;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k + j;
      const bool ok = valid && kk < C;
      const float gp = ok ? __ldg(gamma_p + (TRANSPOSED ? (size_t)kk * C + col : (size_t)col * C + kk)) : 0.f;
      v[j] = ok ? gamma_eff(gp) : 0.f;
    }
    k += 8;
  }
};

// MODE 0: forward, write y.   MODE 1: backward stage 1, write dn (dL/dnorm) and the direct term of gx.
template <int MODE>
__global__ void __launch_bounds__(TG_THREADS) gdn_norm_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ beta_p,
                                                              const float* __restrict__ gamma_p,
                                                              const float* __restrict__ gy, float* __restrict__ out0,
                                                              float* __restrict__ out1, PixGeom g, int inverse) {
  PixA<true> la{x, g};
  GammaB<false> lb{gamma_p, g.C};
  TileAcc acc;
  tile_gemm(la, lb, blockIdx.x * TG_BM, blockIdx.y * TG_BN, 0, g.C, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = acc.m(i);
    if (m >= g.M) continue;
    const int n = m / g.HW, r = m - n * g.HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = acc.n(j);
      if (c >= g.C) continue;
      const size_t idx = ((size_t)n * g.C + c) * g.HW + r;
      const float xv = __ldg(x + idx);
      const float norm = acc.v[i][j] + beta_eff(__ldg(beta_p + c));
      const float s = sqrtf(norm);
      if (MODE == 0) {
        out0[idx] = inverse ? xv * s : xv / s;
      } else {
        const float gv = __ldg(gy + idx);
        if (inverse) {
          out0[idx] = 0.5f * gv * xv / s;  // dL/dnorm
          out1[idx] = gv * s;              // direct term of dL/dx
        } else {
          out0[idx] = -0.5f * gv * xv / (norm * s);
          out1[idx] = gv / s;
        }
      }
    }
  }
}

// backward stage 2: gx[m][j] += 2 x[m][j] * sum_i dn[m][i] gamma_eff[i][j]
__global__ void __launch_bounds__(TG_THREADS) gdn_bwd_dx_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ gamma_p,
                                                                const float* __restrict__ dn, float* __restrict__ gx,
                                                                PixGeom g) {
  PixA<false> la{dn, g};
  GammaB<true> lb{gamma_p, g.C};
  TileAcc acc;
  tile_gemm(la, lb, blockIdx.x * TG_BM, blockIdx.y * TG_BN, 0, g.C, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = acc.m(i);
    if (m >= g.M) continue;
    const int n = m / g.HW, r = m - n * g.HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = acc.n(j);
      if (c >= g.C) continue;
      const size_t idx = ((size_t)n * g.C + c) * g.HW + r;
      gx[idx] = fmaf(2.f * __ldg(x + idx), acc.v[i][j], gx[idx]);
    }
  }
}

// backward stage 3: dgamma_eff[i][j] += sum_m dn[m][i] * x[m][j]^2  (split over m, atomics)
template <bool SQUARE>
struct ChanRow {  // operand[row=channel][red=m] = f(t[n, channel, pix])
  const float* t;
  PixGeom g;
  int ch, m, m_end;
  bool valid;
  __device__ __forceinline__ void init(int ch_, int m_first) {
    ch = ch_;
    valid = ch < g.C;
    m = m_first;
  }
  __device__ __forceinline__ void fetch4(float v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mj = m + j;
      float a = 0.f;
      if (valid && mj < m_end) {
        const int n = mj / g.HW, r = mj - n * g.HW;
        a = __ldg(t + ((size_t)n * g.C + ch) * g.HW + r);
      }
      v[j] = SQUARE ? a * a : a;
    }
    m += 8;
  }
};

__global__ void __launch_bounds__(TG_THREADS) gdn_bwd_dgamma_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ dn,
                                                                    float* __restrict__ dgamma_eff, PixGeom g,
                                                                    int chunk) {
  const int m_begin = blockIdx.z * chunk;
  const int m_end = min(g.M, m_begin + chunk);
  if (m_begin >= m_end) return;
  ChanRow<false> la{dn, g};
  la.m_end = m_end;
  ChanRow<true> lb{x, g};
  lb.m_end = m_end;
  TileAcc acc;
  tile_gemm(la, lb, blockIdx.x * TG_BM, blockIdx.y * TG_BN, m_begin, m_end, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ci = acc.m(i);
    if (ci >= g.C) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cj = acc.n(j);
      if (cj >= g.C) continue;
      atomicAdd(dgamma_eff + (size_t)ci * g.C + cj, acc.v[i][j]);
    }
  }
}

__global__ void __launch_bounds__(256) channel_sum2_kernel(const float* __restrict__ t, float* __restrict__ out,
                                                           int B, int C, int HW) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int n = 0; n < B; ++n) {
    const float* p = t + ((size_t)n * C + c) * HW;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) s += __ldg(p + i);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[c] = s;
}

// backward stage 4: chain through  eff = max(p, bound)^2 - pedestal  with LowerBound's gradient rule
// (GDN.py:18-24): pass iff p >= bound or the incoming gradient is negative.
__global__ void gdn_reparam_bwd_kernel(const float* __restrict__ beta_p, const float* __restrict__ gamma_p,
                                       const float* __restrict__ dbeta_eff, const float* __restrict__ dgamma_eff,
                                       float* __restrict__ dbeta_p, float* __restrict__ dgamma_p, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < C * C) {
    const float p = gamma_p[idx];
    const float b = fmaxf(p, kGammaBound);
    const float gb = dgamma_eff[idx] * 2.f * b;  // dL/d(bounded)
    dgamma_p[idx] = (p >= kGammaBound || gb < 0.f) ? gb : 0.f;
  }
  if (idx < C) {
    const float p = beta_p[idx];
    const float b = fmaxf(p, kBetaBound);
    const float gb = dbeta_eff[idx] * 2.f * b;
    dbeta_p[idx] = (p >= kBetaBound || gb < 0.f) ? gb : 0.f;
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" {

int b2c_gdn_fwd(const float* x, const float* beta_p, const float* gamma_p, float* y, int B, int C, int HW,
                int inverse, void* stream) {
  B2C_REQUIRE(x && beta_p && gamma_p && y && B > 0 && C > 0 && HW > 0, "gdn_fwd: bad arguments");
  PixGeom g{B, C, HW, B * HW};
  dim3 grid((g.M + TG_BM - 1) / TG_BM, (C + TG_BN - 1) / TG_BN);
  gdn_norm_kernel<0><<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(x, beta_p, gamma_p, nullptr, y, nullptr, g,
                                                                   inverse);
  count_launch();
  return check_launch("gdn_fwd");
}

int b2c_gdn_reparam_bwd(const float* beta_p, const float* gamma_p, const float* dbeta_eff, const float* dgamma_eff,
                        float* dbeta_p, float* dgamma_p, int C, void* stream) {
  B2C_REQUIRE(beta_p && gamma_p && dbeta_eff && dgamma_eff && dbeta_p && dgamma_p && C > 0, "gdn_reparam_bwd: bad arguments");
  gdn_reparam_bwd_kernel<<<(C * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(beta_p, gamma_p, dbeta_eff, dgamma_eff,
                                                                                 dbeta_p, dgamma_p, C);
  count_launch();
  return check_launch("gdn_reparam_bwd");
}

size_t b2c_gdn_bwd_workspace_bytes(int B, int C, int HW) {
  return ((size_t)B * C * HW + (size_t)C * C + (size_t)C) * sizeof(float);
}

int b2c_gdn_bwd(const float* x, const float* beta_p, const float* gamma_p, const float* gy, float* gx,
                float* dbeta_p, float* dgamma_p, int B, int C, int HW, int inverse, void* workspace,
                size_t workspace_bytes, void* stream_) {
  B2C_REQUIRE(x && beta_p && gamma_p && gy && gx && dbeta_p && dgamma_p && workspace && B > 0 && C > 0 && HW > 0,
              "gdn_bwd: bad arguments");
  if (workspace_bytes < b2c_gdn_bwd_workspace_bytes(B, C, HW)) {
    set_error("gdn_bwd: workspace too small (%zu < %zu)", workspace_bytes, b2c_gdn_bwd_workspace_bytes(B, C, HW));
    return B2C_EWORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  PixGeom g{B, C, HW, B * HW};
  float* dn = (float*)workspace;
  float* dgamma_eff = dn + (size_t)B * C * HW;
  float* dbeta_eff = dgamma_eff + (size_t)C * C;
  cudaError_t e = cudaMemsetAsync(dgamma_eff, 0, ((size_t)C * C + C) * sizeof(float), stream);
  if (e != cudaSuccess) {
    set_error("gdn_bwd: memset: %s", cudaGetErrorString(e));
    return B2C_ECUDA;
  }
  dim3 grid((g.M + TG_BM - 1) / TG_BM, (C + TG_BN - 1) / TG_BN);
  gdn_norm_kernel<1><<<grid, TG_THREADS, 0, stream>>>(x, beta_p, gamma_p, gy, dn, gx, g, inverse);
  gdn_bwd_dx_kernel<<<grid, TG_THREADS, 0, stream>>>(x, gamma_p, dn, gx, g);
  const int tiles = ((C + TG_BM - 1) / TG_BM) * ((C + TG_BN - 1) / TG_BN);
  int splits = (148 * 2 + tiles - 1) / tiles;
  const int max_splits = (g.M + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int chunk = (g.M + splits - 1) / splits;
  chunk = (chunk + 7) & ~7;
  splits = (g.M + chunk - 1) / chunk;
  dim3 grid3((C + TG_BM - 1) / TG_BM, (C + TG_BN - 1) / TG_BN, splits);
  gdn_bwd_dgamma_kernel<<<grid3, TG_THREADS, 0, stream>>>(x, dn, dgamma_eff, g, chunk);
  channel_sum2_kernel<<<C, 256, 0, stream>>>(dn, dbeta_eff, B, C, HW);
  gdn_reparam_bwd_kernel<<<(C * C + 255) / 256, 256, 0, stream>>>(beta_p, gamma_p, dbeta_eff, dgamma_eff, dbeta_p,
                                                                 dgamma_p, C);
  count_launch(5);
  return check_launch("gdn_bwd");
}

}  // extern "C"
