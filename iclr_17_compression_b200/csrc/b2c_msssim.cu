// SSIM / MS-SSIM on the GPU: the reference's in-training validation copies every reconstruction to
// the CPU and runs models/ms_ssim_torch.py there (train.py:178).  One fused kernel per scale computes the five
// Gaussian-filtered moments (mu_x, mu_y, E[x^2], E[y^2], E[xy]; separable 1-D window, valid padding,
// ms_ssim_torch.py:21-35,59-68), the ssim / cs maps (:70-71) and their per-image sums; avg_pool2d(kernel 2, padding
// = size % 2, zero padding counted) produces the next scale (:164-166); a last kernel forms
// prod_l cs_l^w_l * ssim_L^w_L (:169-170).
#include "b2c_common.cuh"

namespace b2c {

constexpr int SS_T = 16;          // output tile edge
constexpr int SS_MAXW = 33;       // largest supported window

__global__ void __launch_bounds__(256) ssim_level_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                         const float* __restrict__ win, int ws, float C1, float C2,
                                                         float* __restrict__ partial, int H, int W, int tiles_x,
                                                         int tiles_y) {
  extern __shared__ float ss_smem[];
  const int E = SS_T + ws - 1;                       // input tile edge
  float* sx = ss_smem;                               // [E][E]
  float* sy = sx + E * E;                            // [E][E]
  float* hz = sy + E * E;                            // [5][E][SS_T] horizontally filtered moments
  float* sw = hz + 5 * E * SS_T;                     // [ws]
  __shared__ float red[32];
  const int plane = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
  const int Ho = H - ws + 1, Wo = W - ws + 1;
  const int oy0 = ty * SS_T, ox0 = tx * SS_T;
  const float* xp = X + (size_t)plane * H * W;
  const float* yp = Y + (size_t)plane * H * W;
  for (int i = threadIdx.x; i < ws; i += blockDim.x) sw[i] = win[i];
  for (int i = threadIdx.x; i < E * E; i += blockDim.x) {
    const int r = i / E, c = i - r * E;
    const int gy = oy0 + r, gx = ox0 + c;
    const bool in = gy < H && gx < W;
    sx[i] = in ? __ldg(xp + (size_t)gy * W + gx) : 0.f;
    sy[i] = in ? __ldg(yp + (size_t)gy * W + gx) : 0.f;
  }
  __syncthreads();
  // horizontal pass (first F.conv2d of gaussian_filter: 1 x ws window along W)
  for (int i = threadIdx.x; i < E * SS_T; i += blockDim.x) {
    const int r = i / SS_T, c = i - r * SS_T;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
    for (int k = 0; k < ws; ++k) {
      const float w = sw[k], xv = sx[r * E + c + k], yv = sy[r * E + c + k];
      a = fmaf(w, xv, a); b = fmaf(w, yv, b);
      aa = fmaf(w, xv * xv, aa); bb = fmaf(w, yv * yv, bb); ab = fmaf(w, xv * yv, ab);
    }
    hz[0 * E * SS_T + i] = a; hz[1 * E * SS_T + i] = b; hz[2 * E * SS_T + i] = aa;
    hz[3 * E * SS_T + i] = bb; hz[4 * E * SS_T + i] = ab;
  }
  __syncthreads();
  // vertical pass + ssim / cs maps for this thread's output pixel
  const int r = threadIdx.x / SS_T, c = threadIdx.x - r * SS_T;
  float s_ssim = 0.f, s_cs = 0.f;
  if (oy0 + r < Ho && ox0 + c < Wo) {
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ws; ++k) {
      const float w = sw[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(w, hz[q * E * SS_T + (r + k) * SS_T + c], m[q]);
    }
    const float mu1 = m[0], mu2 = m[1];
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s1 = m[2] - mu1_sq, s2 = m[3] - mu2_sq, s12 = m[4] - mu12;
    const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
    s_cs = cs;
    s_ssim = ((2.f * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs;
  }
  s_ssim = block_sum(s_ssim, red);
  __syncthreads();
  s_cs = block_sum(s_cs, red);
  if (threadIdx.x == 0) {
    float* dst = partial + (((size_t)plane * tiles_y + ty) * tiles_x + tx) * 2;
    dst[0] = s_ssim;
    dst[1] = s_cs;
  }
}

// F.avg_pool2d(x, kernel_size=2, padding=(ph, pw)): stride 2, zero padding counted in the divisor (always 4)
__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho, int Wo,
                                int ph, int pw, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const size_t plane = i / ((size_t)Wo * Ho);
    const float* p = in + plane * H * W;
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy - ph + dy, ix = 2 * ox - pw + dx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) s += __ldg(p + (size_t)iy * W + ix);
      }
    out[i] = s * 0.25f;
  }
}

// ms_ssim_torch.py:169-170 evaluates  prod_l ( mcs[l]^w[l] * ssim_last^w[L-1] )  over the first L-1 levels: the
// broadcast puts the ssim factor INSIDE the product, so it enters L-1 times.  Reproduced as written (parity with the
// reference's numbers, not with the textbook formula):  msssim[b] = prod_{l<L-1} mcs[l][b]^w[l] * ssim_last[b]^((L-1) w[L-1])
__global__ void msssim_combine_kernel(const float* __restrict__ sums, const float* __restrict__ counts,
                                      const float* __restrict__ weights, float* __restrict__ out, int levels, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float v = 1.f;
  for (int l = 0; l < levels; ++l) {
    const float ssim = sums[((size_t)l * B + b) * 2] / counts[l], cs = sums[((size_t)l * B + b) * 2 + 1] / counts[l];
    v *= (l == levels - 1) ? powf(ssim, weights[l] * (float)(levels - 1)) : powf(cs, weights[l]);
  }
  out[b] = v;
}

}  // namespace b2c

using namespace b2c;

extern "C" {

size_t b2c_ssim_workspace_bytes(int B, int C, int H, int W, int win_size) {
  if (B <= 0 || C <= 0 || win_size < 1 || H < win_size || W < win_size) return 0;
  const int tx = (W - win_size + 1 + SS_T - 1) / SS_T, ty = (H - win_size + 1 + SS_T - 1) / SS_T;
  return (size_t)B * C * tx * ty * 2 * sizeof(float);
}

int b2c_ssim_level(const float* X, const float* Y, const float* win, int win_size, float C1, float C2, float* sums, int B,
                   int C, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(X && Y && win && sums && workspace && B > 0 && C > 0, "ssim_level: bad arguments");
  B2C_REQUIRE(win_size >= 1 && win_size <= SS_MAXW && (win_size & 1), "ssim_level: window size must be odd, <= %d", SS_MAXW);
  B2C_REQUIRE(H >= win_size && W >= win_size, "ssim_level: image %dx%d smaller than the %d-tap window", H, W, win_size);
  if (workspace_bytes < b2c_ssim_workspace_bytes(B, C, H, W, win_size)) {
    set_error("ssim_level: workspace too small");
    return B2C_EWORKSPACE;
  }
  const int tx = (W - win_size + 1 + SS_T - 1) / SS_T, ty = (H - win_size + 1 + SS_T - 1) / SS_T;
  const int E = SS_T + win_size - 1;
  const size_t smem = (size_t)(2 * E * E + 5 * E * SS_T + win_size) * sizeof(float);
  B2C_REQUIRE(smem <= 48 * 1024, "ssim_level: window too large for shared memory");
  B2C_REQUIRE((size_t)B * C <= 65535 && ty <= 65535, "ssim_level: too many planes / tiles");
  dim3 grid(tx, ty, B * C);
  ssim_level_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(X, Y, win, win_size, C1, C2, (float*)workspace, H, W, tx, ty);
  count_launch();
  int rc = check_launch("ssim_level");
  if (rc) return rc;
  return b2c_partial_reduce((const float*)workspace, sums, B, C * tx * ty, 2, stream);   // fixed-order per-image sums
}

int b2c_avgpool2(const float* in, float* out, int planes, int H, int W, int pad_h, int pad_w, void* stream) {
  B2C_REQUIRE(in && out && planes > 0 && H > 0 && W > 0 && pad_h >= 0 && pad_h <= 1 && pad_w >= 0 && pad_w <= 1,
              "avgpool2: bad arguments");
  const int Ho = (H + 2 * pad_h - 2) / 2 + 1, Wo = (W + 2 * pad_w - 2) / 2 + 1;
  const size_t total = (size_t)planes * Ho * Wo;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  avgpool2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, H, W, Ho, Wo, pad_h, pad_w, total);
  count_launch();
  return check_launch("avgpool2");
}

int b2c_msssim_combine(const float* sums, const float* counts, const float* weights, float* out, int levels, int B,
                       void* stream) {
  B2C_REQUIRE(sums && counts && weights && out && levels > 0 && B > 0, "msssim_combine: bad arguments");
  msssim_combine_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, counts, weights, out, levels, B);
  count_launch();
  return check_launch("msssim_combine");
}

}  // extern "C"
