// Packed-weight layout of the x-folded last synthesis layer (b2c_deconv3x.cu), shared with the multi-job packing
// kernel of b2c_conv_tc.cu.
#pragma once

namespace b2c {
namespace tcp3x {

constexpr int X_NX = 112;                // GEMM N: 4 sub-pixel rows x (12 dx=0, 12 dx=+1, 3 dx=-1, 1 zero)
constexpr int X_PART_COLS = X_NX / 4;    // 28 accumulator columns per sub-pixel row ry

// weights IOHW [C][3][9][9] -> element idx of [tap t = dy + 1][n2 112][c C]; column n2 = ry*28 + j of the GEMM:
//   j in [0,12):  dx = 0,  j = rx*3 + o        j in [12,24): dx = +1, same decomposition of j - 12
//   j in [24,27): dx = -1, rx = 0, o = j - 24  j = 27: zero
// with ky = ry + 4 - 4 dy, kx = rx + 4 - 4 dx (zero where ky > 8)
__device__ __forceinline__ float deconv3x_weight(const float* __restrict__ w, int idx, int C = 128) {
  const int c = idx % C, n2 = (idx / C) % X_NX, t = idx / (C * X_NX);
  const int ry = n2 / X_PART_COLS, j = n2 - ry * X_PART_COLS;
  float v = 0.f;
  if (j < 27) {
    int dx, rx, o;
    if (j < 12) { dx = 0; rx = j / 3; o = j % 3; }
    else if (j < 24) { dx = 1; rx = (j - 12) / 3; o = j % 3; }
    else { dx = -1; rx = 0; o = j - 24; }
    const int dy = t - 1;
    const int ky = ry + 4 - 4 * dy, kx = rx + 4 - 4 * dx;
    if (ky >= 0 && ky < 9 && kx >= 0 && kx < 9) v = w[((c * 3 + o) * 9 + ky) * 9 + kx];
  }
  return v;
}

}  // namespace tcp3x
}  // namespace b2c
