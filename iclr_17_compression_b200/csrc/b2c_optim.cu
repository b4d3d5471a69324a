// Fused training-step tail: element-wise gradient clamp (train.py:106-111) + torch.optim.Adam update
// (train.py:112,233: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) for ALL parameter tensors in one
// launch.  Gradients, first and second moments live in flat fp32 buffers (concatenated in parameter order); the
// parameters stay the reference's separate tensors, addressed through a pointer table.  Step count and learning
// rate are read from device memory so the launch can be replayed inside a CUDA graph.
#include "b2c_common.cuh"

namespace b2c {

constexpr int ADAM_MAX_TENSORS = 32;

struct AdamTable {
  float* p[ADAM_MAX_TENSORS];
  unsigned start[ADAM_MAX_TENSORS + 1];   // flat offset of tensor i; start[n] = total
  int n;
};

__global__ void __launch_bounds__(256) adam_clamp_kernel(const AdamTable T, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         const int* __restrict__ step_dev,
                                                         const float* __restrict__ lr_dev, float beta1, float beta2,
                                                         float omb1, float omb2, float eps, float clip) {
  __shared__ float s_step_size, s_bc2_sqrt;
  __shared__ unsigned s_start[ADAM_MAX_TENSORS + 1];
  if (threadIdx.x <= T.n) s_start[threadIdx.x] = T.start[threadIdx.x];
  if (threadIdx.x == 0) {
    const double step = (double)step_dev[0];
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    s_step_size = (float)((double)lr_dev[0] / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const unsigned total = s_start[T.n];
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  auto update = [&](float gi, float& mi, float& vi, float& pi) {
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);                 // p.grad.data.clamp_(-5, 5)
    // omb = (float)(1.0 - (double)beta): torch forms 1 - beta in double before the scalar becomes fp32
    // (1.f - 0.999f differs from 0.001f by 4.7e-5 relative)
    mi = fmaf(omb1, gi - mi, mi);                                       // exp_avg.lerp_(grad, 1 - beta1)
    vi = fmaf(omb2, gi * gi, vi * beta2);                               // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);                                 // param.addcdiv_(exp_avg, denom, -step_size)
  };
  // four elements per thread (128-bit accesses) wherever a quad lies inside one tensor and both the flat offset and
  // the parameter pointer are 16-byte aligned; scalar otherwise (the 3-element deconv3 bias shifts every later tensor)
  const bool flat_aligned = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                              reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  const unsigned nquads = (total + 3) / 4;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += gridDim.x * blockDim.x) {
    const unsigned i = q * 4;
    int lo = 0, hi = T.n;                       // tensor t with start[t] <= i < start[t+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_start[mid] <= i) lo = mid; else hi = mid;
    }
    float* pp = T.p[lo] + (i - s_start[lo]);
    if (flat_aligned && i + 4 <= s_start[lo + 1] && (reinterpret_cast<uintptr_t>(pp) & 15u) == 0) {
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
      float4 p4 = *reinterpret_cast<float4*>(pp);
      update(g4.x, m4.x, v4.x, p4.x);
      update(g4.y, m4.y, v4.y, p4.y);
      update(g4.z, m4.z, v4.z, p4.z);
      update(g4.w, m4.w, v4.w, p4.w);
      *reinterpret_cast<float4*>(m + i) = m4;
      *reinterpret_cast<float4*>(v + i) = v4;
      *reinterpret_cast<float4*>(pp) = p4;
    } else {
      for (unsigned e = i; e < i + 4 && e < total; ++e) {
        while (e >= s_start[lo + 1]) ++lo;
        float mi = m[e], vi = v[e];
        float* pe = T.p[lo] + (e - s_start[lo]);
        float pi = *pe;
        update(g[e], mi, vi, pi);
        m[e] = mi;
        v[e] = vi;
        *pe = pi;
      }
    }
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_adam_clamp_step(float* const* params, const size_t* numel, int n_tensors, const float* grads_flat,
                                   float* exp_avg, float* exp_avg_sq, const int* step_dev, const float* lr_dev,
                                   double beta1, double beta2, double eps, float grad_clip, void* stream) {
  B2C_REQUIRE(params && numel && grads_flat && exp_avg && exp_avg_sq && step_dev && lr_dev, "adam_clamp_step: null pointer");
  B2C_REQUIRE(n_tensors > 0 && n_tensors <= ADAM_MAX_TENSORS, "adam_clamp_step: 1..%d tensors", ADAM_MAX_TENSORS);
  AdamTable T;
  size_t off = 0;
  for (int i = 0; i < n_tensors; ++i) {
    B2C_REQUIRE(params[i] && numel[i] > 0, "adam_clamp_step: bad tensor %d", i);
    T.p[i] = params[i];
    T.start[i] = (unsigned)off;
    off += numel[i];
    B2C_REQUIRE(off < (1ull << 32), "adam_clamp_step: too many elements");
  }
  T.start[n_tensors] = (unsigned)off;
  T.n = n_tensors;
  unsigned blocks = (unsigned)((off / 4 + 255) / 256);     // four elements per thread
  if (blocks == 0) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_clamp_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(T, grads_flat, exp_avg, exp_avg_sq, step_dev, lr_dev,
                                                             (float)beta1, (float)beta2, (float)(1.0 - beta1),
                                                             (float)(1.0 - beta2), (float)eps, grad_clip);
  count_launch();
  return check_launch("adam_clamp_step");
}
