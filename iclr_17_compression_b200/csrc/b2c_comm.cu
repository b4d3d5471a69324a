// Gradient all-reduce (mean) over NVLink / NVSwitch peer memory for the data-parallel training step (one
// exchange step, 7 MB of fp32 gradients per rank).  NCCL needs ~58 us for this size on 2 GPUs (protocol latency); here
// every rank owns 1/world of the flat buffer and the buffer is SYMMETRIC memory (same allocation mapped into every rank):
//   * multicast form (NVSwitch, NVLS): `multimem.ld_reduce.add` on the multicast address returns the sum over all
//     ranks computed IN THE SWITCH, `multimem.st` writes the scaled result back to all ranks' copies - one load and one
//     store per 16 bytes per rank, in place;
//   * peer form (no multicast): the owner reads its slice from every peer's buffer, sums in rank order
//     (deterministic), and writes the result into every peer's slice (two-shot, in place).
// Callers bracket the launch with a cross-rank barrier on the same stream (inputs complete before, results visible
// after); the slices are disjoint, so no other synchronisation is needed.
#include "b2c_common.cuh"

namespace b2c {

constexpr int AR_MAX_WORLD = 16;

struct PeerPtrs {
  float* p[AR_MAX_WORLD];
};

__global__ void __launch_bounds__(256) allreduce_multimem_kernel(float* __restrict__ mc, size_t begin4, size_t end4, float scale) {
  for (size_t i = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v;
    float* addr = mc + i * 4;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

__global__ void __launch_bounds__(256) allreduce_peer_kernel(const PeerPtrs peers, int world, size_t begin4, size_t end4,
                                                             float scale) {
  for (size_t i = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end4; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {                          // fixed rank order: every rank computes identical sums
      float4 v;                                                 // peers' data: bypass L1 (ld.global.cv semantics)
      asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(peers.p[r] + i * 4) : "memory");
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(peers.p[r] + i * 4) = s;
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_allreduce_symm(void* const* peer_ptrs, void* multicast_ptr, int rank, int world, size_t count,
                                  float scale, void* stream) {
  B2C_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "allreduce_symm: bad rank / world");
  B2C_REQUIRE(count > 0 && count % (4 * (size_t)world) == 0, "allreduce_symm: count must be a multiple of 4 * world");
  B2C_REQUIRE(multicast_ptr || peer_ptrs, "allreduce_symm: need a multicast pointer or the peer pointers");
  const size_t slice4 = count / 4 / world, begin4 = slice4 * rank, end4 = begin4 + slice4;
  size_t blocks = (slice4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  if (multicast_ptr) {
    B2C_REQUIRE(((uintptr_t)multicast_ptr & 15) == 0, "allreduce_symm: misaligned multicast pointer");
    allreduce_multimem_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)multicast_ptr, begin4, end4, scale);
  } else {
    PeerPtrs pp;
    for (int r = 0; r < world; ++r) {
      B2C_REQUIRE(peer_ptrs[r] && ((uintptr_t)peer_ptrs[r] & 15) == 0, "allreduce_symm: bad peer pointer %d", r);
      pp.p[r] = (float*)peer_ptrs[r];
    }
    allreduce_peer_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pp, world, begin4, end4, scale);
  }
  count_launch();
  return check_launch("allreduce_symm");
}
