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

__device__ __forceinline__ float4 mm_ld_reduce(float* addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* addr, float4 v, float scale) {
  v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// A switch round trip is microseconds: four independent reductions per thread are in flight before the first store, so
// a capped grid (a bucket reduced beside compute kernels) is not bound by one outstanding load per thread.
__global__ void __launch_bounds__(256) allreduce_multimem_kernel(float* __restrict__ mc, size_t begin4, size_t end4, float scale) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < end4; i += 4 * stride) {
    float* a0 = mc + i * 4;
    float* a1 = a0 + stride * 4;
    float* a2 = a1 + stride * 4;
    float* a3 = a2 + stride * 4;
    const float4 v0 = mm_ld_reduce(a0), v1 = mm_ld_reduce(a1), v2 = mm_ld_reduce(a2), v3 = mm_ld_reduce(a3);
    mm_st(a0, v0, scale); mm_st(a1, v1, scale); mm_st(a2, v2, scale); mm_st(a3, v3, scale);
  }
  for (; i < end4; i += stride) mm_st(mc + i * 4, mm_ld_reduce(mc + i * 4), scale);
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

// ---------------------------------------------------------------------------------------------------------------
// Fused form: cross-rank barrier -> reduce -> cross-rank barrier inside ONE kernel (the torch symmetric-memory barriers
// around the plain kernel are separate launches of ~8 us each - more than the 7 MB reduction itself).  Every rank keeps,
// in the tail of its symmetric buffer, one flag block per channel:
//     ready[w]  epoch announced by rank w: "my gradients are complete" (written by rank w's CTA 0 through peer memory)
//     done[w]   epoch announced by rank w: "my slice of the result is in everybody's buffer"
//     ticket    CTAs of this rank that have finished their stores (local)
//     epoch     number of completed launches on this channel (local; the launch number is NOT a kernel argument, so a
//               CUDA graph can replay the launch)
// All CTAs wait on the LOCAL ready flags (no co-residency requirement: a CTA scheduled late just starts late; CTA 0,
// which announces, is scheduled first), reduce their part of this rank's slice, and the last CTA to finish announces
// `done` to every peer, waits for every peer's `done` and resets the block - the kernel completes only then, so the next
// kernel in the stream sees the full result.
constexpr int AR_FLAG_WORDS = 2 * AR_MAX_WORLD + 2;     // ready[16] done[16] ticket epoch   (per channel)

struct FusedArgs {
  float* peers[AR_MAX_WORLD];   // base of every rank's symmetric buffer
  float* mc;                    // multicast base or nullptr
  size_t flag_off;              // first flag word of this channel, in elements from the buffer base
  size_t begin4, end4;          // this rank's float4 range
  float scale;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin: a rank that died must surface as a trap, not as a hung GPU
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    if (clock64() - t0 > 20000000000LL) {   // ~10 s
      printf("b2c: all-reduce flag wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(256) allreduce_fused_kernel(const FusedArgs a) {
  __shared__ unsigned s_epoch, s_last;
  unsigned* my = reinterpret_cast<unsigned*>(a.peers[a.rank] + a.flag_off);
  unsigned* ready = my;
  unsigned* done = my + AR_MAX_WORLD;
  unsigned* ticket = my + 2 * AR_MAX_WORLD;
  unsigned* epoch_p = ticket + 1;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(epoch_p) + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  // ---- barrier 1: every rank's input is complete
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(a.peers[threadIdx.x] + a.flag_off) + a.rank, epoch);
  }
  if (threadIdx.x < a.world) wait_flag(ready + threadIdx.x, epoch);
  __syncthreads();
  // ---- reduce this rank's slice, result written to every rank
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = a.begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a.mc) {
    for (; i + 7 * stride < a.end4; i += 8 * stride) {       // eight in-switch reductions in flight per thread
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = mm_ld_reduce(a.mc + (i + k * stride) * 4);
#pragma unroll
      for (int k = 0; k < 8; ++k) mm_st(a.mc + (i + k * stride) * 4, v[k], a.scale);
    }
    for (; i + 3 * stride < a.end4; i += 4 * stride) {
      float* a0 = a.mc + i * 4;
      float* a1 = a0 + stride * 4;
      float* a2 = a1 + stride * 4;
      float* a3 = a2 + stride * 4;
      const float4 v0 = mm_ld_reduce(a0), v1 = mm_ld_reduce(a1), v2 = mm_ld_reduce(a2), v3 = mm_ld_reduce(a3);
      mm_st(a0, v0, a.scale); mm_st(a1, v1, a.scale); mm_st(a2, v2, a.scale); mm_st(a3, v3, a.scale);
    }
    for (; i < a.end4; i += stride) mm_st(a.mc + i * 4, mm_ld_reduce(a.mc + i * 4), a.scale);
  } else {
    auto ldv = [](const float* p) {
      float4 v;
      asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
      return v;
    };
    if (a.world == 2) {      // two ranks: four positions x two ranks = eight loads in flight per thread
      for (; i + 3 * stride < a.end4; i += 4 * stride) {
        float4 v[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[k][0] = ldv(a.peers[0] + (i + k * stride) * 4);
          v[k][1] = ldv(a.peers[1] + (i + k * stride) * 4);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4 s4;
          s4.x = (v[k][0].x + v[k][1].x) * a.scale; s4.y = (v[k][0].y + v[k][1].y) * a.scale;
          s4.z = (v[k][0].z + v[k][1].z) * a.scale; s4.w = (v[k][0].w + v[k][1].w) * a.scale;
          *reinterpret_cast<float4*>(a.peers[0] + (i + k * stride) * 4) = s4;
          *reinterpret_cast<float4*>(a.peers[1] + (i + k * stride) * 4) = s4;
        }
      }
    }
    for (; i < a.end4; i += stride) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < a.world; ++r) {                       // fixed rank order: identical sums on every rank
        float4 v;
        asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a.peers[r] + i * 4) : "memory");
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
      for (int r = 0; r < a.world; ++r) *reinterpret_cast<float4*>(a.peers[r] + i * 4) = s;
    }
  }
  // ---- barrier 2: every rank's slice has landed everywhere.  One system-scope fence per CTA, by the thread that takes
  // the ticket, after the CTA barrier (fences are cumulative: it covers the stores of the whole CTA); 256 fence.sys per
  // CTA, each waiting for the multimem stores to be acknowledged across NVLink, were a visible part of the kernel.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned*>(a.peers[threadIdx.x] + a.flag_off) + AR_MAX_WORLD + a.rank, epoch);
      wait_flag(done + threadIdx.x, epoch);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *ticket = 0u;
      *reinterpret_cast<volatile unsigned*>(epoch_p) = epoch;
      __threadfence();
    }
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_allreduce_symm_flag_words(void) { return AR_FLAG_WORDS; }

extern "C" int b2c_allreduce_symm_fused(void* const* peer_ptrs, void* multicast_ptr, int rank, int world, size_t begin,
                                        size_t count, size_t flag_offset, float scale, int max_blocks, void* stream) {
  B2C_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "allreduce_symm_fused: bad rank / world");
  B2C_REQUIRE(count > 0 && count % 4 == 0 && begin % 4 == 0, "allreduce_symm_fused: begin and count must be multiples of 4");
  B2C_REQUIRE(peer_ptrs && flag_offset >= begin + count, "allreduce_symm_fused: the flag block lies behind the data");
  FusedArgs a;
  for (int r = 0; r < AR_MAX_WORLD; ++r) a.peers[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    B2C_REQUIRE(peer_ptrs[r] && ((uintptr_t)peer_ptrs[r] & 15) == 0, "allreduce_symm_fused: bad peer pointer %d", r);
    a.peers[r] = (float*)peer_ptrs[r];
  }
  B2C_REQUIRE(((uintptr_t)multicast_ptr & 15) == 0, "allreduce_symm_fused: misaligned multicast pointer");
  a.mc = (float*)multicast_ptr;
  a.flag_off = flag_offset;
  const size_t n4 = count / 4, base4 = n4 / world, rem = n4 % world;
  const size_t my4 = base4 + ((size_t)rank < rem ? 1 : 0);
  a.begin4 = begin / 4 + base4 * rank + ((size_t)rank < rem ? (size_t)rank : rem);
  a.end4 = a.begin4 + my4;
  a.scale = scale; a.rank = rank; a.world = world;
  size_t blocks = (my4 + 1023) / 1024;            // >= 4 float4 per thread in flight
  static const int cap_default = [] { const char* e = getenv("B2C_AR_CAP"); const int v = e ? atoi(e) : 148; return v > 0 ? v : 148; }();
  const size_t cap = max_blocks > 0 ? (size_t)max_blocks : (size_t)cap_default;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  allreduce_fused_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  return check_launch("allreduce_symm_fused");
}

// Range form: elements [begin, begin + count) of the symmetric buffer (both multiples of 4); the float4s are split over
// the ranks as evenly as possible (the first `rem` ranks take one more); at most max_blocks CTAs, so that the kernel can
// run beside compute kernels (a gradient bucket reduced while the rest of the backward pass still runs).
extern "C" int b2c_allreduce_symm_range(void* const* peer_ptrs, void* multicast_ptr, int rank, int world, size_t begin,
                                        size_t count, float scale, int max_blocks, void* stream) {
  B2C_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "allreduce_symm_range: bad rank / world");
  B2C_REQUIRE(count > 0 && count % 4 == 0 && begin % 4 == 0, "allreduce_symm_range: begin and count must be multiples of 4");
  B2C_REQUIRE(multicast_ptr || peer_ptrs, "allreduce_symm_range: need a multicast pointer or the peer pointers");
  const size_t n4 = count / 4, base4 = n4 / world, rem = n4 % world;
  const size_t my4 = base4 + ((size_t)rank < rem ? 1 : 0);
  const size_t begin4 = begin / 4 + base4 * rank + ((size_t)rank < rem ? (size_t)rank : rem), end4 = begin4 + my4;
  size_t blocks = (my4 + 255) / 256;
  const size_t cap = max_blocks > 0 ? (size_t)max_blocks : 148 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;      // (an empty slice still launches: keeps the launch sequence identical on every rank)
  if (multicast_ptr) {
    B2C_REQUIRE(((uintptr_t)multicast_ptr & 15) == 0, "allreduce_symm_range: misaligned multicast pointer");
    allreduce_multimem_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)multicast_ptr, begin4, end4, scale);
  } else {
    PeerPtrs pp;
    for (int r = 0; r < world; ++r) {
      B2C_REQUIRE(peer_ptrs[r] && ((uintptr_t)peer_ptrs[r] & 15) == 0, "allreduce_symm_range: bad peer pointer %d", r);
      pp.p[r] = (float*)peer_ptrs[r];
    }
    allreduce_peer_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pp, world, begin4, end4, scale);
  }
  count_launch();
  return check_launch("allreduce_symm_range");
}

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
