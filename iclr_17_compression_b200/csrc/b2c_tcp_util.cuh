// Device helpers shared by the persistent tensor-core kernels (b2c_conv_tcp.cu, b2c_conv1_tcp.cu): swizzled 16-byte
// stores, packed fp32x2 arithmetic, approximate sqrt / rsqrt, pair-scope mbarrier arrivals and TMA-store groups.
#pragma once
#include "b2c_tc.cuh"
#include "b2c_tc_im2col.cuh"

namespace b2c {
namespace tc {

__device__ __forceinline__ void sts128(uint8_t* slot, int r, int cidx, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cidx ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(slot + off) = make_uint4(a, b, c, d);
}
// packed fp32x2 arithmetic (FADD2 / FMUL2): halves the ALU instruction count of the epilogues
__device__ __forceinline__ uint64_t pk2(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t pkf2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_fast(float x) {   // MUFU.RSQ without the denormal fix-up code around rsqrtf()
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Arrive on the mbarrier at the same offset in CTA 0 of the pair with the DEFAULT semantics (release at CTA scope),
// as CUTLASS' ClusterBarrier::arrive does.  The explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR
// (~1.1 k clk on the critical path of every x^2 / accumulator hand-off); the data these arrivals publish is this
// CTA's own shared memory / TMEM, already ordered by fence.proxy.async / tcgen05.fence + the named barrier.
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// NON-blocking phase test for polling (try_wait may suspend the warp for a system-dependent time when the phase is
// not complete - that suspended the MMA warp for 2-3.5 k clk at every tile boundary)
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity, bool cluster_scope) {
  uint32_t ok;
  if (cluster_scope)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
// L2 prefetch of one activation box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// lo = bf16x2(x - float(hi)) for a pair: the two residuals in one packed subtraction (the hi / lo split is 40 % of the
// instructions of the operand builders and epilogues)
__device__ __forceinline__ uint32_t lo_of(float a, float b, uint32_t h) {
  float r0, r1;
  upk2(sub2(pkf2(a, b), pk2(h << 16, h & 0xffff0000u)), r0, r1);
  return cvt_bf16x2(r0, r1);
}

}  // namespace tc
}  // namespace b2c
