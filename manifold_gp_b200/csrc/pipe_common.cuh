// mbarrier / bulk-copy (TMA) / cp.async helpers shared by the pipelined SpMM kernels (lap_spmm_pipe.cu, lap_spmm_wi.cu).
#pragma once

#include "common.cuh"

namespace mgp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.release.cta.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// Bounded wait: a logic error traps (the launch fails with an error) instead of hanging the GPU.
// Round 2: the r01 loop (try_wait + spin counter + compare + branch, no suspend hint) was the single hottest instruction
// block of the C = 16 SpMM -- ncu source page of lap_spmm_wi_kernel<float,16>: 4.15 M + 1.92 M iterations x 6 instructions =
// 28 % of all issued warp instructions were producers / consumers polling a barrier, on a kernel whose issue slots are 68 %
// busy.  Now: one plain try_wait on the fast path; on the slow path try_wait with a suspend-time hint (the warp sleeps in
// hardware until the phase completes or the hint expires) and a wall-clock guard read only there.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  if (done) return;
  uint64_t t0 = 0;
  for (;;) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.b32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(200000u)      // suspend-time hint in ns
        : "memory");
    if (done) return;
    uint64_t now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    else if (now - t0 > 20000000000ull) __trap();   // 20 s
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// the mbarrier receives one arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

}  // namespace mgp
