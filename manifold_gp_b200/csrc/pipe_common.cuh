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
// Round 2, from the ncu source page of lap_spmm_wi_kernel<float,16> (per-SASS-line executed counts): the r01 loop (try_wait + spin
// counter + compare + branch) ran 4.15 M + 1.92 M iterations per launch -- 28 % of all issued warp instructions were 16 producer
// and 16 consumer warps polling barriers, on a kernel whose issue slots are 68 % busy.  A suspend-time hint on try_wait did not
// change that (the hardware still returns within tens of ns; measured 2.06 M iterations).  Now: one plain try_wait on the fast
// path; on the slow path the warp SLEEPS (nanosleep, no issue slots) between polls, and only one producer warp polls at all
// (lap_spmm_wi.cu releases the others through a named barrier, which blocks in hardware).
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
template <unsigned SLEEP_NS = 96>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_test(bar, parity)) return;
  for (uint32_t spins = 0;; ++spins) {
    __nanosleep(SLEEP_NS);
    if (mbar_test(bar, parity)) return;
    if (spins > (1u << 26)) __trap();               // >= 6 s of sleeping alone
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
