// Blackwell/Hopper asynchronous-copy plumbing used by the scan and aggregate kernels: 1-D bulk copies global -> shared
// executed by the TMA unit (cp.async.bulk, SASS UBLKCP) that signal completion on a shared-memory mbarrier
// (SASS SYNCS), plus the mbarrier operations of a producer / consumer pipeline. The streamed column vectors are
// contiguous per chunk, so no tensor map is needed: source, destination and size only have to be multiples of 16 bytes
// (segment buffers are >= 16-byte aligned with a 64-byte readable tail, see Arena / hyb_blocks_upload).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace hyb {

__device__ __forceinline__ uint32_t shared_address(const void* pointer) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(pointer));
}

__device__ __forceinline__ void mbarrier_init(unsigned long long* barrier, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(shared_address(barrier)), "r"(arrivals) : "memory");
}

// Makes the initialised barriers visible to the async proxy (the TMA unit) before the first copy is issued.
__device__ __forceinline__ void mbarrier_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbarrier_arrive(unsigned long long* barrier) {
  asm volatile("{\n\t.reg .b64 state;\n\tmbarrier.arrive.shared::cta.b64 state, [%0];\n\t}" ::"r"(shared_address(barrier))
               : "memory");
}

// One arrival that also announces `bytes` of asynchronous copies which will complete on this barrier.
__device__ __forceinline__ void mbarrier_arrive_expect_tx(unsigned long long* barrier, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 state;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 state, [%0], %1;\n\t}" ::"r"(
                   shared_address(barrier)),
               "r"(bytes)
               : "memory");
}

// Blocks until the phase with the given parity has completed (try_wait suspends the thread in hardware between polls).
__device__ __forceinline__ void mbarrier_wait(unsigned long long* barrier, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred done;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1, 0x989680;\n\t"  // suspend-time hint: do not busy-poll
      "@done bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(shared_address(barrier)),
      "r"(parity)
      : "memory");
}

// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned); completes `bytes` on `barrier`.
__device__ __forceinline__ void bulk_copy_to_shared(void* shared_destination, const void* global_source, uint32_t bytes,
                                                    unsigned long long* barrier) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   shared_address(shared_destination)),
               "l"(global_source), "r"(bytes), "r"(shared_address(barrier))
               : "memory");
}

}  // namespace hyb
