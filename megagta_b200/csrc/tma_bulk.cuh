// tma_bulk.cuh -- 1-D bulk asynchronous copies (the TMA engine without a tensor map: cp.async.bulk, SASS UBLKCP) and the
// mbarrier that signals their completion.  One elected thread arms the barrier with the byte count and issues the copies;
// every thread of the CTA then waits on the barrier's phase.  No LSU instruction, no register, no address arithmetic per
// element: a tile of an SoA item array is IW copies of one contiguous run each.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mgta {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

// arrive (1) and add `bytes` to the transaction count the current phase waits for
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar` (complete_tx)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace mgta
