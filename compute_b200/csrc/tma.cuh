// tma.cuh -- bulk-copy engine (TMA, 1-D cp.async.bulk), mbarrier and named-barrier helpers for sm_100a.
// SASS: cp.async.bulk -> UBLKCP, mbarrier ops -> SYNCS.*, bar.sync/arrive -> BAR.SYNC/ARV.
#pragma once

#include <cuda_runtime.h>

namespace bcb {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// arrive (release) and add `bytes` to the transaction count the current phase waits for
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait (acquire) for the phase with the given parity to complete
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// global -> shared, completion counted in bytes on an mbarrier; addresses and size are multiples of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, bulk-group completion (commit separately); addresses and size are multiples of 16
__device__ __forceinline__ void tma_store_issue(void *gmem_dst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// the same with an L2 eviction-priority policy (createpolicy) for the written lines
__device__ __forceinline__ void tma_store_issue_hint(void *gmem_dst, const void *smem_src, unsigned bytes, unsigned long long policy)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes), "l"(policy)
                 : "memory");
}
// one 16-byte chunk of which only the bytes selected by `mask` (bit i = byte i) are written (PTX 8.6, sm_100+)
__device__ __forceinline__ void tma_store_masked16(void *gmem_dst, const void *smem_src, unsigned mask)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "h"((unsigned short)mask)
                 : "memory");
}
// ask the bulk-copy engine to pull a global range into L2 (no shared-memory destination, nothing to wait for)
__device__ __forceinline__ void tma_prefetch_l2(const void *gmem_src, unsigned bytes, unsigned long long policy)
{
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(gmem_src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, unsigned bytes)
{
    tma_store_issue(gmem_dst, smem_src, bytes);
    tma_commit();
}
// wait until all but the N most recent bulk groups of this thread have finished READING shared memory
template <int N> __device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the bulk-copy engine (async proxy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// named barriers (ids 1..15; 0 is __syncthreads): `count` threads take part, arrive does not wait
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

}  // namespace bcb
