/*
 * dsim_async.h — asynchronous global -> shared copies of sm_90+/sm_100a used by k_move:
 *   - bulk copies by the TMA engine (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier in shared memory:
 *     one instruction moves a whole reach row, no registers, no per-thread address arithmetic;
 *   - per-thread 16-byte gathers (cp.async, SASS LDGSTS) for the candidate views, which are scattered.
 * In the CPU debugging build of the test-suite (-DDSIM_EMU) they are plain synchronous copies.
 */
#ifndef DSIM_ASYNC_H
#define DSIM_ASYNC_H

#include "dsim_rt.h"

struct __align__(8) dsim_mbar { unsigned long long state; };

#if defined(DSIM_EMU)
static inline void dsim_mbar_init(dsim_mbar *b, unsigned) { b->state = 0; }
static inline void dsim_mbar_expect_tx(dsim_mbar *, unsigned) {}
static inline void dsim_mbar_wait(dsim_mbar *, unsigned) { __syncwarp(); } /* the copies are made by the lanes themselves: wait for all of them */
static inline void dsim_bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, dsim_mbar *) { std::memcpy(smem_dst, gsrc, bytes); }
static inline void dsim_cp_async16(void *smem_dst, const void *gsrc) { std::memcpy(smem_dst, gsrc, 16); }
static inline void dsim_cp_async8(void *smem_dst, const void *gsrc) { std::memcpy(smem_dst, gsrc, 8); }
static inline void dsim_cp_async4(void *smem_dst, const void *gsrc) { std::memcpy(smem_dst, gsrc, 4); }
static inline unsigned long long dsim_policy_evict_last() { return 0; }
static inline unsigned long long dsim_policy_evict_first() { return 0; }
static inline void dsim_bulk_g2s_hint(void *smem_dst, const void *gsrc, unsigned bytes, dsim_mbar *, unsigned long long) { std::memcpy(smem_dst, gsrc, bytes); }
static inline void dsim_cp_async16_hint(void *smem_dst, const void *gsrc, unsigned long long) { std::memcpy(smem_dst, gsrc, 16); }
static inline void dsim_cp_async8_hint(void *smem_dst, const void *gsrc, unsigned long long) { std::memcpy(smem_dst, gsrc, 8); }
static inline void dsim_cp_async4_hint(void *smem_dst, const void *gsrc, unsigned long long) { std::memcpy(smem_dst, gsrc, 4); }
static inline void dsim_cp_async_commit() {}
static inline void dsim_cp_async_wait_all() {}
static inline void dsim_fence_proxy_async() {}
#else
__device__ __forceinline__ unsigned dsim_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
/* one arrival per phase: the thread that announces the byte count */
__device__ __forceinline__ void dsim_mbar_init(dsim_mbar *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsim_smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); /* visible to the async proxy */
}
__device__ __forceinline__ void dsim_mbar_expect_tx(dsim_mbar *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsim_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dsim_mbar_wait(dsim_mbar *b, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DSIM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DSIM_DONE;\n"
        "bra DSIM_WAIT;\n"
        "DSIM_DONE:\n"
        "}\n" ::"r"(dsim_smem_u32(b)),
        "r"(phase)
        : "memory");
}
/* TMA bulk copy: 16-byte aligned source and destination, size a multiple of 16 */
__device__ __forceinline__ void dsim_bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, dsim_mbar *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dsim_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(dsim_smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void dsim_cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dsim_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void dsim_cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dsim_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void dsim_cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dsim_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
/* L2 eviction priorities (createpolicy + .L2::cache_hint): the small arrays every decision gathers from (candidate views,
 * band records: 64 MB at 2 M nodes) are kept in L2 ahead of the reach rows and history rings that stream through it */
__device__ __forceinline__ unsigned long long dsim_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long dsim_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void dsim_bulk_g2s_hint(void *smem_dst, const void *gsrc, unsigned bytes, dsim_mbar *b, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     dsim_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(dsim_smem_u32(b)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void dsim_cp_async16_hint(void *smem_dst, const void *gsrc, unsigned long long pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dsim_smem_u32(smem_dst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void dsim_cp_async8_hint(void *smem_dst, const void *gsrc, unsigned long long pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(dsim_smem_u32(smem_dst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void dsim_cp_async4_hint(void *smem_dst, const void *gsrc, unsigned long long pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(dsim_smem_u32(smem_dst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void dsim_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void dsim_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
/* orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (TMA) accesses */
__device__ __forceinline__ void dsim_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#endif
