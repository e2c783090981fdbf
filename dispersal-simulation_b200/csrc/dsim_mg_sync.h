/*
 * dsim_mg_sync.h — flag words of the peer-memory transport (dsim_mg_kernels.cu, k_children).
 *
 * A message is written with plain remote stores into the receiver's buffer (NVLink peer memory mapped through CUDA IPC);
 * the LAST block of the writing kernel to finish then publishes the exchange round in the receiver's flag word
 * (relaxed system-scope store after a system-scope fence); the consuming kernel starts by spinning on its own flag words
 * (ld.acquire.sys, bounded).  In the CPU debugging build of the test-suite these are plain accesses.
 */
#ifndef DSIM_MG_SYNC_H
#define DSIM_MG_SYNC_H

#include "dsim_kernels.h"

#if defined(DSIM_EMU)
static inline void mg_st_flag(uint32_t *p, uint32_t v) { *p = v; }
static inline uint32_t mg_ld_acquire_sys(const uint32_t *p) { return *p; }
static inline void mg_fence_sys() {}
static inline void mg_backoff() {}
#else
/* a flag word is stored relaxed at system scope AFTER one system-scope fence of the publishing thread (mg_last_block):
 * fence + relaxed store is the release pattern, and several flag words (both neighbours, or every rank for SPLIT) then
 * cost one fence — a st.release.sys each would wait for the peer's acknowledgement once per flag */
__device__ __forceinline__ void mg_st_flag(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t mg_ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mg_fence_sys() { __threadfence_system(); }
__device__ __forceinline__ void mg_backoff() { __nanosleep(100); }
#endif

/* Called by every thread of a consuming kernel before it reads a message: thread k < n_slots of each block waits until
 * flag word slot0 + k has reached `round` (rounds only grow; a neighbour that is a step ahead may already have
 * published the next one).  A sender that never arrives raises DSIM_MODEL_COMM_TIMEOUT instead of hanging the device. */
__device__ __forceinline__ void mg_wait_flags(const uint32_t *flags, uint32_t slot0, uint32_t n_slots, uint32_t slot_mask, uint32_t round,
                                              uint32_t *errors) {
    if (threadIdx.x < n_slots && ((slot_mask >> threadIdx.x) & 1u)) {
        uint32_t spins = 0;
        while ((int32_t)(mg_ld_acquire_sys(flags + slot0 + threadIdx.x) - round) < 0) {
            mg_backoff();
            if (++spins > (1u << 25)) { /* several seconds */
                atomicOr(errors, (uint32_t)DSIM_MODEL_COMM_TIMEOUT);
                break;
            }
        }
    }
    __syncthreads(); /* the acquire of the waiting threads orders the whole block's later reads behind the message */
}

#endif
