/*
 * dsim_rt.h — CUDA runtime include + launch macro.
 *
 * The product is compiled by nvcc for sm_100a.  The same sources also compile with g++ against
 * tests/emu/cuda_emu.h (-DDSIM_EMU) for the CPU debugging harness of the test-suite; that build is
 * never loaded by the product binding.
 */
#ifndef DSIM_RT_H
#define DSIM_RT_H

#include <stdint.h>

#if defined(DSIM_EMU)
#include "cuda_emu.h"
#define DSIM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(#kernel, dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })
#define DSIM_DYN_SMEM(name) unsigned char *name = emu::dyn_smem
#define DSIM_LAUNCH_PDL DSIM_LAUNCH
static inline void dsim_pdl_wait() {}
static inline void dsim_pdl_trigger() {}
#else
#include <cuda_runtime.h>
#define DSIM_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DSIM_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
/* Programmatic dependent launch (sm_90+): a kernel launched with this attribute may be scheduled while its predecessor
 * in the stream (or in the captured graph) is still draining — its blocks take the SM resources that become free and
 * park at dsim_pdl_wait() until the predecessor has completed and its memory is visible; the launch latency of the small
 * kernels of a step disappears behind the tail of the one before.  ONLY kernels that start with dsim_pdl_wait() may be
 * launched this way.  dsim_pdl_trigger() (after the wait) lets the next kernel in line do the same to this one. */
template <typename... KArgs, typename... Args>
static inline cudaError_t dsim_launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define DSIM_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) dsim_launch_pdl(kernel, (grid), (block), (smem), (stream), __VA_ARGS__)
#if defined(__CUDACC__)
__device__ __forceinline__ void dsim_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void dsim_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
#endif

#if defined(DSIM_EMU)
#define DSIM_NOINLINE __attribute__((noinline))
#define DSIM_LDCG(p) (*(p))
#else
#define DSIM_NOINLINE __noinline__
#define DSIM_LDCG(p) __ldcg(p) /* read at L2: data another block of the same launch has just written */
#endif

/* release / acquire at device scope for hand-offs between kernels that run concurrently (ChildTask.ready) */
#if defined(DSIM_EMU)
static inline void dsim_store_release_u32(uint32_t *p, uint32_t v) { *p = v; }
static inline uint32_t dsim_load_acquire_u32(const uint32_t *p) { return *p; }
static inline void dsim_backoff() {}
#elif defined(__CUDACC__)
__device__ __forceinline__ void dsim_store_release_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t dsim_load_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dsim_backoff() { __nanosleep(64); }
#endif

#define DSIM_FULL 0xffffffffu

#endif
