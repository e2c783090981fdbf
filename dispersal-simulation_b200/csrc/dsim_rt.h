/*
 * dsim_rt.h — CUDA runtime include + launch macro.
 *
 * The product is compiled by nvcc for sm_100a.  The same sources also compile with g++ against
 * tests/emu/cuda_emu.h (-DDSIM_EMU) for the CPU debugging harness of the test-suite; that build is
 * never loaded by the product binding.
 */
#ifndef DSIM_RT_H
#define DSIM_RT_H

#if defined(DSIM_EMU)
#include "cuda_emu.h"
#define DSIM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(#kernel, dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })
#define DSIM_DYN_SMEM(name) unsigned char *name = emu::dyn_smem
#else
#include <cuda_runtime.h>
#define DSIM_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DSIM_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#if defined(DSIM_EMU)
#define DSIM_NOINLINE __attribute__((noinline))
#define DSIM_LDCG(p) (*(p))
#else
#define DSIM_NOINLINE __noinline__
#define DSIM_LDCG(p) __ldcg(p) /* read at L2: data another block of the same launch has just written */
#endif

#define DSIM_FULL 0xffffffffu

#endif
