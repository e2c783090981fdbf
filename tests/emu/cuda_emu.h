/*
 * cuda_emu.h — a small CUDA-on-CPU shim.  TEST INFRASTRUCTURE ONLY.
 *
 * This container has no GPU.  To debug the warp-cooperative kernels of
 * dispersal-simulation_b200/csrc before spending GPU minutes, the very same .cu sources are
 * compiled by g++ with -DDSIM_EMU against this header: every CUDA thread becomes a fiber, blocks
 * run one after another, and warp collectives (__shfl_sync, __ballot_sync, __syncwarp, ...) are
 * rendezvous points that CHECK their masks (a lane that exited or that sits in a different
 * collective is reported, not tolerated).  The result, tests/emu/libdsim_emu.so, is loaded only by
 * tests marked `emu`; the product binding never looks for it and the product has no CPU fallback.
 *
 * What it does not model: concurrency between warps/blocks (blocks and warps interleave only at
 * collectives, so data races stay hidden), memory alignment faults, shared-memory limits, timing.
 */
#ifndef CUDA_EMU_H
#define CUDA_EMU_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

/* ---- per-thread context ------------------------------------------------------------------- */
namespace emu {
struct ThreadCtx {
    uint3 tid;
    uint3 bid;
    unsigned linear; /* thread index within the block */
};
extern ThreadCtx *cur;
extern dim3 g_block, g_grid;
extern unsigned char *dyn_smem;

enum Op { OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_ANY, OP_ALL, OP_SYNCWARP };
unsigned long long warp_collective(int op, unsigned mask, unsigned long long value, int arg);
void block_barrier();
void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
unsigned long long launches();
} // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::g_block)
#define gridDim (emu::g_grid)
#define warpSize 32

/* ---- warp collectives --------------------------------------------------------------------- */
template <typename T> static inline unsigned long long emu_pack(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    unsigned long long u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <typename T> static inline T emu_unpack(unsigned long long u) {
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
/* `arg` packs the shuffle operand (low 16 bits) and the segment width (high 16 bits) */
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return emu_unpack<T>(emu::warp_collective(emu::OP_SHFL, mask, emu_pack(v), (src & 0xFFFF) | (width << 16)));
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return emu_unpack<T>(emu::warp_collective(emu::OP_SHFL_UP, mask, emu_pack(v), (int)(d & 0xFFFF) | (width << 16)));
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return emu_unpack<T>(emu::warp_collective(emu::OP_SHFL_DOWN, mask, emu_pack(v), (int)(d & 0xFFFF) | (width << 16)));
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    return emu_unpack<T>(emu::warp_collective(emu::OP_SHFL_XOR, mask, emu_pack(v), (lanemask & 0xFFFF) | (width << 16)));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    return (unsigned)emu::warp_collective(emu::OP_BALLOT, mask, pred ? 1 : 0, 0);
}
static inline int __any_sync(unsigned mask, int pred) {
    return (int)emu::warp_collective(emu::OP_ANY, mask, pred ? 1 : 0, 0);
}
static inline int __all_sync(unsigned mask, int pred) {
    return (int)emu::warp_collective(emu::OP_ALL, mask, pred ? 1 : 0, 0);
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_collective(emu::OP_SYNCWARP, mask, 0, 0); }
static inline void __syncthreads() { emu::block_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

/* ---- integer intrinsics --------------------------------------------------------------------- */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }

/* ---- atomics (single OS thread: plain read-modify-write) -------------------------------------- */
template <typename T> struct emu_id { typedef T type; };
#define EMU_V typename emu_id<T>::type
template <typename T> static inline T atomicAdd(T *p, EMU_V v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicSub(T *p, EMU_V v) { T o = *p; *p = o - v; return o; }
template <typename T> static inline T atomicExch(T *p, EMU_V v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, EMU_V v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T *p, EMU_V v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicMax(T *p, EMU_V v) { T o = *p; *p = std::max<T>(o, v); return o; }
template <typename T> static inline T atomicMin(T *p, EMU_V v) { T o = *p; *p = std::min<T>(o, v); return o; }
template <typename T> static inline T atomicCAS(T *p, EMU_V cmp, EMU_V v) { T o = *p; if (o == cmp) *p = v; return o; }

/* ---- runtime API ---------------------------------------------------------------------------- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct emuEvent *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaHostAllocDefault = 0 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
    size_t totalGlobalMem, sharedMemPerBlockOptin;
};
static inline const char *cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    std::memset(p, 0, sizeof *p);
    std::strcpy(p->name, "emulated sm_100a");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 4;
    p->totalGlobalMem = (size_t)8 << 30; p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    *p = (T *)std::malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { *least = 0; *greatest = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int) { return cudaStreamCreateWithFlags(s, flags); }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

#endif
