// Random-sector read rate of HBM (B200): every thread gathers 16-byte words at hashed, sector-aligned offsets of a
// buffer far larger than L2.  Prints giga-sectors/s and the equivalent GB/s for several access widths.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rand_sector_bw rand_sector_bw.cu && ./rand_sector_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int SECTORS>  // consecutive sectors per gather
__global__ void gather(const uint4 *buf, uint64_t n_sectors, uint32_t per_thread, uint32_t *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per_thread; ++k) {
        const uint64_t s = ((uint64_t)mix(t * 7919u + k) * (n_sectors / SECTORS)) >> 32;  // random group of SECTORS sectors
#pragma unroll
        for (int j = 0; j < SECTORS; ++j) { const uint4 v = buf[(s * SECTORS + j) * 2]; acc += v.x; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
int main() {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    uint32_t *out; cudaMalloc(&out, 4);
    const double sizes_gb[] = {0.5, 2, 8, 32, 96};
    for (double gb : sizes_gb) {
        const uint64_t bytes = (uint64_t)(gb * (1ull << 30)); uint4 *buf;
        if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("%.1f GB: alloc failed\n", gb); continue; }
        cudaMemset(buf, 1, bytes);
        const uint64_t n_sectors = bytes / 32;
        const uint32_t blocks = 148 * 8, threads = 256, per = 256;
        for (int w = 1; w <= 4; w *= 4) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (w == 1) gather<1><<<blocks, threads>>>(buf, n_sectors, per, out);
                if (w == 4) gather<4><<<blocks, threads>>>(buf, n_sectors, per, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const double sectors = (double)blocks * threads * per * w;
            printf("buffer %5.1f GB  run of %d sectors: %.1f Gsector/s = %.0f GB/s\n", gb, w, sectors / best * 1e-6, sectors * 32 / best * 1e-6);
        }
        cudaFree(buf);
    }
    return 0;
}
