// group_by_patch_ab.cu — A/B for SURVEY.md §8 row N2 / lib.rs:602-612 (families_by_location): how to bring the
// families of a step into per-patch groups.
//   A  the product's counting sort: arrival ranks from one global atomicAdd per family on its node's counter (in the
//      product this is fused into the epilogue of k_move / k_children; timed here as a pass of its own), segment starts
//      claimed by the first arrival of every node with one cursor bump per block (k_segments), scatter (k_scatter)
//   B  sort-by-patch: cub::DeviceRadixSort::SortPairs on (node, family) + head flags + exclusive scan for the segment
//      table (the scan-based ranking; no atomics at all)
//   C  per-patch sum of a per-family double ("effort"): one atomicAdd(double) per family into its node's accumulator
//      against a warp-shuffle segmented sum over the sorted order (B's output)
// Sizes: the populations of BASELINE configs C2 (68 k live families), C3 (1 M), C4 (10 M) at their co-residence.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cudart shared -o gpurun_out/group_by_patch_ab tools/micro/group_by_patch_ab.cu
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_hist(const uint32_t *loc, uint32_t n, uint32_t *cnt, uint32_t *fslot) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) fslot[j] = atomicAdd(&cnt[loc[j]], 1u);
}
__global__ void __launch_bounds__(256) k_seg(const uint32_t *loc, const uint32_t *fslot, uint32_t n, const uint32_t *cnt, uint32_t *seg, uint32_t *cursor) {
    __shared__ uint32_t s_tot[8], s_base[8];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t n_round = (n + 255u) & ~255u;
    for (uint32_t j0 = blockIdx.x * 256u; j0 < n_round; j0 += gridDim.x * 256u) {
        const uint32_t j = j0 + threadIdx.x;
        const bool first = j < n && fslot[j] == 0u;
        uint32_t p = 0, c = 0;
        if (first) { p = loc[j]; c = cnt[p]; }
        uint32_t incl = c;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += o; }
        if (lane == 31u) s_tot[w] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int k = 0; k < 8; ++k) tot += s_tot[k];
            uint32_t base = tot ? atomicAdd(cursor, tot) : 0u;
            for (int k = 0; k < 8; ++k) { s_base[k] = base; base += s_tot[k]; }
        }
        __syncthreads();
        if (first) seg[p] = s_base[w] + incl - c;
    }
}
__global__ void k_scat(const uint32_t *loc, const uint32_t *fslot, uint32_t n, const uint32_t *seg, uint32_t *members) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) members[seg[loc[j]] + fslot[j]] = j;
}
__global__ void k_clear(const uint32_t *loc, uint32_t n, uint32_t *cnt) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) cnt[loc[j]] = 0u;
}
__global__ void k_heads(const uint32_t *keys, uint32_t n, uint32_t *flag) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) flag[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}
__global__ void k_segtab(const uint32_t *keys, const uint32_t *flag, const uint32_t *rank, uint32_t n, uint32_t *seg_start, uint32_t *seg_node) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        if (flag[j]) { seg_start[rank[j]] = j; seg_node[rank[j]] = keys[j]; }
}
__global__ void k_sum_atomic(const uint32_t *loc, const double *v, uint32_t n, double *acc) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) atomicAdd(&acc[loc[j]], v[j]);
}
// warp-shuffle segmented sum over the sorted order: lanes of a warp hold consecutive sorted positions; a segmented
// inclusive scan by key, the last lane of every run adds the run's sum to its node (runs that cross warps: one more add)
__global__ void k_sum_shuffle(const uint32_t *skeys, const uint32_t *sidx, const double *v, uint32_t n, double *acc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x) {
        const bool in = j < n;
        const uint32_t key = in ? skeys[j] : 0xFFFFFFFFu;
        double x = in ? v[sidx[j]] : 0.0;
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, x, d);
            const uint32_t ko = __shfl_up_sync(0xffffffffu, key, d);
            if ((int)lane >= d && ko == key) x += o;
        }
        const uint32_t kn = __shfl_down_sync(0xffffffffu, key, 1);
        if (in && (lane == 31u || kn != key)) atomicAdd(&acc[key], x); // one add per run and warp
    }
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1000.f; }
};

int main() {
    struct Case { const char *name; uint32_t n, n_nodes, occupied; } cases[] = {
        {"C2 (68 k families, rho 1.75)", 68000u, 2000460u, 39000u},
        {"C3 (1 M families, rho 8)", 1000000u, 2000460u, 125000u},
        {"C4 (10 M families, rho 8)", 10000000u, 10000086u, 1250000u}};
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    printf("| case | A: histogram atomics | A: segments | A: scatter | A total (product path: segments + scatter) | B: radix sort pairs | B: heads + scan + table | B total | C: atomicAdd(double) per family | C: shuffle segmented sum |\n|---|---|---|---|---|---|---|---|---|---|\n");
    for (const Case &c : cases) {
        std::vector<uint32_t> occ(c.occupied), loc(c.n);
        uint64_t s = 88172645463325252ull;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
        for (auto &p : occ) p = (uint32_t)(rnd() % c.n_nodes);
        for (auto &l : loc) l = occ[rnd() % c.occupied];
        uint32_t *d_loc, *d_cnt, *d_fslot, *d_seg, *d_members, *d_cursor, *d_keys2, *d_idx, *d_idx2, *d_flag, *d_rank, *d_segstart, *d_segnode;
        double *d_v, *d_acc;
        CK(cudaMalloc(&d_loc, c.n * 4)); CK(cudaMalloc(&d_cnt, (size_t)c.n_nodes * 4)); CK(cudaMalloc(&d_fslot, c.n * 4));
        CK(cudaMalloc(&d_seg, (size_t)c.n_nodes * 4)); CK(cudaMalloc(&d_members, c.n * 4)); CK(cudaMalloc(&d_cursor, 4));
        CK(cudaMalloc(&d_keys2, c.n * 4)); CK(cudaMalloc(&d_idx, c.n * 4)); CK(cudaMalloc(&d_idx2, c.n * 4)); CK(cudaMalloc(&d_flag, c.n * 4));
        CK(cudaMalloc(&d_rank, c.n * 4)); CK(cudaMalloc(&d_segstart, c.n * 4)); CK(cudaMalloc(&d_segnode, c.n * 4));
        CK(cudaMalloc(&d_v, (size_t)c.n * 8)); CK(cudaMalloc(&d_acc, (size_t)c.n_nodes * 8));
        CK(cudaMemcpy(d_loc, loc.data(), c.n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_cnt, 0, (size_t)c.n_nodes * 4)); CK(cudaMemset(d_acc, 0, (size_t)c.n_nodes * 8)); CK(cudaMemset(d_v, 0x3f, (size_t)c.n * 8));
        std::vector<uint32_t> iota(c.n);
        for (uint32_t i = 0; i < c.n; ++i) iota[i] = i;
        CK(cudaMemcpy(d_idx, iota.data(), c.n * 4, cudaMemcpyHostToDevice));
        size_t tmp_bytes = 0, tmp2 = 0;
        int end_bit = 1;
        while ((1ull << end_bit) < c.n_nodes) ++end_bit;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_loc, d_keys2, d_idx, d_idx2, (int)c.n, 0, end_bit);
        cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_flag, d_rank, (int)c.n);
        void *d_tmp;
        CK(cudaMalloc(&d_tmp, tmp_bytes > tmp2 ? tmp_bytes : tmp2));
        const int grid = n_sm * 8;
        Timer t;
        float a_h = 1e9f, a_s = 1e9f, a_c = 1e9f, b_s = 1e9f, b_t = 1e9f, c_a = 1e9f, c_s = 1e9f;
        for (int rep = 0; rep < 7; ++rep) {
            k_clear<<<grid, 256>>>(d_loc, c.n, d_cnt);
            CK(cudaMemset(d_cursor, 0, 4));
            t.start(); k_hist<<<grid, 256>>>(d_loc, c.n, d_cnt, d_fslot); a_h = fminf(a_h, t.stop());
            t.start(); k_seg<<<grid, 256>>>(d_loc, d_fslot, c.n, d_cnt, d_seg, d_cursor); a_s = fminf(a_s, t.stop());
            t.start(); k_scat<<<grid, 256>>>(d_loc, d_fslot, c.n, d_seg, d_members); a_c = fminf(a_c, t.stop());
            t.start(); cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_loc, d_keys2, d_idx, d_idx2, (int)c.n, 0, end_bit); b_s = fminf(b_s, t.stop());
            t.start();
            k_heads<<<grid, 256>>>(d_keys2, c.n, d_flag);
            cub::DeviceScan::ExclusiveSum(d_tmp, tmp2, d_flag, d_rank, (int)c.n);
            k_segtab<<<grid, 256>>>(d_keys2, d_flag, d_rank, c.n, d_segstart, d_segnode);
            b_t = fminf(b_t, t.stop());
            t.start(); k_sum_atomic<<<grid, 256>>>(d_loc, d_v, c.n, d_acc); c_a = fminf(c_a, t.stop());
            t.start(); k_sum_shuffle<<<grid, 256>>>(d_keys2, d_idx2, d_v, c.n, d_acc); c_s = fminf(c_s, t.stop());
        }
        CK(cudaDeviceSynchronize());
        printf("| %s | %.1f us | %.1f us | %.1f us | %.1f us | %.1f us | %.1f us | %.1f us | %.1f us | %.1f us |\n", c.name, a_h, a_s, a_c, a_s + a_c, b_s, b_t,
               b_s + b_t, c_a, c_s);
        cudaFree(d_loc); cudaFree(d_cnt); cudaFree(d_fslot); cudaFree(d_seg); cudaFree(d_members); cudaFree(d_cursor); cudaFree(d_keys2); cudaFree(d_idx);
        cudaFree(d_idx2); cudaFree(d_flag); cudaFree(d_rank); cudaFree(d_segstart); cudaFree(d_segnode); cudaFree(d_v); cudaFree(d_acc); cudaFree(d_tmp);
    }
    return 0;
}
