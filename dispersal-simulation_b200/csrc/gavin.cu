/*
 * gavin.cu — the gavin2017processbased model batched over a parameter sweep (include/gavin.h), sm_100a.
 *
 * One WARP per sweep member.  The model is sequential by construction — a language grows cell after cell with
 * the fractional growth carried from one cell to the next (language.py:18-33), languages follow one another
 * (model.py:352-368) — so the parallelism is the sweep (members are independent) and, inside a member, the seven
 * candidates of a cell's neighbourhood: lane 0 holds the cell, lanes 1..6 its neighbours; they fetch their
 * cell's state together, draw their permutation keys together and divide their D'Hondt quotients together; the
 * ordered decisions (arg-max with ties to the earlier candidate, the visiting order of `distribute`, the
 * insertion order of the `grow_into` counter) are taken warp-uniformly from shuffled values.
 * Every rule cites the reference line it restates; canonical orders and draws: oracle/gavin.py.
 * Floating point: IEEE doubles in the reference's operation order (-fmad=false).
 */
#include "../../include/gavin.h"
#include "../../include/dsim.h"
#include "dsim_rng.h"
#include "dsim_rt.h"

#include <algorithm>
#include <string>
#include <vector>

enum { GAVIN_TAG_START = 1, GAVIN_TAG_LANGCAP = 2, GAVIN_TAG_PERM = 3, GAVIN_TAG_SEED = 4 };

struct GavinArgs {
    uint32_t R, C, n, S, n_caps, max_lang;
    const double *popcap, *gf, *caps;
    const uint64_t *seed;
    int32_t *language, *population;
    uint32_t *n_languages;
    int32_t *cells, *grow_cnt, *touched; /* [S][n] scratch: cells of the growing language, grow_into counts and key order */
    uint8_t *zone;                       /* [S][n] expansion zone marks */
    unsigned long long *counters;        /* [0] grow calls, [1] cells visited */
};

/* Grid.neighbors, model.py:236-262: neighbour L of cell c, the cell itself where the grid ends */
__device__ __forceinline__ uint32_t gavin_nb(uint32_t c, uint32_t L, uint32_t R, uint32_t C) {
    const uint32_t RC = R * C, m = c >= RC ? 1u : 0u, r = c - m * RC;
    const int i = (int)(r / C), j = (int)(r % C);
    int nm, ni, nj;
    if (m == 0u) {
        switch (L) {
        case 0: nm = 0; ni = i; nj = j + 1; break;
        case 1: nm = 1; ni = i; nj = j; break;
        case 2: nm = 1; ni = i; nj = j - 1; break;
        case 3: nm = 0; ni = i; nj = j - 1; break;
        case 4: nm = 1; ni = i - 1; nj = j - 1; break;
        default: nm = 1; ni = i - 1; nj = j; break;
        }
    } else {
        switch (L) {
        case 0: nm = 1; ni = i; nj = j + 1; break;
        case 1: nm = 0; ni = i; nj = j + 1; break;
        case 2: nm = 0; ni = i; nj = j; break;
        case 3: nm = 1; ni = i; nj = j - 1; break;
        case 4: nm = 0; ni = i + 1; nj = j; break;
        default: nm = 0; ni = i + 1; nj = j + 1; break;
        }
    }
    if (ni < 0 || ni >= (int)R || nj < 0 || nj >= (int)C) return c; /* `g or self`, model.py:258 */
    return (uint32_t)nm * RC + (uint32_t)ni * C + (uint32_t)nj;
}

__device__ __forceinline__ dsim_u4 gavin_draw(uint64_t seed, uint32_t tag, uint32_t x, uint32_t y, uint32_t z) {
    return dsim_philox((uint32_t)seed, (uint32_t)(seed >> 32), tag, x, y, z);
}
__device__ __forceinline__ uint64_t gavin_bounded(uint64_t word, uint64_t n) { return __umul64hi(word, n); }
__device__ __forceinline__ int32_t gavin_founders(double popcap) { /* max(1, min(10, int(popcap))), language.py:11 */
    const int32_t k = popcap >= 10.0 ? 10 : (int32_t)popcap;
    return k < 1 ? 1 : k;
}

__global__ void __launch_bounds__(128) k_gavin(GavinArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = a.n;
    unsigned long long my_calls = 0, my_cells = 0;

    for (uint32_t member = gwarp; member < a.S; member += nwarps) {
        const double *pc = a.popcap + (size_t)member * n;
        int32_t *lang = a.language + (size_t)member * n, *pop = a.population + (size_t)member * n;
        int32_t *cells = a.cells + (size_t)member * n, *cnt = a.grow_cnt + (size_t)member * n, *touched = a.touched + (size_t)member * n;
        uint8_t *zone = a.zone + (size_t)member * n;
        const double gf = a.gf[member];
        const uint64_t seed = a.seed[member];
        for (uint32_t c = lane; c < n; c += 32u) {
            lang[c] = 0;
            pop[c] = 0;
            cnt[c] = 0;
            zone[c] = 0;
        }
        __syncwarp();

        /* a random starting cell that can hold an individual, model.py:346-349 */
        uint32_t c0 = 0xFFFFFFFFu;
        if (lane == 0) {
            for (uint32_t k = 0; k < (1u << 20); ++k) {
                const dsim_u4 w = gavin_draw(seed, GAVIN_TAG_START, k, 0u, 0u);
                const uint32_t m = (uint32_t)gavin_bounded((uint64_t)w.x << 32, 2u), i = (uint32_t)gavin_bounded((uint64_t)w.y << 32, a.R),
                               j = (uint32_t)gavin_bounded((uint64_t)w.z << 32, a.C);
                const uint32_t c = m * a.R * a.C + i * a.C + j;
                if (pc[c] >= 1.0) {
                    c0 = c;
                    break;
                }
            }
        }
        c0 = __shfl_sync(DSIM_FULL, c0, 0);
        if (c0 == 0xFFFFFFFFu) {
            if (lane == 0) a.n_languages[member] = 0u;
            continue;
        }
        uint32_t lid = 1u, n_cells = 1u;
        double cap;
        { /* Language.__init__, language.py:9-16 */
            if (lane == 0) {
                pop[c0] = gavin_founders(pc[c0]);
                lang[c0] = (int32_t)lid;
                cells[0] = (int32_t)c0;
            }
            const dsim_u4 w = gavin_draw(seed, GAVIN_TAG_LANGCAP, lid, 0u, 0u);
            cap = a.caps[gavin_bounded(((uint64_t)w.y << 32) | w.x, a.n_caps)]; /* random_popcap, popcaps.py:16-17 */
        }
        __syncwarp();

        for (;;) { /* one language after the other, model.py:352-368 */
            for (uint32_t call = 0;; ++call) { /* Language.grow until StopIteration, model.py:354-361 */
                uint32_t n_touched = 0;
                double growth = 0.0;
                for (uint32_t ordinal = 0; ordinal < n_cells; ++ordinal) { /* for cell in self.cells, language.py:21 */
                    const uint32_t cell = (uint32_t)cells[ordinal];
                    /* lane 0: the cell; lanes 1..6: its neighbours (cell.neighbors(), language.py:27) */
                    uint32_t q = cell;
                    double pcq = 0.0;
                    int32_t popq = 0;
                    bool ok = false;
                    if (lane <= 6u) {
                        if (lane > 0u) q = gavin_nb(cell, lane - 1u, a.R, a.C);
                        const int32_t lq = lang[q];
                        pcq = pc[q];
                        popq = pop[q];
                        ok = lane > 0u && (lq == (int32_t)lid || lq == 0) && pcq >= 1.0; /* model.py:259-262 */
                    }
                    /* region sums in neighbour order; a neighbour listed twice counts twice (language.py:27-30) */
                    double region_popcap = __shfl_sync(DSIM_FULL, pcq, 0);
                    long long region_population = __shfl_sync(DSIM_FULL, popq, 0);
                    const int32_t cell_pop = (int32_t)region_population;
                    bool first = ok && q != cell; /* member of the region set: first occurrence only */
                    for (uint32_t L = 1; L <= 6u; ++L) {
                        const bool o = __shfl_sync(DSIM_FULL, (int)ok, (int)L) != 0;
                        const double p = __shfl_sync(DSIM_FULL, pcq, (int)L);
                        const int32_t pp = __shfl_sync(DSIM_FULL, popq, (int)L);
                        const uint32_t qq = __shfl_sync(DSIM_FULL, q, (int)L);
                        if (o) {
                            region_popcap = region_popcap + p;
                            region_population += pp;
                            if (L < lane && qq == q) first = false;
                        }
                    }
                    growth = growth + gf * (double)cell_pop * (1.0 - (double)region_population / region_popcap); /* language.py:32 */

                    /* distribute, language.py:46-84: candidates = [cell, distinct neighbours in list order] */
                    const bool is_cand = lane == 0u || (lane <= 6u && first);
                    const uint32_t cand_mask = __ballot_sync(DSIM_FULL, is_cand);
                    const uint32_t my_index = (uint32_t)__popc(cand_mask & ((1u << lane) - 1u));
                    const double space = pcq - (double)popq; /* growth_space, language.py:53 */
                    uint64_t key = 0;
                    if (is_cand && lane > 0u) {
                        const dsim_u4 w = gavin_draw(seed, GAVIN_TAG_PERM, lid, call, 8u * ordinal + my_index);
                        key = ((uint64_t)w.y << 32) | w.x;
                    }
                    uint32_t pos = 0; /* place in the visiting order [0] + shuffled rest, language.py:57 */
                    if (lane > 0u) pos = 1u;
                    for (uint32_t L = 1; L <= 6u; ++L) {
                        const uint64_t kk = __shfl_sync(DSIM_FULL, key, (int)L);
                        if (((cand_mask >> L) & 1u) && is_cand && lane > 0u && L != lane && (kk < key || (kk == key && L < lane))) pos += 1u;
                    }
                    /* dhondt, language.py:86-99 */
                    uint32_t result = 0, divisor = 1, total = 0;
                    while ((double)total < growth && total < (1u << 24)) {
                        const double quot = space / (double)divisor;
                        int best_lane = 0;
                        double best = __shfl_sync(DSIM_FULL, quot, 0);
                        for (uint32_t L = 1; L <= 6u; ++L) {
                            const double qv = __shfl_sync(DSIM_FULL, quot, (int)L);
                            if (((cand_mask >> L) & 1u) && qv > best) { /* ties: the earlier candidate */
                                best = qv;
                                best_lane = (int)L;
                            }
                        }
                        if ((int)lane == best_lane) {
                            result += 1u;
                            divisor += 1u;
                        }
                        total += 1u;
                    }
                    /* visiting loop, language.py:68-81; grow_into += more_grow_into, language.py:35 */
                    const uint32_t n_cand = (uint32_t)__popc(cand_mask);
                    for (uint32_t s = 0; s < n_cand; ++s) {
                        const uint32_t who = __ballot_sync(DSIM_FULL, is_cand && pos == s);
                        const int src = __ffs((int)who) - 1;
                        const uint32_t cg = __shfl_sync(DSIM_FULL, result, src);
                        const uint32_t target = __shfl_sync(DSIM_FULL, q, src);
                        if (cg > 0u) {
                            const int32_t old = cnt[target];
                            __syncwarp();
                            if (lane == 0) {
                                if (old == 0) touched[n_touched] = (int32_t)target; /* a new key of the counter */
                                cnt[target] = old + (int32_t)cg;
                            }
                            if (old == 0) n_touched += 1u;
                        }
                        growth = growth - (double)cg;
                        if (growth < 1.0) break;
                    }
                    __syncwarp();
                }
                my_calls += 1ull;
                my_cells += n_cells;
                /* apply the growth in the counter's key order, language.py:36-44 */
                bool stop = n_touched == 0u; /* "No more growth" */
                uint32_t applied = 0;
                for (; applied < n_touched; ++applied) {
                    const uint32_t target = (uint32_t)touched[applied];
                    const int32_t g = cnt[target];
                    const bool fresh = lang[target] != (int32_t)lid;
                    __syncwarp();
                    if (lane == 0) {
                        pop[target] += g;
                        if (fresh) {
                            lang[target] = (int32_t)lid;
                            cells[n_cells] = (int32_t)target;
                        }
                    }
                    if (fresh) n_cells += 1u;
                    cap = cap - (double)g;
                    if (cap <= 0.0) { /* "popcap reached": the rest of the counter is dropped */
                        stop = true;
                        break;
                    }
                }
                __syncwarp();
                for (uint32_t t = lane; t < n_touched; t += 32u) cnt[touched[t]] = 0;
                __syncwarp();
                if (stop) break;
            }
            if (a.max_lang != 0u && lid >= a.max_lang) break;
            /* expansion zone: free neighbours of filled cells that can hold an individual, model.py:362-366 */
            for (uint32_t c = lane; c < n; c += 32u) {
                const int32_t lc = lang[c];
                if (lc == 0) continue;
                for (uint32_t L = 0; L < 6u; ++L) {
                    const uint32_t q = gavin_nb(c, L, a.R, a.C);
                    if (lang[q] == 0 && pc[q] >= 1.0) zone[q] = 1;
                }
            }
            __syncwarp();
            uint32_t total = 0;
            for (uint32_t base = 0; base < n; base += 32u) total += (uint32_t)__popc(__ballot_sync(DSIM_FULL, base + lane < n && zone[base + lane] != 0));
            if (total == 0u) break;
            const dsim_u4 w = gavin_draw(seed, GAVIN_TAG_SEED, lid, 0u, 0u);
            uint32_t pick = (uint32_t)gavin_bounded(((uint64_t)w.y << 32) | w.x, total); /* free[randint(len(free))], ascending cells */
            uint32_t next = 0;
            for (uint32_t base = 0; base < n; base += 32u) {
                uint32_t m = __ballot_sync(DSIM_FULL, base + lane < n && zone[base + lane] != 0);
                const uint32_t k = (uint32_t)__popc(m);
                if (pick < k) {
                    for (; pick > 0u; --pick) m &= m - 1u;
                    next = base + (uint32_t)__ffs((int)m) - 1u;
                    break;
                }
                pick -= k;
            }
            __syncwarp();
            for (uint32_t c = lane; c < n; c += 32u) zone[c] = 0;
            lid += 1u;
            n_cells = 1u;
            if (lane == 0) {
                pop[next] = gavin_founders(pc[next]);
                lang[next] = (int32_t)lid;
                cells[0] = (int32_t)next;
            }
            const dsim_u4 wc = gavin_draw(seed, GAVIN_TAG_LANGCAP, lid, 0u, 0u);
            cap = a.caps[gavin_bounded(((uint64_t)wc.y << 32) | wc.x, a.n_caps)];
            __syncwarp();
        }
        if (lane == 0) a.n_languages[member] = lid;
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(&a.counters[0], my_calls);
        atomicAdd(&a.counters[1], my_cells);
    }
}

/* ---- host side ----------------------------------------------------------------------------------- */
static std::string g_gavin_error;
const char *gavin_last_error(void) { return g_gavin_error.c_str(); }

#define GCK(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            g_gavin_error = std::string(#call) + " failed: " + cudaGetErrorString(e_);                    \
            for (void *p_ : allocs) cudaFree(p_);                                                         \
            return DSIM_ERR_CUDA;                                                                       \
        }                                                                                                 \
    } while (0)

int gavin_run(const gavin_config *cfg, const double *popcap, const double *growth_factor, const uint64_t *seed,
              const double *lang_caps, int32_t *language, int32_t *population, uint32_t *n_languages, gavin_result *res) {
    std::vector<void *> allocs;
    if (!cfg || !popcap || !growth_factor || !seed || !lang_caps || !language || !population || !n_languages || cfg->rows == 0 ||
        cfg->cols == 0 || cfg->n_members == 0 || cfg->n_lang_caps == 0) {
        g_gavin_error = "gavin_run: invalid argument";
        return DSIM_ERR_INVALID;
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || (int)cfg->device >= n_dev) {
        g_gavin_error = "no CUDA device: the batched model has no CPU fallback";
        return DSIM_ERR_NO_DEVICE;
    }
    GCK(cudaSetDevice((int)cfg->device));
    const size_t n = 2ull * cfg->rows * cfg->cols, S = cfg->n_members, sn = S * n;
    GavinArgs a;
    a.R = cfg->rows;
    a.C = cfg->cols;
    a.n = (uint32_t)n;
    a.S = cfg->n_members;
    a.n_caps = cfg->n_lang_caps;
    a.max_lang = cfg->max_languages;
    double *d_popcap, *d_gf, *d_caps;
    uint64_t *d_seed;
#define GA(ptr, count)                                                    \
    GCK(cudaMalloc(&(ptr), std::max<size_t>((count), 1) * sizeof(*(ptr)))); \
    allocs.push_back(ptr);
    GA(d_popcap, sn) GA(d_gf, S) GA(d_caps, cfg->n_lang_caps) GA(d_seed, S)
    GA(a.language, sn) GA(a.population, sn) GA(a.n_languages, S) GA(a.cells, sn) GA(a.grow_cnt, sn) GA(a.touched, sn) GA(a.zone, sn)
    GA(a.counters, 2)
#undef GA
    GCK(cudaMemcpy(d_popcap, popcap, sn * sizeof(double), cudaMemcpyHostToDevice));
    GCK(cudaMemcpy(d_gf, growth_factor, S * sizeof(double), cudaMemcpyHostToDevice));
    GCK(cudaMemcpy(d_caps, lang_caps, cfg->n_lang_caps * sizeof(double), cudaMemcpyHostToDevice));
    GCK(cudaMemcpy(d_seed, seed, S * sizeof(uint64_t), cudaMemcpyHostToDevice));
    GCK(cudaMemset(a.counters, 0, 2 * sizeof(unsigned long long)));
    a.popcap = d_popcap;
    a.gf = d_gf;
    a.caps = d_caps;
    a.seed = d_seed;
    cudaEvent_t e0, e1;
    GCK(cudaEventCreate(&e0));
    GCK(cudaEventCreate(&e1));
    const uint32_t blocks = (uint32_t)((S + 3) / 4); /* 4 members (warps) per block, one member per warp */
    GCK(cudaEventRecord(e0, 0));
    DSIM_LAUNCH(k_gavin, blocks, 128, 0, 0, a);
    GCK(cudaEventRecord(e1, 0));
    GCK(cudaEventSynchronize(e1));
    GCK(cudaGetLastError());
    float ms = 0.f;
    GCK(cudaEventElapsedTime(&ms, e0, e1));
    GCK(cudaMemcpy(language, a.language, sn * sizeof(int32_t), cudaMemcpyDeviceToHost));
    GCK(cudaMemcpy(population, a.population, sn * sizeof(int32_t), cudaMemcpyDeviceToHost));
    GCK(cudaMemcpy(n_languages, a.n_languages, S * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    unsigned long long counters[2] = {0, 0};
    GCK(cudaMemcpy(counters, a.counters, sizeof counters, cudaMemcpyDeviceToHost));
    if (res) {
        res->device_ms = ms;
        res->grow_calls = counters[0];
        res->cell_updates = counters[1];
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (void *p : allocs) cudaFree(p);
    return DSIM_OK;
}
