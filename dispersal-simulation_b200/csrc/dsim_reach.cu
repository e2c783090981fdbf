/*
 * dsim_reach.cu — construction of the reach table, once per graph: on the host (OpenMP, the default) or on the
 * device (a warp per node; second half of this file).
 *
 * For every node: (a) the set of nodes that a Dijkstra search started there ever assigns a score to
 * while it settles nodes in order of distance and stops at the first one beyond
 * MAX_SEASON_RANGE = 320/6*1000/0.7740708353271509 s (lib.rs:1183) — i.e. the nodes within range and
 * the out-neighbours of those, the latter with their tentative distances — minus the entries whose
 * distance is not a normal float (this removes the start node, d = 0; lib.rs:1147); cost in OYR =
 * (1/SECONDS_PER_YEAR*3) * seconds (lib.rs:1139-1140,1148).  (b) the 2-hop out-neighbourhood
 * (lib.rs:1158-1178).  Both sorted by node id: the reference keeps them in hash containers whose
 * iteration order is random, the canonical order is ascending.
 */
#include "dsim_reach.h"
#include "dsim_rt.h"
#include "dsim_math.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <queue>
#include <utility>
#include <cstdlib>
#include <thread>
#include <omp.h>

namespace {

const double kSecondsPerYear = 365.24219 * 24. * 60. * 60.;
const double kMaxSeasonRange = 320. / 6. * 1000. / 0.7740708353271509;

/* open-addressing map node -> (distance, settled) that is cleared by walking its own key list */
struct LocalMap {
    std::vector<uint32_t> key;
    std::vector<double> dist;
    std::vector<uint8_t> settled;
    std::vector<uint32_t> used;
    uint32_t mask;
    explicit LocalMap(uint32_t cap_pow2 = 1024) { resize(cap_pow2); }
    void resize(uint32_t cap) {
        key.assign(cap, 0xFFFFFFFFu);
        dist.assign(cap, 0.0);
        settled.assign(cap, 0);
        mask = cap - 1;
        used.clear();
    }
    void clear() {
        for (uint32_t s : used) key[s] = 0xFFFFFFFFu;
        used.clear();
    }
    uint32_t slot(uint32_t k) const {
        uint32_t s = (k * 2654435761u) & mask;
        while (key[s] != 0xFFFFFFFFu && key[s] != k) s = (s + 1) & mask;
        return s;
    }
    void grow() {
        std::vector<std::pair<uint32_t, std::pair<double, uint8_t>>> items;
        for (uint32_t s : used) items.push_back({key[s], {dist[s], settled[s]}});
        resize((mask + 1) * 2);
        for (auto &it : items) {
            const uint32_t s = slot(it.first);
            key[s] = it.first;
            dist[s] = it.second.first;
            settled[s] = it.second.second;
            used.push_back(s);
        }
    }
};

void reach_of(const dsim_graph *g, uint32_t start, LocalMap &map, std::vector<std::pair<uint32_t, double>> &out) {
    typedef std::pair<double, uint32_t> Item;
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
    map.clear();
    {
        const uint32_t s = map.slot(start);
        map.key[s] = start;
        map.dist[s] = 0.0;
        map.settled[s] = 0;
        map.used.push_back(s);
    }
    heap.push({0.0, start});
    while (!heap.empty()) {
        const Item it = heap.top();
        heap.pop();
        const uint32_t us = map.slot(it.second);
        if (map.settled[us]) continue;
        if (kMaxSeasonRange < it.first) break;
        for (uint64_t e = g->edge_off[it.second]; e < g->edge_off[it.second + 1]; ++e) {
            const uint32_t v = g->edge_dst[e];
            if (map.used.size() * 2 > map.mask) map.grow();
            const uint32_t vs = map.slot(v);
            double nd = it.first + g->edge_sec[e];
            if (map.key[vs] == v) {
                if (map.settled[vs]) continue;
                if (nd < map.dist[vs])
                    map.dist[vs] = nd;
                else
                    nd = map.dist[vs];
            } else {
                map.key[vs] = v;
                map.dist[vs] = nd;
                map.settled[vs] = 0;
                map.used.push_back(vs);
            }
            heap.push({nd, v});
        }
        map.settled[map.slot(it.second)] = 1; /* the map may have grown: look the slot up again */
    }
    const double cost_per_second = 1. / kSecondsPerYear * 3.;
    out.clear();
    for (uint32_t s : map.used)
        if (std::isnormal(map.dist[s])) out.push_back({map.key[s], cost_per_second * map.dist[s]});
    std::sort(out.begin(), out.end());
}

void nearby_of(const dsim_graph *g, uint32_t node, std::vector<uint32_t> &out) {
    out.clear();
    for (uint64_t e = g->edge_off[node]; e < g->edge_off[node + 1]; ++e) {
        const uint32_t nb = g->edge_dst[e];
        for (uint64_t f = g->edge_off[nb]; f < g->edge_off[nb + 1]; ++f) out.push_back(g->edge_dst[f]);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

} // namespace

/* host threads for the one-time table build: DSIM_HOST_THREADS, else all hardware threads (launchers such
 * as torchrun export OMP_NUM_THREADS=1, which would make this step 16x slower for no reason) */
static int host_threads() {
    if (const char *e = std::getenv("DSIM_HOST_THREADS")) {
        const int k = std::atoi(e);
        if (k > 0) return k;
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

int dsim_build_reach(const dsim_graph *g, HostReach *out, std::string *err) {
    const uint32_t n = g->n_nodes;
    omp_set_num_threads(host_threads());
    const uint32_t kChunk = 2048;
    const uint32_t n_chunks = (n + kChunk - 1) / kChunk;
    struct Chunk {
        std::vector<uint32_t> r_len, n_len, r_dst, n_dst, r_sensed;
        std::vector<double> r_cost;
    };
    std::vector<Chunk> chunks(n_chunks);
    int failed = 0;

#pragma omp parallel
    {
        LocalMap map;
        std::vector<std::pair<uint32_t, double>> row;
        std::vector<uint32_t> near;
#pragma omp for schedule(dynamic, 1)
        for (long c = 0; c < (long)n_chunks; ++c) {
            Chunk &ch = chunks[c];
            const uint32_t v0 = (uint32_t)c * kChunk, v1 = std::min(n, v0 + kChunk);
            for (uint32_t v = v0; v < v1; ++v) {
                reach_of(g, v, map, row);
                nearby_of(g, v, near);
                if (near.size() >= 0xFFFFu) {
#pragma omp atomic write
                    failed = 1;
                }
                ch.r_len.push_back((uint32_t)row.size());
                ch.n_len.push_back((uint32_t)near.size());
                for (uint32_t q : near) ch.n_dst.push_back(q);
                /* sensed entries (also in the nearby list) first, then the rest; both ascending */
                uint32_t n_sensed = 0;
                for (int pass = 0; pass < 2; ++pass)
                    for (const auto &qc : row) {
                        const auto it = std::lower_bound(near.begin(), near.end(), qc.first);
                        const bool sensed = (it != near.end() && *it == qc.first);
                        if (sensed != (pass == 0)) continue;
                        ch.r_dst.push_back(qc.first);
                        ch.r_cost.push_back(qc.second);
                        n_sensed += sensed ? 1u : 0u;
                    }
                ch.r_sensed.push_back(n_sensed);
            }
        }
    }
    if (failed) {
        if (err) *err = "a node has 65535 or more nodes in its 2-hop neighbourhood";
        return DSIM_ERR_INVALID;
    }
    /* offsets: reach rows padded to multiples of 8 entries (16-byte aligned starts for every column) */
    out->hdr.assign((size_t)n, HostRowHdr{0, 0, 0, 0});
    out->near_off.assign((size_t)n + 1, 0);
    uint64_t ro = 0, no = 0, n_reach = 0;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t v0 = c * kChunk;
        for (size_t k = 0; k < chunks[c].r_len.size(); ++k) {
            const uint32_t padded = (chunks[c].r_len[k] + 7u) & ~7u;
            if (ro > 0xFFFFFFF0ull) break;
            out->hdr[v0 + k] = HostRowHdr{(uint32_t)ro, padded, chunks[c].r_sensed[k], chunks[c].n_len[k]};
            out->near_off[v0 + k] = (uint32_t)no;
            ro += padded;
            no += chunks[c].n_len[k];
            n_reach += chunks[c].r_len[k];
        }
    }
    if (ro > 0xFFFFFFF0ull || no > 0xFFFFFFF0ull) {
        if (err) *err = "reach table exceeds 2^32 entries";
        return DSIM_ERR_CAPACITY;
    }
    out->near_off[n] = (uint32_t)no;
    out->n_reach = n_reach;
    out->dst.assign(ro, 0xFFFFFFFFu);
    out->cost.assign(ro, std::numeric_limits<double>::infinity());
    out->near_dst.assign(no, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)n_chunks; ++c) {
        const Chunk &ch = chunks[c];
        const uint32_t v0 = (uint32_t)c * kChunk;
        size_t rp = 0, np = 0;
        for (size_t k = 0; k < ch.r_len.size(); ++k) {
            std::copy(ch.r_dst.begin() + rp, ch.r_dst.begin() + rp + ch.r_len[k], out->dst.begin() + out->hdr[v0 + k].start);
            std::copy(ch.r_cost.begin() + rp, ch.r_cost.begin() + rp + ch.r_len[k], out->cost.begin() + out->hdr[v0 + k].start);
            std::copy(ch.n_dst.begin() + np, ch.n_dst.begin() + np + ch.n_len[k], out->near_dst.begin() + out->near_off[v0 + k]);
            rp += ch.r_len[k];
            np += ch.n_len[k];
        }
    }
    /* per-row index for the travel-cost look-ups of remembered patches: 32 buckets of 8 slots, slot =
     * tag << 8 | entry position (0xFFFF = empty); a full bucket overflows into the next one */
    out->hash.assign((size_t)n * 256u, (uint16_t)0xFFFF);
#pragma omp parallel for schedule(static)
    for (long v = 0; v < (long)n; ++v) {
        const HostRowHdr &hd = out->hdr[v];
        if (hd.len > DSIM_ROW_HASH_MAX) continue;
        uint16_t *tbl = out->hash.data() + (size_t)v * 256u;
        for (uint32_t k = 0; k < hd.len; ++k) {
            const uint32_t q = out->dst[hd.start + k];
            if (q == 0xFFFFFFFFu) continue;
            const uint32_t h = dsim_row_hash(q);
            uint32_t b = h >> 8;
            for (;;) {
                uint32_t s = 0;
                while (s < 8u && tbl[b * 8u + s] != 0xFFFF) ++s;
                if (s < 8u) {
                    tbl[b * 8u + s] = (uint16_t)(((h & 0xFFu) << 8) | k);
                    break;
                }
                b = (b + 1u) & 31u;
            }
        }
    }
    return DSIM_OK;
}

/* ------------------------------------------------------------------------------------------ *
 * The same table on the device.  One warp per start node.  The search is label-correcting instead of Dijkstra's
 * ordered one: every node whose tentative distance is within range relaxes its out-edges, again whenever its
 * distance has improved, until nothing changes.  The fixed point is the same — d(q) = min over predecessors u within
 * range of d(u) + w(u, q), evaluated in the same floating-point order along every path, and the same set of nodes
 * ever given a distance (a node relaxes only when its final distance is within range too, since tentative >= final).
 * ------------------------------------------------------------------------------------------ */
#define RB_WARPS 2
#define RB_CAP 512u  /* slots of the per-warp distance table (at most DSIM_REACH_DEV_MAX used) */
#define NB_CAP 1024u /* slots of the per-warp 2-hop set (at most DSIM_NEAR_DEV_MAX used)       */
#define RB_EMPTY 0xFFFFFFFFu
#define RB_INF 0x7FF0000000000000ull

struct ReachWarpMem {
    unsigned long long dist[RB_CAP]; /* bit patterns of non-negative doubles order like the doubles; later: row keys */
    unsigned long long done[RB_CAP]; /* distance at which the slot last relaxed its edges; later: compacted distances */
    uint32_t key[RB_CAP];            /* later: sensed flag of compacted entry i */
    uint32_t ckey[RB_CAP];           /* compacted keys */
    uint32_t nset[NB_CAP];
    uint32_t nkey[DSIM_NEAR_DEV_MAX];
    uint16_t tbl[256];
    uint32_t cnt[4]; /* entries in dist table, entries in 2-hop set, unsupported */
};

__device__ __forceinline__ uint32_t rb_hash(uint32_t q) { return q * 2654435761u; }

template <bool FILL>
__global__ void __launch_bounds__(RB_WARPS * 32) k_reach_build(ReachDevGraph g, ReachDevOut o, double bound, double cost_per_second) {
    __shared__ ReachWarpMem mem[RB_WARPS];
    ReachWarpMem &M = mem[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;

    for (uint32_t v = g.v_lo + gwarp; v < g.v_hi; v += nwarps) {
        for (uint32_t s = lane; s < RB_CAP; s += 32u) {
            M.key[s] = RB_EMPTY;
            M.dist[s] = RB_INF;
            M.done[s] = ~0ull;
        }
        for (uint32_t s = lane; s < NB_CAP; s += 32u) M.nset[s] = RB_EMPTY;
        for (uint32_t s = lane; s < 256u; s += 32u) M.tbl[s] = 0xFFFFu;
        if (lane < 4u) M.cnt[lane] = 0u;
        __syncwarp();
        if (lane == 0) {
            const uint32_t s = rb_hash(v) & (RB_CAP - 1u);
            M.key[s] = v;
            M.dist[s] = 0ull;
            M.cnt[0] = 1u;
        }
        __syncwarp();

        /* ---- distances (movementgraph.rs:64-115) ---- */
        for (;;) {
            bool changed = false;
            for (uint32_t s = lane; s < RB_CAP; s += 32u) {
                const uint32_t k = M.key[s];
                if (k == RB_EMPTY) continue;
                const unsigned long long d = M.dist[s];
                if (d >= M.done[s]) continue;
                const double dd = dsim_bits_to_f64(d);
                if (bound < dd) continue; /* beyond the range: given a distance, but does not relax (the search stops there) */
                M.done[s] = d;
                for (uint64_t e = g.edge_off[k]; e < g.edge_off[k + 1]; ++e) {
                    const uint32_t q = g.edge_dst[e];
                    const unsigned long long nd = dsim_f64_to_bits(dd + g.edge_sec[e]);
                    uint32_t slot = rb_hash(q) & (RB_CAP - 1u);
                    bool have = false;
                    for (;;) {
                        const uint32_t cur = *(volatile uint32_t *)&M.key[slot];
                        if (cur == q) {
                            have = true;
                            break;
                        }
                        if (cur == RB_EMPTY) {
                            if (*(volatile uint32_t *)&M.cnt[0] >= DSIM_REACH_DEV_MAX) { /* keeps the table from filling up */
                                M.cnt[2] = 1u;
                                break;
                            }
                            const uint32_t prev = atomicCAS(&M.key[slot], RB_EMPTY, q);
                            if (prev == RB_EMPTY) {
                                atomicAdd(&M.cnt[0], 1u);
                                have = true;
                                break;
                            }
                            if (prev == q) {
                                have = true;
                                break;
                            }
                        }
                        slot = (slot + 1u) & (RB_CAP - 1u);
                    }
                    if (have && nd < atomicMin(&M.dist[slot], nd)) changed = true;
                }
            }
            __syncwarp();
            if (*(volatile uint32_t *)&M.cnt[2] != 0u) break;
            if (!__any_sync(DSIM_FULL, changed)) break;
        }
        /* ---- 2-hop out-neighbourhood (lib.rs:1158-1178) ---- */
        for (uint64_t e1 = g.edge_off[v]; e1 < g.edge_off[v + 1]; ++e1) {
            const uint32_t nb = g.edge_dst[e1];
            for (uint64_t e2 = g.edge_off[nb] + lane; e2 < g.edge_off[nb + 1]; e2 += 32u) {
                const uint32_t q = g.edge_dst[e2];
                uint32_t slot = rb_hash(q) & (NB_CAP - 1u);
                for (;;) {
                    const uint32_t cur = *(volatile uint32_t *)&M.nset[slot];
                    if (cur == q) break;
                    if (cur == RB_EMPTY) {
                        if (*(volatile uint32_t *)&M.cnt[1] >= DSIM_NEAR_DEV_MAX) {
                            M.cnt[2] = 1u;
                            break;
                        }
                        const uint32_t prev = atomicCAS(&M.nset[slot], RB_EMPTY, q);
                        if (prev == RB_EMPTY) {
                            atomicAdd(&M.cnt[1], 1u);
                            break;
                        }
                        if (prev == q) break;
                    }
                    slot = (slot + 1u) & (NB_CAP - 1u);
                }
            }
        }
        __syncwarp();
        if (M.cnt[2] != 0u) { /* does not fit the per-warp tables: the caller builds the table on the host */
            if (lane == 0) atomicOr(&o.stats[0], 1u);
            __syncwarp();
            continue;
        }
        /* ---- compaction: entries with a normal distance (lib.rs:1147 drops the start node, d = 0) ---- */
        uint32_t m = 0, nn = 0;
        for (uint32_t s0 = 0; s0 < RB_CAP; s0 += 32u) {
            const uint32_t k = M.key[s0 + lane];
            const unsigned long long d = M.dist[s0 + lane];
            const uint32_t ex = (uint32_t)(d >> 52) & 0x7FFu;
            const bool ok = k != RB_EMPTY && ex != 0u && ex != 0x7FFu;
            const uint32_t mask = __ballot_sync(DSIM_FULL, ok);
            if (ok) {
                const uint32_t pos = m + (uint32_t)__popc(mask & lt);
                M.ckey[pos] = k;
                M.done[pos] = d;
            }
            m += (uint32_t)__popc(mask);
        }
        for (uint32_t s0 = 0; s0 < NB_CAP; s0 += 32u) {
            const uint32_t k = M.nset[s0 + lane];
            const uint32_t mask = __ballot_sync(DSIM_FULL, k != RB_EMPTY);
            if (k != RB_EMPTY) M.nkey[nn + (uint32_t)__popc(mask & lt)] = k;
            nn += (uint32_t)__popc(mask);
        }
        __syncwarp();
        /* sensed = also in the 2-hop set */
        uint32_t n_sensed = 0, span = 0;
        for (uint32_t i0 = 0; i0 < m; i0 += 32u) {
            const uint32_t i = i0 + lane;
            bool sensed = false;
            if (i < m) {
                const uint32_t q = M.ckey[i];
                uint32_t slot = rb_hash(q) & (NB_CAP - 1u);
                for (;;) {
                    const uint32_t cur = M.nset[slot];
                    if (cur == q) {
                        sensed = true;
                        break;
                    }
                    if (cur == RB_EMPTY) break;
                    slot = (slot + 1u) & (NB_CAP - 1u);
                }
                M.key[i] = sensed ? 1u : 0u;
                const uint32_t dq = q > v ? q - v : v - q;
                span = dq > span ? dq : span;
            }
            n_sensed += (uint32_t)__popc(__ballot_sync(DSIM_FULL, sensed));
        }
        for (uint32_t i = lane; i < nn; i += 32u) {
            const uint32_t q = M.nkey[i];
            const uint32_t dq = q > v ? q - v : v - q;
            span = dq > span ? dq : span;
        }
        __syncwarp();
        if (!FILL) {
            for (int d = 16; d > 0; d >>= 1) {
                const uint32_t other = __shfl_xor_sync(DSIM_FULL, span, d);
                span = other > span ? other : span;
            }
            if (lane == 0) {
                o.lens[4u * v + 0u] = m;
                o.lens[4u * v + 1u] = n_sensed;
                o.lens[4u * v + 2u] = nn;
                o.lens[4u * v + 3u] = 0u;
                atomicMax(&o.stats[1], span);
            }
            __syncwarp();
            continue;
        }
        /* ---- rows: [sensed, ascending][others, ascending][padding]; nearby list ascending ---- */
        const HostRowHdr hd = o.hdr[v];
        uint32_t *rowkey = reinterpret_cast<uint32_t *>(M.dist);
        for (uint32_t i = lane; i < m; i += 32u) {
            const uint32_t ki = M.ckey[i], si = M.key[i];
            uint32_t r = 0;
            for (uint32_t j = 0; j < m; ++j) r += (M.key[j] == si && M.ckey[j] < ki) ? 1u : 0u;
            const uint32_t pos = si ? r : n_sensed + r;
            o.dst[hd.start + pos] = ki;
            o.cost[hd.start + pos] = cost_per_second * dsim_bits_to_f64(M.done[i]);
            rowkey[pos] = ki;
        }
        for (uint32_t k = m + lane; k < hd.len; k += 32u) {
            o.dst[hd.start + k] = 0xFFFFFFFFu;
            o.cost[hd.start + k] = dsim_bits_to_f64(RB_INF);
        }
        const uint32_t n0 = o.near_off[v];
        for (uint32_t i = lane; i < nn; i += 32u) {
            const uint32_t ki = M.nkey[i];
            uint32_t r = 0;
            for (uint32_t j = 0; j < nn; ++j) r += M.nkey[j] < ki ? 1u : 0u;
            o.near_dst[n0 + r] = ki;
        }
        __syncwarp();
        /* ---- the row's bucket index, filled in row order like the host does (a full bucket overflows into the next) ---- */
        if (lane == 0 && hd.len <= DSIM_ROW_HASH_MAX) {
            for (uint32_t k = 0; k < m; ++k) {
                const uint32_t hsh = (rowkey[k] * 2654435761u) >> 19; /* = dsim_row_hash */
                uint32_t b = hsh >> 8;
                for (;;) {
                    uint32_t s = 0;
                    while (s < 8u && M.tbl[b * 8u + s] != 0xFFFFu) ++s;
                    if (s < 8u) {
                        M.tbl[b * 8u + s] = (uint16_t)(((hsh & 0xFFu) << 8) | k);
                        break;
                    }
                    b = (b + 1u) & 31u;
                }
            }
        }
        __syncwarp();
        for (uint32_t s = lane; s < 256u; s += 32u) o.hash[(size_t)v * 256u + s] = M.tbl[s];
        __syncwarp();
    }
}

double dsim_reach_cost_per_second(void) { return 1. / kSecondsPerYear * 3.; }

void dsim_reach_device_pass(int fill, const ReachDevGraph &g, const ReachDevOut &o, uint32_t n_sm, void *stream) {
    const uint32_t grid = n_sm * 6u; /* 37 KB of shared memory per block */
    cudaStream_t s = (cudaStream_t)stream;
    (void)s;
    if (fill)
        DSIM_LAUNCH(k_reach_build<true>, grid, RB_WARPS * 32, 0, s, g, o, kMaxSeasonRange, dsim_reach_cost_per_second());
    else
        DSIM_LAUNCH(k_reach_build<false>, grid, RB_WARPS * 32, 0, s, g, o, kMaxSeasonRange, dsim_reach_cost_per_second());
}
