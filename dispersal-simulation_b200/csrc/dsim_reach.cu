/*
 * dsim_reach.cu — host-side (OpenMP) construction of the reach table, once per graph.
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
