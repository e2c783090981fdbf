/*
 * oracle.cpp — CPU restatement of the step loop of Anaphory/dispersal-simulation
 * (supplement/dispersal_model_rust/src/lib.rs; all `lib.rs:` citations below refer to that file,
 * other files are cited as src/<name>.rs).
 *
 * TEST INFRASTRUCTURE ONLY — see oracle.h.  PARITY UNPINNED by the reference (no tests, no
 * fixtures, does not compile); pinned by the known-answer vectors in tests/.
 *
 * The code keeps the reference's shape on purpose: array-of-structs families with a dense
 * 304-entry adaptation vector (src/ecology.rs:140-161), an unbounded history (lib.rs:218), one
 * bounded Dijkstra per family per decision (lib.rs:843), per-patch vectors of bands.  The
 * canonicalisations that make it deterministic (the reference is not: unseeded thread-local
 * fastrand, randomised hash iteration orders) are listed in DESIGN.md §3 and marked CANON below.
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fopenmp  (no FMA contraction: decisions depend on
 * double-precision comparisons that the GPU path has to reproduce bit for bit).
 */
#include "oracle.h"

#if defined(_OPENMP)
#include <omp.h>
#endif
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <queue>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {

/* ------------------------------------------------------------------------------------------ *
 * Random stream.  The reference draws from an unseeded thread-local `fastrand` generator under
 * rayon (lib.rs:859,1271,1278-1279,1296,1310,1318,1326,1334) and is therefore irreproducible.
 * CANON: every draw is a pure function Philox4x32-10(key = seed, counter = (step, entity id,
 * purpose tag, sub index)).  Restated from Salmon et al., "Parallel random numbers: as easy as
 * 1, 2, 3" (SC'11); pinned by the Random123 known-answer vectors in tests/test_oracle_kat.py.
 * ------------------------------------------------------------------------------------------ */
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum Purpose : uint32_t {
    P_MEM = 1,     /* fastrand::f64() of lib.rs:859, history entry n           entity = family id */
    P_NOISE = 2,   /* resource_uncertainty() lib.rs:1317, candidate i           entity = family id */
    P_SHUF = 3,    /* shuffle lib.rs:1333                                       entity = family id */
    P_FIGHT = 4,   /* fight_is_deadly lib.rs:1325 (both calls of one encounter) entity = newcomer  */
    P_PIVOT = 5,   /* select_pivot_family lib.rs:1295, band b                   entity = node id   */
    P_MUT = 6,     /* time_till_next_mutation / change_random_bitvec_component  entity = family id */
    P_REGROW = 7,  /* recover_resources_proportion lib.rs:1309, ecoregion slot  entity = node id   */
    P_MEM2 = 8,    /* low 32 bits of the memory draw of history entry n         entity = family id */
    P_MEMT = 9     /* skip sampling over the tail of the history (stream v2)    entity = family id */
};

struct Stream {
    uint32_t key[2];
    void draw(uint64_t step, uint64_t entity, Purpose tag, uint32_t sub, uint32_t out[4]) const {
        const uint32_t ctr[4] = {(uint32_t)step, (uint32_t)entity, (uint32_t)(entity >> 32),
                                 ((uint32_t)tag << 24) | (sub & 0xFFFFFFu)};
        philox4x32_10(ctr, key, out);
    }
};

/* uniform on [0,1) with 53 random bits */
inline double u53(uint32_t lo, uint32_t hi) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

/* ---- the canonical draws, one function per call site of the reference's `stochasticity` module (lib.rs:1260-1336);
 * the step below and the statistical tests (oracle_draw_* exports, tests/test_oracle_stats.py) both go through these */
inline double draw_noise(const Stream &rng, uint64_t t, uint64_t fid, uint32_t noise_index) {
    /* resource_uncertainty, lib.rs:1317.  Two draws per Philox block: indices i and i + 32 of every run of 64 share a
     * block, block = (i / 64) * 32 + i % 32, words (0,1) if i % 64 < 32, else (2,3). */
    uint32_t out[4];
    rng.draw(t, fid, P_NOISE, ((noise_index >> 6) << 5) | (noise_index & 31u), out);
    return (noise_index & 32u) ? u53(out[2], out[3]) : u53(out[0], out[1]);
}
inline uint64_t draw_shuffle_key(const Stream &rng, uint64_t t, uint64_t fid) { /* shuffle, lib.rs:1333 */
    uint32_t out[4];
    rng.draw(t, fid, P_SHUF, 0, out);
    return ((uint64_t)out[1] << 32) | out[0];
}
inline void draw_fight(const Stream &rng, uint64_t t, uint64_t newcomer, uint32_t opponent_position, uint32_t deadliness,
                       bool *opponent_hit, bool *newcomer_hit) { /* fight_is_deadly x2, lib.rs:1219,1226,1325 */
    uint32_t out[4];
    rng.draw(t, newcomer, P_FIGHT, opponent_position, out);
    *opponent_hit = out[0] < deadliness;
    *newcomer_hit = out[1] < deadliness;
}
inline uint64_t draw_pivot(const Stream &rng, uint64_t t, uint64_t node, uint32_t band, uint64_t len) {
    /* select_pivot_family, lib.rs:1295-1298: fastrand::usize(0..len) */
    uint32_t out[4];
    rng.draw(t, node, P_PIVOT, band, out);
    const uint64_t r = ((uint64_t)out[1] << 32) | out[0];
    return (uint64_t)(((unsigned __int128)r * (unsigned __int128)len) >> 64);
}
inline double draw_regrowth(const Stream &rng, uint64_t t, uint64_t node, uint32_t eco_slot) {
    /* recover_resources_proportion, lib.rs:1309-1311: 2 * f64() */
    uint32_t out[4];
    rng.draw(t, node, P_REGROW, eco_slot, out);
    return 2.0 * u53(out[0], out[1]);
}

/* CANON: natural logarithm evaluated with +,-,*,/ only, in a fixed order, so that host and device
 * agree bit for bit (lib.rs:1271 truncates a quotient of logarithms to an integer).
 * x = 2^e * m, m in (sqrt(1/2), sqrt(2)];  ln m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716;
 * atanh series to s^27 (next term < 2^-70). */
double det_log(double x) {
    if (x != x) return x;
    if (x < 0.0) return std::numeric_limits<double>::quiet_NaN();
    if (x == 0.0) return -std::numeric_limits<double>::infinity();
    if (x == std::numeric_limits<double>::infinity()) return x;
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    int e = 0;
    if ((bits >> 52) == 0) { /* subnormal */
        x *= 18014398509481984.0; /* 2^54 */
        std::memcpy(&bits, &x, 8);
        e = -54;
    }
    e += (int)(bits >> 52) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m;
    std::memcpy(&m, &bits, 8);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double p = 1.0 / 27.0;
    p = p * z + 1.0 / 25.0;
    p = p * z + 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z + 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z + 1.0;
    const double r = (2.0 * s) * p;
    const double de = (double)e;
    return de * 6.93147180369123816490e-01 + (de * 1.90821492927058770002e-10 + r);
}

/* Rust's `as u32` on an f64: truncate toward zero, saturate, NaN -> 0 (lib.rs:535,1271,1878,1916). */
uint32_t f64_as_u32(double x) {
    if (x != x) return 0;
    if (x <= 0.0) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}

/* OneYearResources ordering (src/ecology.rs:13-34): partial_cmp with "incomparable = Equal". */
inline double oyr_min(double a, double b) { return (a > b) ? b : a; } /* std::cmp::min(a, b)  */
inline double oyr_max(double a, double b) { return (a > b) ? a : b; } /* Ord::max(a, b)       */

/* ------------------------------------------------------------------------------------------ *
 * Types (lib.rs:142-152, 178-229, 253-258; src/parameters.rs:5-20; src/movementgraph.rs:8-9)
 * ------------------------------------------------------------------------------------------ */
constexpr int N_ECO = DSIM_N_ECOREGIONS; /* src/ecology.rs:137,143 */

struct Family {
    uint64_t id = 0;            /* descendence (lib.rs:227) as a number: birth order */
    uint64_t parent = DSIM_NO_PARENT;
    uint16_t birth_index = 0;   /* the ":<k>" of lib.rs:1880 */
    uint64_t effective_size = 0;
    uint64_t culture = 0;       /* Culture.in_memory[0], lib.rs:257 */
    uint32_t location = 0;
    double stored_resources = 0.0;
    uint16_t number_offspring = 0;
    std::vector<std::pair<uint32_t, double>> history;
    bool has_mutation_clock = false; /* Option<Seasons>, lib.rs:220 */
    uint32_t seasons_till_next_mutation = 0;
    uint32_t seasons_till_next_adult = 0;
    std::array<double, N_ECO> adaptation; /* Ecovector.entries, src/ecology.rs:143 */
    double last_harvest = 0.0;
};

struct Graph {
    uint32_t n = 0;
    std::vector<uint64_t> h3;
    std::vector<double> lon, lat;
    std::vector<uint64_t> eco_off;
    std::vector<uint32_t> eco_id;
    std::vector<double> popcap;
    std::vector<uint64_t> edge_off;
    std::vector<uint32_t> edge_dst;
    std::vector<double> edge_sec;
};

typedef std::vector<std::pair<uint64_t, uint64_t>> CultureMap; /* FxHashMap<Culture,(usize,OYR)>: key -> size */

void culture_map_insert(CultureMap &m, uint64_t culture, uint64_t size) {
    for (auto &kv : m)
        if (kv.first == culture) {
            kv.second = size; /* HashMap::insert overwrites */
            return;
        }
    m.emplace_back(culture, size);
}

const double SECONDS_PER_YEAR = 365.24219 * 24. * 60. * 60.;                 /* lib.rs:100  */
const double MAX_SEASON_RANGE = 320. / 6. * 1000. / 0.7740708353271509;      /* lib.rs:1183 */
const double YEARS_BETWEEN_ADULTS = (2.85 - 1.35) / (1. - 0.488);            /* lib.rs:1910 */

} // namespace

struct oracle_handle {
    dsim_params p;
    Graph g;
    Stream rng;
    uint64_t t = 0;
    uint64_t next_id = 0;
    int mode = 0;
    std::vector<Family> families;
    std::vector<double> res, max_res;                /* Patch.resources, aligned with g.eco_off */
    std::vector<CultureMap> storage;                 /* per node; empty = no entry */
    std::vector<std::unique_ptr<std::vector<uint32_t>>> nearby_cache;            /* lib.rs:487 */
    std::vector<std::unique_ptr<std::vector<std::pair<uint32_t, double>>>> reach_cache; /* mode 1 */
    std::vector<double> mem_table; /* memory after n decays, lib.rs:847-866 */
    uint32_t mem_tail0 = 0;        /* first history entry sampled by skipping (CANON, see remembered_entries) */
    std::vector<double> mem_tail;  /* mem_tail[j] = prod_{k <= j} (1 - memory_{mem_tail0 + k}) */
    std::string err;
    uint32_t model_errors = 0;
    uint64_t births = 0, deaths = 0, updates = 0, occupied = 0;
};

namespace {

std::string g_create_error;

/* ------------------------------------------------------------------------------------------ *
 * 4.7 Sensing (lib.rs:1110-1184, src/movementgraph.rs:64-115)
 * ------------------------------------------------------------------------------------------ */

/* src/movementgraph.rs:64-115.  Returns every node that was ever given a score: nodes settled at a
 * distance <= max_distance AND the nodes relaxed from them (tentative distances). */
std::unordered_map<uint32_t, double> bounded_dijkstra(const Graph &g, uint32_t start, double max_distance) {
    std::unordered_set<uint32_t> visited;
    std::unordered_map<uint32_t, double> scores;
    typedef std::pair<double, uint32_t> Scored;
    auto cmp = [](const Scored &a, const Scored &b) { return a.first > b.first; }; /* MinScored :23-62 */
    std::priority_queue<Scored, std::vector<Scored>, decltype(cmp)> visit_next(cmp);
    scores[start] = 0.0;
    visit_next.push({0.0, start});
    while (!visit_next.empty()) {
        const Scored top = visit_next.top();
        visit_next.pop();
        const double node_score = top.first;
        const uint32_t node = top.second;
        if (visited.count(node)) continue;
        if (max_distance < node_score) break;
        for (uint64_t e = g.edge_off[node]; e < g.edge_off[node + 1]; ++e) {
            const uint32_t next = g.edge_dst[e];
            if (visited.count(next)) continue;
            double next_score = node_score + g.edge_sec[e];
            auto it = scores.find(next);
            if (it != scores.end()) {
                if (next_score < it->second)
                    it->second = next_score;
                else
                    next_score = it->second;
            } else {
                scores.emplace(next, next_score);
            }
            visit_next.push({next_score, next});
        }
        visited.insert(node);
    }
    return scores;
}

/* lib.rs:1133-1154.  CANON: returned sorted by node id (the reference returns a HashMap; only
 * look-ups are performed on it, so the order is immaterial). */
std::vector<std::pair<uint32_t, double>> many_nearby_locations(const Graph &g, uint32_t location) {
    const double cost_per_walking_second = 1. / SECONDS_PER_YEAR * 3.; /* lib.rs:1139-1140 */
    std::vector<std::pair<uint32_t, double>> out;
    for (const auto &nd : bounded_dijkstra(g, location, MAX_SEASON_RANGE))
        if (std::isnormal(nd.second)) /* drops the start node (d = 0), lib.rs:1147 */
            out.emplace_back(nd.first, cost_per_walking_second * nd.second);
    std::sort(out.begin(), out.end());
    return out;
}

/* lib.rs:1158-1178.  CANON: ascending node id (the reference drains a HashSet in arbitrary order,
 * which feeds the order-dependent best_location and the `nearby[1..]` of lib.rs:1846). */
std::vector<uint32_t> nearby_locations(const Graph &g, uint32_t location) {
    std::vector<uint32_t> out;
    for (uint64_t e = g.edge_off[location]; e < g.edge_off[location + 1]; ++e) {
        const uint32_t neighbor = g.edge_dst[e];
        for (uint64_t ne = g.edge_off[neighbor]; ne < g.edge_off[neighbor + 1]; ++ne)
            out.push_back(g.edge_dst[ne]);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

const std::vector<uint32_t> &cached_nearby(oracle_handle &h, uint32_t loc) {
    /* nearby_cache.entry(loc).or_insert_with(...), lib.rs:1816-1818 */
    std::vector<uint32_t> *p;
#pragma omp critical(oracle_nearby)
    {
        if (!h.nearby_cache[loc]) h.nearby_cache[loc].reset(new std::vector<uint32_t>(nearby_locations(h.g, loc)));
        p = h.nearby_cache[loc].get();
    }
    return *p;
}

/* ------------------------------------------------------------------------------------------ *
 * 4.4 Objectives (lib.rs:989-1062)
 * ------------------------------------------------------------------------------------------ */
bool will_cooperate(uint64_t a, uint64_t b, const dsim_params &p) {
    /* emergence::will_cooperate is missing from the reference; see DESIGN.md §3 (a15). */
    const uint32_t d = (uint32_t)__builtin_popcountll(a ^ b);
    return p.cooperation_inclusive ? (d <= p.cooperation_threshold) : (d < p.cooperation_threshold);
}

double expected_harvest(const oracle_handle &h, uint32_t node) { /* lib.rs:989-994 */
    double sum = 0.0; /* CANON: ascending ecoregion id (the reference iterates an FxHashMap) */
    for (uint64_t k = h.g.eco_off[node]; k < h.g.eco_off[node + 1]; ++k)
        sum += h.res[k] + h.max_res[k] * h.p.resource_recovery_per_season;
    return sum;
}

bool has_enemies(const oracle_handle &h, uint32_t node, uint64_t culture) { /* lib.rs:996-1017 */
    for (const auto &kv : h.storage[node])
        if (!will_cooperate(culture, kv.first, h.p)) return true;
    return false;
}

double destination_quality(const oracle_handle &h, uint32_t i, const Family &family, uint64_t population) {
    /* lib.rs:1034-1062 */
    const uint64_t pop_at_destination = population + (family.location == i ? 0 : family.effective_size);
    if (has_enemies(h, i, family.culture)) return 0.0;
    return expected_harvest(h, i) / (double)pop_at_destination;
}

/* ------------------------------------------------------------------------------------------ *
 * 4.3 Adaptation: decide_on_moving + best_location (lib.rs:830-929)
 * ------------------------------------------------------------------------------------------ */
struct Decision {
    uint32_t target;
    double cost;
};

/* Which of the `len` newest history entries (0 = newest) a family remembers this step: the reference keeps entry n
 * iff `fastrand::f64() < memory_n`, memory_n = 0.5^(n * season / 2) by repeated multiplication, one independent draw
 * per entry while memory_n >= f64::EPSILON (lib.rs:847-866).  CANON (stream v2), same distribution:
 *  - head, n < tail0: a 40-bit uniform U_n = (B_n * 2^32 + W_n) * 2^-40 per entry, B_n = byte n&15 of Philox block
 *    n>>4 of the MEM stream (16 entries per block), W_n = word n&3 of block n>>2 of the MEM2 stream; kept iff
 *    U_n < memory_n.  (The device draws W_n only when B_n alone does not decide.)
 *  - tail, n >= tail0 (every memory_n < 2^-8: 0.07 entries expected out of 500): the independent draws are replaced
 *    by exact skip sampling.  With S_j = prod_{k <= j} (1 - memory_{tail0 + k}) (`tail`, the probability that none of
 *    the first j+1 tail entries is kept), given that the entries before j0 are decided and base = S_{j0-1} (1 at the
 *    start), the next kept entry is the smallest j >= j0 with S_j <= U * base for one uniform U — P(J > j) =
 *    S_j / base, exactly the law of independent draws; then j0 = J + 1, base = S_J and the next uniform decides the
 *    entry after that.  Uniform d of the family and step is u53 of words (0,1) (d even) or (2,3) (d odd) of block
 *    d>>1 of the MEMT stream.
 * Output: kept entry indices, ascending (the order the reference considers them in). */
void remembered_entries(const Stream &rng, uint64_t t, uint64_t fid, uint32_t len, const std::vector<double> &mem,
                        uint32_t tail0, const std::vector<double> &tail, std::vector<uint32_t> &kept) {
    uint32_t out[4];
    kept.clear();
    const uint32_t n_head = std::min(len, tail0);
    for (uint32_t n = 0; n < n_head; ++n) {
        rng.draw(t, fid, P_MEM, n >> 4, out);
        const uint64_t b = (out[(n >> 2) & 3] >> (8 * (n & 3))) & 0xFFu;
        rng.draw(t, fid, P_MEM2, n >> 2, out);
        const uint64_t x = (b << 32) | out[n & 3];
        const double u = (double)x * (1.0 / 1099511627776.0);
        if (u < mem[n]) kept.push_back(n);
    }
    if (len <= tail0) return;
    const uint32_t L = len - tail0; /* tail entries this family has */
    uint32_t j0 = 0, d = 0;
    double base = 1.0;
    while (j0 < L) {
        rng.draw(t, fid, P_MEMT, d >> 1, out);
        const double u = (d & 1u) ? u53(out[2], out[3]) : u53(out[0], out[1]);
        const double thr = u * base;
        uint32_t j = j0;
        while (j < L && tail[j] > thr) ++j;
        if (j >= L) break;
        kept.push_back(tail0 + j);
        base = tail[j];
        j0 = j + 1;
        d += 1;
    }
}

Decision decide_on_moving(oracle_handle &h, const Family &family, const uint32_t *known, size_t n_known,
                          const std::vector<uint64_t> &population) {
    /* lib.rs:843: one bounded Dijkstra per call */
    std::vector<std::pair<uint32_t, double>> local;
    const std::vector<std::pair<uint32_t, double>> *travel_costs;
    if (h.mode == 0) {
        local = many_nearby_locations(h.g, family.location);
        travel_costs = &local;
    } else {
        std::vector<std::pair<uint32_t, double>> *p;
#pragma omp critical(oracle_reach)
        {
            if (!h.reach_cache[family.location])
                h.reach_cache[family.location].reset(
                    new std::vector<std::pair<uint32_t, double>>(many_nearby_locations(h.g, family.location)));
            p = h.reach_cache[family.location].get();
        }
        travel_costs = p;
    }
    auto cost_of = [&](uint32_t q, double *c) {
        auto it = std::lower_bound(travel_costs->begin(), travel_costs->end(), std::make_pair(q, -1.0));
        if (it == travel_costs->end() || it->first != q) return false;
        *c = it->second;
        return true;
    };

    /* best_location state, lib.rs:911-916 */
    uint32_t target = family.location;
    double max_gain = 0.0, t_cost = 0.0;
    double m = max_gain - t_cost;

    auto consider = [&](uint32_t q, double gain, double l_cost, uint32_t noise_index) {
        /* lib.rs:918-927.  CANON noise index: the reference draws once per candidate that has a travel
         * cost, in iteration order; a sensed patch uses its ordinal among the sensed patches that have a
         * travel cost, the n-th newest history entry uses n_known + n (whether or not earlier entries were
         * kept); two draws per Philox block: indices i and i + 32 of every run of 64 share a block,
         * block = (i / 64) * 32 + i % 32, words (0,1) if i % 64 < 32, else (2,3). */
        const double u = draw_noise(h.rng, h.t, family.id, noise_index);
        const double g = gain * u - l_cost;
        if (g > m) {
            target = q;
            max_gain = oyr_max(max_gain, gain);
            t_cost = l_cost;
            m = max_gain - t_cost;
        }
    };

    /* known_destinations.iter().map(|d| (d, None)), lib.rs:851-853, filtered by lib.rs:869 */
    uint32_t n_considered = 0;
    for (size_t i = 0; i < n_known; ++i) {
        double c;
        if (!cost_of(known[i], &c)) continue;
        const double harvest = destination_quality(h, known[i], family, population[known[i]]);
        consider(known[i], harvest, c, n_considered++);
    }
    /* family.history.iter().rev().filter_map(...), lib.rs:854-867 */
    std::vector<uint32_t> kept;
    remembered_entries(h.rng, h.t, family.id, (uint32_t)std::min<size_t>(family.history.size(), h.mem_table.size()), h.mem_table,
                       h.mem_tail0, h.mem_tail, kept);
    for (const uint32_t n : kept) {
        const auto &entry = family.history[family.history.size() - 1 - n];
        double c;
        if (!cost_of(entry.first, &c)) continue;
        consider(entry.first, entry.second, c, (uint32_t)(n_known + n));
    }
    return Decision{target, t_cost};
}

/* ------------------------------------------------------------------------------------------ *
 * 7.x family_lifecycle (lib.rs:1788-1942)
 * ------------------------------------------------------------------------------------------ */
double resources_at_season_end(double resources, uint64_t size, double season) { /* lib.rs:1861-1867 */
    return resources - (double)size * season;
}

bool use_resources_and_maybe_shrink(uint64_t &size, double &resources, double season) { /* lib.rs:1923-1937 */
    bool has_shrunk = false;
    while (resources_at_season_end(resources, size, season) < 0. && size > 0) {
        size -= 1;
        has_shrunk = true;
    }
    resources = resources_at_season_end(resources, size, season);
    return has_shrunk;
}

void maybe_grow(Family &family, double season) { /* lib.rs:1912-1920 */
    if (family.seasons_till_next_adult == 0) {
        family.effective_size += 1;
        family.seasons_till_next_adult = f64_as_u32(YEARS_BETWEEN_ADULTS / season);
    } else {
        family.seasons_till_next_adult -= 1;
    }
}

bool can_survive(const Family &family) { return family.effective_size >= 2; } /* lib.rs:1939-1941 */

bool maybe_procreate(Family &family, double season, uint64_t child_id, Family *child) { /* lib.rs:1869-1897 */
    if (family.effective_size < 10) return false;
    family.number_offspring += 1;
    const double resources_they_take = family.stored_resources * 2.0 / (double)family.effective_size;
    family.stored_resources -= resources_they_take;
    family.effective_size -= 2;
    const uint32_t time_to_grow_adult = f64_as_u32(15. / season) + 1;
    child->id = child_id;
    child->parent = family.id;
    child->birth_index = family.number_offspring;
    child->location = family.location;
    const size_t len = family.history.size();
    const size_t keep = std::min<size_t>(len, time_to_grow_adult);
    child->history.assign(family.history.begin() + (len - keep), family.history.end());
    child->seasons_till_next_adult = time_to_grow_adult;
    child->culture = family.culture;
    child->effective_size = 2;
    child->number_offspring = 0;
    child->has_mutation_clock = false;
    child->seasons_till_next_mutation = 0;
    child->stored_resources = resources_they_take;
    child->adaptation = family.adaptation;
    child->last_harvest = 0.0;
    return true;
}

/* ------------------------------------------------------------------------------------------ *
 * Part 1 of the step: update_each_family_step (lib.rs:511-552)
 * ------------------------------------------------------------------------------------------ */
void update_each_family_step(oracle_handle &h) {
    const double season = h.p.season_length_in_years;
    std::vector<Family> &families = h.families;
    h.updates += families.size();

    /* census at step start, lib.rs:522-525 */
    std::vector<uint64_t> cc(h.g.n, 0);
    for (const Family &f : families) cc[f.location] += f.effective_size;

    /* lib.rs:527-539 */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)families.size(); ++i) {
        Family &family = families[i];
        if (use_resources_and_maybe_shrink(family.effective_size, family.stored_resources, season))
            family.seasons_till_next_adult = f64_as_u32(15. / season) + 1; /* lib.rs:535 */
        maybe_grow(family, season);
    }

    /* families.retain(can_survive), lib.rs:541 */
    const size_t before = families.size();
    families.erase(std::remove_if(families.begin(), families.end(), [](const Family &f) { return !can_survive(f); }),
                   families.end());
    h.deaths += before - families.size();

    /* procreate_and_migrate, lib.rs:1795-1859.  CANON: children get consecutive ids in parent order,
     * which is the order `collect()` + `append` (lib.rs:551,1858) put them into the vector. */
    std::vector<long> child_slot(families.size(), -1);
    long n_children = 0;
    for (size_t i = 0; i < families.size(); ++i)
        if (families[i].effective_size >= 10) child_slot[i] = n_children++;
    std::vector<Family> children((size_t)n_children);
    const uint64_t first_child_id = h.next_id;
    h.next_id += (uint64_t)n_children;

#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)families.size(); ++i) {
        Family &family = families[i];
        const std::vector<uint32_t> &nearby = cached_nearby(h, family.location);
        const Decision d = decide_on_moving(h, family, nearby.data(), nearby.size(), cc);
        /* lib.rs:1826-1833 */
        family.history.emplace_back(family.location, family.last_harvest / (double)family.effective_size);
        family.stored_resources += family.last_harvest;
        family.last_harvest = 0.0;
        family.location = d.target;
        family.stored_resources -= d.cost;
        if (child_slot[i] >= 0) {
            Family &descendant = children[(size_t)child_slot[i]];
            maybe_procreate(family, season, first_child_id + (uint64_t)child_slot[i], &descendant);
            /* lib.rs:1842-1853: `&nearby[1..]` of the parent's OLD patch; CANON: drops the smallest id */
            const size_t skip = nearby.empty() ? 0 : 1;
            const Decision d2 = decide_on_moving(h, descendant, nearby.data() + skip, nearby.size() - skip, cc);
            descendant.location = d2.target;
            descendant.stored_resources -= d2.cost;
        }
    }
    h.births += (uint64_t)n_children;
    for (Family &c : children) families.push_back(std::move(c)); /* lib.rs:551 */
}

/* ------------------------------------------------------------------------------------------ *
 * 4.8/4.10 Interaction and collectives (lib.rs:1196-1246, 1351-1418)
 * ------------------------------------------------------------------------------------------ */

/* `Band` is missing from the reference; inferred from lib.rs:1390-1393 as {culture, families}.
 * Each member's position in the shuffled list is carried alongside so that the fight draws can be
 * keyed by (newcomer id, opponent position). */
struct Member {
    Family *family;
    uint32_t position;
};
struct PBand {
    uint64_t culture;
    std::vector<Member> families;
};

bool encounter_maybe_join(oracle_handle &h, Family *family, PBand &group) {
    bool join = true;
    for (Member &other : group.families) {
        Family *other_family = other.family;
        if (!will_cooperate(family->culture, other_family->culture, h.p)) {
            join = false;
            const uint64_t reduction = std::min(family->effective_size, other_family->effective_size);
            bool other_hit, self_hit;
            draw_fight(h.rng, h.t, family->id, other.position, h.p.fight_deadliness, &other_hit, &self_hit);
            if (other_hit) {
                other_family->effective_size -= reduction;
                if (other_family->effective_size == 0) {
                    family->stored_resources += other_family->stored_resources;
                    other_family->stored_resources = 0.0;
                }
            }
            if (self_hit) family->effective_size -= reduction;
        }
    }
    return join;
}

/* lib.rs:1372-1417 */
std::vector<PBand> cooperate_or_fight(oracle_handle &h, uint32_t node, std::vector<Family *> families_in_this_location) {
    /* stochasticity::shuffle, lib.rs:1376.  CANON: the uniformly random permutation is realised by
     * ordering the families by a keyed 64-bit random number (ties by id); unlike Fisher–Yates this
     * does not depend on the order the families were pushed (lib.rs:607-612 is a parallel push). */
    std::vector<std::pair<std::pair<uint64_t, uint64_t>, Family *>> keyed;
    keyed.reserve(families_in_this_location.size());
    for (Family *f : families_in_this_location) keyed.push_back({{draw_shuffle_key(h.rng, h.t, f->id), f->id}, f});
    std::sort(keyed.begin(), keyed.end(),
              [](const auto &a, const auto &b) { return a.first < b.first; });

    std::vector<PBand> bands;
    for (uint32_t pos = 0; pos < keyed.size(); ++pos) {
        Family *family = keyed[pos].second;
        PBand *joined_band = nullptr;
        for (PBand &band : bands) {
            if (encounter_maybe_join(h, family, band)) {
                joined_band = &band;
                break;
            }
        }
        if (joined_band == nullptr)
            bands.push_back(PBand{family->culture, {Member{family, pos}}});
        else
            joined_band->families.push_back(Member{family, pos});
    }
    std::vector<PBand> alive;
    for (uint32_t b = 0; b < bands.size(); ++b) {
        PBand &band = bands[b];
        /* select_pivot_family, lib.rs:1295-1298: fastrand::usize(0..len) */
        const uint64_t index = draw_pivot(h.rng, h.t, node, b, band.families.size());
        band.culture = band.families[index].family->culture;
        uint64_t bandsize = 0;
        for (Member &m : band.families) bandsize += m.family->effective_size;
        if (bandsize > 0) alive.push_back(std::move(band));
    }
    return alive;
}

/* ------------------------------------------------------------------------------------------ *
 * 7.2 Culture (lib.rs:1754-1785), stochasticity (lib.rs:1270-1285)
 * ------------------------------------------------------------------------------------------ */
uint32_t time_till_next_mutation(double u, double culture_mutation_rate) { /* lib.rs:1270-1272 */
    /* f64::log(self, base) = ln(self) / ln(base) */
    return f64_as_u32(det_log(u) / det_log(1. - culture_mutation_rate));
}

void mutate(oracle_handle &h, Family &f, uint64_t target_culture) {
    const double rate = h.p.culture_mutation_rate;
    uint32_t a[4], b[4];
    h.rng.draw(h.t, f.id, P_MUT, 0, a);
    f.culture = target_culture; /* lib.rs:1761 */
    if (!f.has_mutation_clock) { /* None => draw, then fall through, lib.rs:1763-1774 */
        f.has_mutation_clock = true;
        f.seasons_till_next_mutation = time_till_next_mutation(u53(a[0], a[1]), rate);
    }
    if (f.seasons_till_next_mutation == 0) { /* Some(0), lib.rs:1775-1780 */
        /* change_random_bitvec_component, lib.rs:1277-1285: component 0 of 1, bit u32(0..64) */
        f.culture ^= (uint64_t)1 << (a[2] >> 26);
        h.rng.draw(h.t, f.id, P_MUT, 1, b);
        f.seasons_till_next_mutation = time_till_next_mutation(u53(b[0], b[1]), rate);
    } else { /* Some(k), lib.rs:1781-1783 */
        f.seasons_till_next_mutation -= 1;
    }
}

void adjust_culture(oracle_handle &h, PBand &group) { /* lib.rs:1960-1970 */
    for (Member &m : group.families) mutate(h, *m.family, group.culture);
}

/* ------------------------------------------------------------------------------------------ *
 * 7.x ecology (lib.rs:1944-2110, src/ecology.rs:147-161)
 * ------------------------------------------------------------------------------------------ */
double group_size_effect(double raw_group_size) { return raw_group_size * raw_group_size; } /* powi(2), lib.rs:1243 */

double adapt_entry_decay(double v, double minimum) { return std::fmax(minimum, v * 0.95); } /* ecology.rs:153 */

double raw_family_contribution(Family &family, uint32_t ecoregion, double minimum, double *sum_effort) {
    /* lib.rs:1947-1958 */
    for (int k = 0; k < N_ECO; ++k) family.adaptation[k] = adapt_entry_decay(family.adaptation[k], minimum);
    family.adaptation[ecoregion] += 0.05;
    const double contribution = (double)family.effective_size * family.adaptation[ecoregion];
    *sum_effort += contribution;
    return contribution;
}

double recover(double patch_resources, double patch_max_resources, double rate, double proportion) {
    /* lib.rs:2095-2109; proportion = 2 * U, lib.rs:1309-1311 */
    if (patch_resources < patch_max_resources)
        patch_resources = oyr_min(patch_max_resources, patch_max_resources * rate * proportion + patch_resources);
    return patch_resources;
}

/* the per-band part of exploit_patch for one ecoregion, lib.rs:2026-2064 */
double exploit_ecoregion(oracle_handle *h, std::vector<PBand> &groups, uint32_t ecoregion, double res, double cap,
                         double season, double minimum, uint32_t *errors) {
    (void)h;
    double total_squared_groups_effort = 1.0;
    std::vector<std::pair<double, std::vector<double>>> work;
    for (PBand &group : groups) {
        double sum_effort = 0.0;
        std::vector<double> contributions;
        for (Member &m : group.families)
            contributions.push_back(raw_family_contribution(*m.family, ecoregion, minimum, &sum_effort));
        total_squared_groups_effort += group_size_effect(sum_effort);
        work.emplace_back(sum_effort, std::move(contributions));
    }
    for (size_t b = 0; b < groups.size(); ++b) {
        const double group_effort = work[b].first;
        if (!(group_effort > 0.)) *errors |= DSIM_MODEL_EFFORT_NOT_POSITIVE;
        const double group_harvest =
            oyr_min(res * group_size_effect(group_effort) / total_squared_groups_effort, cap * season * group_effort);
        if (!(group_harvest > 0.)) *errors |= DSIM_MODEL_HARVEST_NOT_POSITIVE;
        const double coefficient = oyr_min(group_harvest / group_effort, 2.5);
        if (!(coefficient > 0.)) *errors |= DSIM_MODEL_COEFF_NOT_POSITIVE;
        for (size_t k = 0; k < groups[b].families.size(); ++k)
            groups[b].families[k].family->last_harvest = coefficient * season * work[b].second[k];
        const double effort = coefficient * group_effort;
        res -= effort;
        if (!(res >= 0.)) *errors |= DSIM_MODEL_RESOURCES_NEGATIVE;
    }
    return res;
}

/* ------------------------------------------------------------------------------------------ *
 * Part 2 of the step: distribute_each_patch_resources_step (lib.rs:591-655)
 * ------------------------------------------------------------------------------------------ */
void distribute_each_patch_resources_step(oracle_handle &h) {
    std::vector<Family> &families = h.families;
    /* families_by_location, lib.rs:602-612 */
    std::vector<uint32_t> order(families.size());
    for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return families[a].location < families[b].location; });
    std::vector<std::pair<size_t, size_t>> segments;
    for (size_t i = 0; i < order.size();) {
        size_t j = i;
        while (j < order.size() && families[order[j]].location == families[order[i]].location) ++j;
        segments.emplace_back(i, j);
        i = j;
    }
    h.occupied = segments.size();

    std::vector<CultureMap> stored_resources(h.g.n); /* lib.rs:614 */
    uint32_t errors = 0;

#pragma omp parallel for schedule(dynamic, 16) reduction(| : errors)
    for (long s = 0; s < (long)segments.size(); ++s) {
        const uint32_t patch_id = families[order[segments[s].first]].location;
        std::vector<Family *> here;
        for (size_t k = segments[s].first; k < segments[s].second; ++k) here.push_back(&families[order[k]]);

        std::vector<PBand> groups = cooperate_or_fight(h, patch_id, std::move(here)); /* lib.rs:620 */
        CultureMap &cs = stored_resources[patch_id];
        for (PBand &group : groups) { /* lib.rs:625-633 */
            adjust_culture(h, group);
            uint64_t group_size = 0;
            for (Member &m : group.families) group_size += m.family->effective_size;
            culture_map_insert(cs, group.culture, group_size);
        }
        /* exploit_patch, lib.rs:2021-2067.  CANON: ecoregions in ascending id (FxHashMap order in the
         * reference decides whose `last_harvest` survives, lib.rs:2058). */
        for (uint64_t k = h.g.eco_off[patch_id]; k < h.g.eco_off[patch_id + 1]; ++k) {
            double res = exploit_ecoregion(&h, groups, h.g.eco_id[k], h.res[k],
                                           h.p.maximum_resources_one_adult_can_harvest, h.p.season_length_in_years,
                                           h.p.minimum_adaptation, &errors);
            const double proportion = draw_regrowth(h.rng, h.t, patch_id, (uint32_t)(k - h.g.eco_off[patch_id]));
            h.res[k] = recover(res, h.max_res[k], h.p.resource_recovery_per_season, proportion);
        }
    }
    h.model_errors |= errors;
    h.storage.swap(stored_resources); /* `stored_resources = step(...)`, src/cli.rs:128 */
}

void step(oracle_handle &h) { /* lib.rs:477-496 and `s.t += 1`, src/cli.rs:154 */
    update_each_family_step(h);
    distribute_each_patch_resources_step(h);
    h.t += 1;
}

void build_mem_table(oracle_handle &h) {
    /* lib.rs:845-866: memory starts at 1 and is multiplied by 0.5^(season/2) per entry; entries are
     * examined while memory >= f64::EPSILON */
    const double memory_decay_per_season = std::pow(0.5, h.p.season_length_in_years / 2.0);
    h.mem_table.clear();
    double memory = 1.0;
    while (!(memory < std::numeric_limits<double>::epsilon()) && h.mem_table.size() < (1u << 20)) {
        h.mem_table.push_back(memory);
        memory *= memory_decay_per_season;
    }
    /* CANON stream v2: the head is made of the leading 16-entry blocks that hold an entry with memory_n >= 2^-8 (top
     * byte of ceil(memory_n * 2^40) non-zero), at most 16 blocks; the entries after it are sampled by skipping */
    uint32_t head_blocks = 0;
    for (size_t k = 0; k < h.mem_table.size(); ++k) {
        const double c = std::ceil(h.mem_table[k] * 1099511627776.0);
        if (c >= 4294967296.0) head_blocks = (uint32_t)(k / 16u) + 1u;
    }
    if (head_blocks > 16u) head_blocks = 16u;
    h.mem_tail0 = 16u * head_blocks;
    h.mem_tail.clear();
    double survive = 1.0;
    for (size_t k = h.mem_tail0; k < h.mem_table.size(); ++k) {
        survive = survive * (1.0 - h.mem_table[k]);
        h.mem_tail.push_back(survive);
    }
}

int fail(oracle_handle *h, int code, const char *msg) {
    if (h)
        h->err = msg;
    else
        g_create_error = msg;
    return code;
}

} // namespace

/* ------------------------------------------------------------------------------------------ *
 * C API
 * ------------------------------------------------------------------------------------------ */
extern "C" {

int oracle_abi_version(void) { return DSIM_ABI_VERSION; }

void oracle_default_params(dsim_params *p) { /* src/parameters.rs:22-54 */
    std::memset(p, 0, sizeof *p);
    p->resource_recovery_per_season = 0.10;
    p->culture_mutation_rate = 6e-3;
    p->culture_dimensionality = 20;
    p->cooperation_threshold = 6;
    p->maximum_resources_one_adult_can_harvest = 1e50;
    p->evidence_needed = 0.1;
    p->payoff_std = 0.1;
    p->minimum_adaptation = 0.5;
    p->fight_deadliness = 1u << 31;
    p->enemy_discount = std::pow(0.5, 0.2);
    p->season_length_in_years = 1. / 6.;
    p->cooperation_inclusive = 0;
}

void oracle_default_options(dsim_options *o) { std::memset(o, 0, sizeof *o); o->use_graphs = 1; }

int oracle_create(const dsim_params *p, const dsim_graph *g, uint64_t seed, const dsim_options *opt,
                  oracle_handle **out) {
    (void)opt;
    if (!p || !g || !out || !g->eco_off || !g->edge_off) return fail(nullptr, DSIM_ERR_INVALID, "null argument");
    std::unique_ptr<oracle_handle> h(new oracle_handle);
    h->p = *p;
    h->rng.key[0] = (uint32_t)seed;
    h->rng.key[1] = (uint32_t)(seed >> 32);
    Graph &G = h->g;
    G.n = g->n_nodes;
    const uint64_t n_eco = g->eco_off[G.n], n_edge = g->edge_off[G.n];
    if (g->h3) G.h3.assign(g->h3, g->h3 + G.n);
    if (g->lon) G.lon.assign(g->lon, g->lon + G.n);
    if (g->lat) G.lat.assign(g->lat, g->lat + G.n);
    G.eco_off.assign(g->eco_off, g->eco_off + G.n + 1);
    G.eco_id.assign(g->eco_id, g->eco_id + n_eco);
    if (g->popcap) G.popcap.assign(g->popcap, g->popcap + n_eco);
    G.edge_off.assign(g->edge_off, g->edge_off + G.n + 1);
    G.edge_dst.assign(g->edge_dst, g->edge_dst + n_edge);
    G.edge_sec.assign(g->edge_sec, g->edge_sec + n_edge);
    for (uint32_t v = 0; v < G.n; ++v)
        for (uint64_t k = G.eco_off[v]; k < G.eco_off[v + 1]; ++k) {
            if (G.eco_id[k] >= (uint32_t)N_ECO) return fail(nullptr, DSIM_ERR_INVALID, "ecoregion id out of range");
            if (k > G.eco_off[v] && G.eco_id[k] <= G.eco_id[k - 1])
                return fail(nullptr, DSIM_ERR_INVALID, "ecoregion ids must ascend within a node");
        }
    for (uint64_t e = 0; e < n_edge; ++e)
        if (G.edge_dst[e] >= G.n) return fail(nullptr, DSIM_ERR_INVALID, "edge target out of range");
    h->res.assign(n_eco, 0.0);
    h->max_res.assign(n_eco, 0.0);
    h->storage.assign(G.n, CultureMap());
    h->nearby_cache.resize(G.n);
    h->reach_cache.resize(G.n);
    build_mem_table(*h);
    *out = h.release();
    return DSIM_OK;
}

void oracle_destroy(oracle_handle *h) { delete h; }

const char *oracle_last_error(const oracle_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int oracle_set_mode(oracle_handle *h, int mode) {
    if (!h || mode < 0 || mode > 1) return DSIM_ERR_INVALID;
    h->mode = mode;
    return DSIM_OK;
}

/* OpenMP threads of the oracle's parallel loops (the rayon sites of lib.rs:527,607,615,1813).  n <= 0 restores the
 * runtime's default; returns the number of threads the next parallel region will use.  bench.py sets it explicitly:
 * torch.distributed.run exports OMP_NUM_THREADS=1 into every rank. */
int oracle_set_threads(int n) {
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int oracle_set_patches(oracle_handle *h, const double *res, const double *max_res) {
    if (!h || !res || !max_res) return DSIM_ERR_INVALID;
    std::copy(res, res + h->res.size(), h->res.begin());
    std::copy(max_res, max_res + h->max_res.size(), h->max_res.begin());
    return DSIM_OK;
}

int oracle_get_patches(oracle_handle *h, double *res, double *max_res) {
    if (!h) return DSIM_ERR_INVALID;
    if (res) std::copy(h->res.begin(), h->res.end(), res);
    if (max_res) std::copy(h->max_res.begin(), h->max_res.end(), max_res);
    return DSIM_OK;
}

int oracle_set_families(oracle_handle *h, const dsim_families *f) {
    if (!h || !f) return DSIM_ERR_INVALID;
    if (f->n && (!f->id || !f->culture || !f->location || !f->size || !f->stored))
        return fail(h, DSIM_ERR_INVALID, "missing family array");
    std::vector<Family> fam((size_t)f->n);
    uint64_t max_id = 0;
    for (uint64_t i = 0; i < f->n; ++i) {
        Family &x = fam[i];
        x.id = f->id[i];
        x.parent = f->parent_id ? f->parent_id[i] : DSIM_NO_PARENT;
        x.birth_index = f->birth_index ? f->birth_index[i] : 0;
        x.culture = f->culture[i];
        x.location = f->location[i];
        if (x.location >= h->g.n) return fail(h, DSIM_ERR_INVALID, "family location out of range");
        x.effective_size = f->size[i];
        x.stored_resources = f->stored[i];
        x.last_harvest = f->last_harvest ? f->last_harvest[i] : 0.0;
        x.seasons_till_next_adult = f->till_adult ? f->till_adult[i] : 0;
        x.has_mutation_clock = f->has_mutation_clock ? f->has_mutation_clock[i] != 0 : false;
        x.seasons_till_next_mutation = (f->till_mutation && x.has_mutation_clock) ? f->till_mutation[i] : 0;
        x.number_offspring = f->n_offspring ? f->n_offspring[i] : 0;
        if (f->hist_off)
            for (uint64_t k = f->hist_off[i]; k < f->hist_off[i + 1]; ++k)
                x.history.emplace_back(f->hist_patch[k], f->hist_harvest[k]);
        x.adaptation.fill(h->p.minimum_adaptation); /* Ecovector::from(min), lib.rs:1614 */
        if (f->adapt_off)
            for (uint64_t k = f->adapt_off[i]; k < f->adapt_off[i + 1]; ++k) {
                if (f->adapt_eco[k] >= (uint32_t)N_ECO) return fail(h, DSIM_ERR_INVALID, "adaptation ecoregion out of range");
                x.adaptation[f->adapt_eco[k]] = f->adapt_val[k];
            }
        if (i > 0 && x.id <= fam[i - 1].id) return fail(h, DSIM_ERR_INVALID, "family ids must ascend");
        max_id = x.id;
    }
    h->families.swap(fam);
    h->next_id = f->n ? max_id + 1 : 0;
    return DSIM_OK;
}

int oracle_family_counts(oracle_handle *h, uint64_t *n, uint64_t *n_hist, uint64_t *n_adapt) {
    if (!h) return DSIM_ERR_INVALID;
    uint64_t nh = 0, na = 0;
    for (const Family &f : h->families) {
        nh += f.history.size();
        for (double v : f.adaptation) na += (v != h->p.minimum_adaptation);
    }
    if (n) *n = h->families.size();
    if (n_hist) *n_hist = nh;
    if (n_adapt) *n_adapt = na;
    return DSIM_OK;
}

int oracle_get_families(oracle_handle *h, dsim_families *f) {
    if (!h || !f) return DSIM_ERR_INVALID;
    const uint64_t n = h->families.size();
    f->n = n;
    uint64_t nh = 0, na = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const Family &x = h->families[i];
        if (f->id) f->id[i] = x.id;
        if (f->parent_id) f->parent_id[i] = x.parent;
        if (f->birth_index) f->birth_index[i] = x.birth_index;
        if (f->culture) f->culture[i] = x.culture;
        if (f->location) f->location[i] = x.location;
        if (f->size) f->size[i] = (uint32_t)x.effective_size;
        if (f->stored) f->stored[i] = x.stored_resources;
        if (f->last_harvest) f->last_harvest[i] = x.last_harvest;
        if (f->till_adult) f->till_adult[i] = x.seasons_till_next_adult;
        if (f->till_mutation) f->till_mutation[i] = x.has_mutation_clock ? x.seasons_till_next_mutation : 0;
        if (f->has_mutation_clock) f->has_mutation_clock[i] = x.has_mutation_clock;
        if (f->n_offspring) f->n_offspring[i] = x.number_offspring;
        if (f->hist_off) {
            f->hist_off[i] = nh;
            for (const auto &e : x.history) {
                if (f->hist_patch) f->hist_patch[nh] = e.first;
                if (f->hist_harvest) f->hist_harvest[nh] = e.second;
                ++nh;
            }
        }
        if (f->adapt_off) {
            f->adapt_off[i] = na;
            for (int k = 0; k < N_ECO; ++k)
                if (x.adaptation[k] != h->p.minimum_adaptation) {
                    if (f->adapt_eco) f->adapt_eco[na] = (uint32_t)k;
                    if (f->adapt_val) f->adapt_val[na] = x.adaptation[k];
                    ++na;
                }
        }
    }
    if (f->hist_off) f->hist_off[n] = nh;
    if (f->adapt_off) f->adapt_off[n] = na;
    return DSIM_OK;
}

int oracle_set_time(oracle_handle *h, uint64_t t) {
    if (!h) return DSIM_ERR_INVALID;
    h->t = t;
    return DSIM_OK;
}
int oracle_get_time(oracle_handle *h, uint64_t *t) {
    if (!h || !t) return DSIM_ERR_INVALID;
    *t = h->t;
    return DSIM_OK;
}

int oracle_storage_count(oracle_handle *h, uint64_t *n) {
    if (!h || !n) return DSIM_ERR_INVALID;
    uint64_t c = 0;
    for (const CultureMap &m : h->storage) c += m.size();
    *n = c;
    return DSIM_OK;
}

int oracle_get_storage(oracle_handle *h, uint32_t *node, uint64_t *culture, uint64_t *size) {
    if (!h) return DSIM_ERR_INVALID;
    uint64_t c = 0;
    for (uint32_t v = 0; v < h->g.n; ++v) {
        CultureMap m = h->storage[v];
        std::sort(m.begin(), m.end());
        for (const auto &kv : m) {
            if (node) node[c] = v;
            if (culture) culture[c] = kv.first;
            if (size) size[c] = kv.second;
            ++c;
        }
    }
    return DSIM_OK;
}

int oracle_set_storage(oracle_handle *h, uint64_t n, const uint32_t *node, const uint64_t *culture,
                       const uint64_t *size) {
    if (!h || (n && (!node || !culture))) return DSIM_ERR_INVALID;
    for (CultureMap &m : h->storage) m.clear();
    for (uint64_t i = 0; i < n; ++i) {
        if (node[i] >= h->g.n) return fail(h, DSIM_ERR_INVALID, "storage node out of range");
        culture_map_insert(h->storage[node[i]], culture[i], size ? size[i] : 0);
    }
    return DSIM_OK;
}

static void fill_stats(oracle_handle *h, dsim_step_stats *stats) {
    if (!stats) return;
    std::memset(stats, 0, sizeof *stats);
    stats->t = h->t;
    stats->live_families = h->families.size();
    stats->births = h->births;
    stats->deaths = h->deaths;
    stats->family_updates = h->updates;
    stats->occupied_patches = h->occupied;
    stats->model_errors = h->model_errors;
}

int oracle_step(oracle_handle *h, uint32_t n_steps, dsim_step_stats *stats) {
    if (!h) return DSIM_ERR_INVALID;
    h->births = h->deaths = h->updates = 0;
    for (uint32_t s = 0; s < n_steps; ++s) step(*h);
    fill_stats(h, stats);
    return h->model_errors ? DSIM_ERR_MODEL : DSIM_OK;
}

int oracle_sync(oracle_handle *h) { return h ? DSIM_OK : DSIM_ERR_INVALID; }

int oracle_step_part(oracle_handle *h, int part) {
    if (!h) return DSIM_ERR_INVALID;
    if (part == 1)
        update_each_family_step(*h);
    else if (part == 2) {
        distribute_each_patch_resources_step(*h);
        h->t += 1;
    } else
        return DSIM_ERR_INVALID;
    return DSIM_OK;
}

int oracle_reach_counts(oracle_handle *h, uint64_t *n_nearby, uint64_t *n_reach) {
    if (!h) return DSIM_ERR_INVALID;
    uint64_t a = 0, b = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : a, b)
    for (long v = 0; v < (long)h->g.n; ++v) {
        a += nearby_locations(h->g, (uint32_t)v).size();
        b += many_nearby_locations(h->g, (uint32_t)v).size();
    }
    if (n_nearby) *n_nearby = a;
    if (n_reach) *n_reach = b;
    return DSIM_OK;
}

int oracle_get_reach(oracle_handle *h, uint64_t *near_off, uint32_t *near_dst, uint64_t *reach_off,
                     uint32_t *reach_dst, double *reach_cost) {
    if (!h) return DSIM_ERR_INVALID;
    uint64_t a = 0, b = 0;
    for (uint32_t v = 0; v < h->g.n; ++v) {
        const std::vector<uint32_t> nb = nearby_locations(h->g, v);
        const std::vector<std::pair<uint32_t, double>> rc = many_nearby_locations(h->g, v);
        if (near_off) near_off[v] = a;
        if (reach_off) reach_off[v] = b;
        for (uint32_t q : nb) {
            if (near_dst) near_dst[a] = q;
            ++a;
        }
        for (const auto &qc : rc) {
            if (reach_dst) reach_dst[b] = qc.first;
            if (reach_cost) reach_cost[b] = qc.second;
            ++b;
        }
    }
    if (near_off) near_off[h->g.n] = a;
    if (reach_off) reach_off[h->g.n] = b;
    return DSIM_OK;
}

/* --- primitives for the known-answer tests --------------------------------------------------- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr, key, out);
}
double oracle_log(double x) { return det_log(x); }
uint32_t oracle_f64_as_u32(double x) { return f64_as_u32(x); }

int oracle_use_resources_and_maybe_shrink(uint64_t *size, double *stored, double season) {
    return use_resources_and_maybe_shrink(*size, *stored, season) ? 1 : 0;
}

void oracle_maybe_grow(uint32_t *size, uint32_t *till_adult, double season) {
    Family f;
    f.effective_size = *size;
    f.seasons_till_next_adult = *till_adult;
    maybe_grow(f, season);
    *size = (uint32_t)f.effective_size;
    *till_adult = f.seasons_till_next_adult;
}

double oracle_recover(double res, double max_res, double rate, double u) { return recover(res, max_res, rate, 2.0 * u); }

int oracle_best_location(int n, const double *gain, const double *cost, const double *u, double *t_cost_out) {
    /* lib.rs:907-929 with the draws supplied */
    int target = -1;
    double max_gain = 0.0, t_cost = 0.0;
    double m = max_gain - t_cost;
    for (int i = 0; i < n; ++i) {
        const double g = gain[i] * u[i] - cost[i];
        if (g > m) {
            target = i;
            max_gain = oyr_max(max_gain, gain[i]);
            t_cost = cost[i];
            m = max_gain - t_cost;
        }
    }
    if (t_cost_out) *t_cost_out = t_cost;
    return target;
}

double oracle_exploit_one(int n, const uint32_t *size, const double *adapt, const int *band, double res, double cap,
                          double season, double *last_harvest) {
    /* families arrive with adaptation[0] = adapt[i] BEFORE the update; ecoregion 0 */
    std::vector<Family> fam((size_t)n);
    std::vector<PBand> groups;
    for (int i = 0; i < n; ++i) {
        fam[i].effective_size = size[i];
        fam[i].adaptation.fill(0.5);
        fam[i].adaptation[0] = adapt[i];
        if (groups.empty() || band[i] != band[i - 1]) groups.push_back(PBand{0, {}});
        groups.back().families.push_back(Member{&fam[i], (uint32_t)i});
    }
    uint32_t errors = 0;
    const double left = exploit_ecoregion(nullptr, groups, 0, res, cap, season, 0.5, &errors);
    for (int i = 0; i < n; ++i) last_harvest[i] = fam[i].last_harvest;
    return left;
}

double oracle_adapt_touch(double v, double minimum) { return adapt_entry_decay(v, minimum) + 0.05; }

uint32_t oracle_time_till_next_mutation(double u, double rate) { return time_till_next_mutation(u, rate); }

int oracle_will_cooperate(uint64_t a, uint64_t b, uint32_t threshold, int inclusive) {
    dsim_params p;
    oracle_default_params(&p);
    p.cooperation_threshold = threshold;
    p.cooperation_inclusive = inclusive ? 1 : 0;
    return will_cooperate(a, b, p) ? 1 : 0;
}

/* ---- the canonical draws, for tests/test_oracle_stats.py (distribution of each draw against the reference's formula) */
static Stream stream_of_seed(uint64_t seed) {
    Stream r;
    r.key[0] = (uint32_t)seed;
    r.key[1] = (uint32_t)(seed >> 32);
    return r;
}
/* kept[n] = 1 iff the family remembers its n-th newest history entry this step (lib.rs:854-866); returns n_mem */
uint32_t oracle_draw_remembered(double season, uint64_t seed, uint64_t step, uint64_t fid, uint32_t len, uint8_t *kept) {
    oracle_handle h;
    h.p.season_length_in_years = season;
    build_mem_table(h);
    std::vector<uint32_t> k;
    const uint32_t n = (uint32_t)std::min<size_t>(len, h.mem_table.size());
    remembered_entries(stream_of_seed(seed), step, fid, n, h.mem_table, h.mem_tail0, h.mem_tail, k);
    std::memset(kept, 0, len);
    for (uint32_t e : k) kept[e] = 1;
    return (uint32_t)h.mem_table.size();
}
double oracle_draw_noise(uint64_t seed, uint64_t step, uint64_t fid, uint32_t index) { return draw_noise(stream_of_seed(seed), step, fid, index); }
uint64_t oracle_draw_shuffle_key(uint64_t seed, uint64_t step, uint64_t fid) { return draw_shuffle_key(stream_of_seed(seed), step, fid); }
void oracle_draw_fight(uint64_t seed, uint64_t step, uint64_t fid, uint32_t position, uint32_t deadliness, int out[2]) {
    bool a, b;
    draw_fight(stream_of_seed(seed), step, fid, position, deadliness, &a, &b);
    out[0] = a;
    out[1] = b;
}
uint64_t oracle_draw_pivot(uint64_t seed, uint64_t step, uint64_t node, uint32_t band, uint64_t len) {
    return draw_pivot(stream_of_seed(seed), step, node, band, len);
}
double oracle_draw_regrowth(uint64_t seed, uint64_t step, uint64_t node, uint32_t slot) { return draw_regrowth(stream_of_seed(seed), step, node, slot); }
/* the mutation draws of one family and step through `mutate` itself (lib.rs:1754-1785): clock[0] = the clock drawn for a
 * family without one, clock[1] = the clock drawn after a flip, *bit = the bit flipped when the clock is at zero */
void oracle_draw_mutation(uint64_t seed, uint64_t step, uint64_t fid, double rate, uint32_t clock[2], uint32_t *bit) {
    oracle_handle h;
    oracle_default_params(&h.p);
    h.p.culture_mutation_rate = rate;
    h.rng = stream_of_seed(seed);
    h.t = step;
    Family f;
    f.id = fid;
    f.has_mutation_clock = false;
    mutate(h, f, 0);                      /* None: draws clock[0] (then decrements it, or flips at once when it is 0) */
    uint32_t a[4];
    h.rng.draw(step, fid, P_MUT, 0, a);
    clock[0] = time_till_next_mutation(u53(a[0], a[1]), rate);
    f = Family();
    f.id = fid;
    f.has_mutation_clock = true;
    f.seasons_till_next_mutation = 0;
    mutate(h, f, 0);                      /* Some(0): flips one bit of culture 0 and redraws */
    *bit = (uint32_t)__builtin_ctzll(f.culture);
    clock[1] = f.seasons_till_next_mutation;
}

void oracle_constants(double season, uint32_t *adult_reset, uint32_t *grow_reset, double *range_s, double *cost_per_s,
                      double *decay, uint32_t *n_mem) {
    oracle_handle h;
    h.p.season_length_in_years = season;
    build_mem_table(h);
    if (adult_reset) *adult_reset = f64_as_u32(15. / season) + 1;
    if (grow_reset) *grow_reset = f64_as_u32(YEARS_BETWEEN_ADULTS / season);
    if (range_s) *range_s = MAX_SEASON_RANGE;
    if (cost_per_s) *cost_per_s = 1. / SECONDS_PER_YEAR * 3.;
    if (decay) *decay = std::pow(0.5, season / 2.0);
    if (n_mem) *n_mem = (uint32_t)h.mem_table.size();
}

} /* extern "C" */
