/*
 * dsim_kernels.cu — the step loop of the dispersal model as sm_100a kernels.
 *
 * One season = `step` (lib.rs:477-496) = seven launches, replayed as a CUDA graph:
 *
 *   part 1, update_each_family_step (lib.rs:511-552)
 *     k_lifecycle   consume / shrink / grow per family (lib.rs:527-539, 1912-1937); survivor and
 *                   split bitmaps + per-256 counts (the `retain` of lib.rs:541 becomes a scan)
 *     k_control     single block: exclusive scan of the per-256 counts, step bookkeeping
 *     k_move        one warp per family: decide_on_moving + best_location (lib.rs:830-929),
 *                   destination_quality (lib.rs:1034-1062), move bookkeeping (lib.rs:1826-1833),
 *                   maybe_procreate + the child's own move (lib.rs:1840-1855, 1869-1897); writes the
 *                   compacted next family buffer; counts arrivals per node (histogram of the
 *                   counting sort by patch)
 *   part 2, distribute_each_patch_resources_step (lib.rs:591-655)
 *     k_segments    one thread per occupied node: claim a segment of `members`
 *     k_scatter     one thread per family: counting-sort scatter; clears the views of vacated nodes
 *     k_patch       one warp per occupied node: cooperate_or_fight (lib.rs:1372-1417), adjust_culture
 *                   / mutate (lib.rs:1754-1785, 1960-1970), exploit_patch + recover (lib.rs:2021-2109),
 *                   next step's census, quality and band cultures (lib.rs:522-525, 989-994, 622-633)
 *     k_advance     one thread: heavy-slot allocator bookkeeping, t += 1 (src/cli.rs:154)
 *
 * Floating point: doubles, IEEE operations only, compiled with -fmad=false so that every decision
 * (`g > m`, `res < max`, `stored - size*season < 0`) matches the CPU oracle bit for bit.
 */
#include "dsim_kernels.h"
#include "dsim_math.h"

#define LIFE_BLOCK 256
#define MOVE_THREADS 256
#define PATCH_WARPS 4
#define PATCH_M 64 /* families per node handled in shared memory; larger nodes spill to HBM scratch */

__device__ __forceinline__ dsim_stream stream_of(const ModelConst &mc) {
    dsim_stream s;
    s.k0 = mc.k0;
    s.k1 = mc.k1;
    return s;
}

/* resource_uncertainty (lib.rs:1317) for candidate `i` of family `fid`: two draws per Philox block */
__device__ __forceinline__ double noise_draw(const ModelConst &mc, uint32_t step, uint64_t fid, uint32_t i) {
    const dsim_u4 o = dsim_draw(stream_of(mc), step, fid, DSIM_P_NOISE, i >> 1);
    return (i & 1u) ? dsim_u53(o.z, o.w) : dsim_u53(o.x, o.y);
}

/* ------------------------------------------------------------------------------------------ *
 * part 1a: lifecycle
 * ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(LIFE_BLOCK) k_lifecycle(StepArgs a) {
    const uint32_t n = a.ctl->n_cur[a.parity];
    const uint32_t nvb = (n + LIFE_BLOCK - 1) / LIFE_BLOCK;
    __shared__ uint32_t s_alive[LIFE_BLOCK / 32], s_split[LIFE_BLOCK / 32];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const double season = a.mc.season;
    for (uint32_t vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
        const uint32_t i = vb * LIFE_BLOCK + threadIdx.x;
        bool alive = false, split = false;
        if (i < n) {
            uint32_t size = a.cur.size[i];
            double stored = a.cur.stored[i];
            uint32_t till_adult = a.cur.till_adult[i];
            /* use_resources_and_maybe_shrink, lib.rs:1923-1937 */
            bool has_shrunk = false;
            while ((stored - (double)size * season) < 0.0 && size > 0u) {
                size -= 1u;
                has_shrunk = true;
            }
            stored = stored - (double)size * season;
            if (has_shrunk) till_adult = a.mc.adult_reset; /* lib.rs:535 */
            /* maybe_grow, lib.rs:1912-1920 */
            if (till_adult == 0u) {
                size += 1u;
                till_adult = a.mc.grow_reset;
            } else {
                till_adult -= 1u;
            }
            a.cur.size[i] = size;
            a.cur.stored[i] = stored;
            a.cur.till_adult[i] = till_adult;
            alive = size >= 2u;  /* can_survive, lib.rs:1939-1941 */
            split = size >= 10u; /* maybe_procreate, lib.rs:1870 (the move does not change the size) */
        }
        const uint32_t ab = __ballot_sync(DSIM_FULL, alive);
        const uint32_t sb = __ballot_sync(DSIM_FULL, split);
        if (lane == 0) {
            a.sc.alive_bits[vb * (LIFE_BLOCK / 32) + w] = ab;
            a.sc.split_bits[vb * (LIFE_BLOCK / 32) + w] = sb;
            s_alive[w] = (uint32_t)__popc(ab);
            s_split[w] = (uint32_t)__popc(sb);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t sa = 0, ss = 0;
            for (int k = 0; k < LIFE_BLOCK / 32; ++k) {
                sa += s_alive[k];
                ss += s_split[k];
            }
            a.sc.blk_alive[vb] = sa;
            a.sc.blk_split[vb] = ss;
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------ *
 * part 1b: scan of the per-256 counts + bookkeeping (single block)
 * ------------------------------------------------------------------------------------------ */
#define CTL_THREADS 1024
__global__ void __launch_bounds__(CTL_THREADS) k_control(StepArgs a) {
    __shared__ uint32_t sa[CTL_THREADS], ss[CTL_THREADS];
    __shared__ uint32_t carry_a, carry_s;
    DevCtl *ctl = a.ctl;
    const uint32_t n = ctl->n_cur[a.parity];
    const uint32_t nvb = (n + LIFE_BLOCK - 1) / LIFE_BLOCK;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        carry_a = 0;
        carry_s = 0;
    }
    __syncthreads();
    for (uint32_t base = 0; base < nvb; base += CTL_THREADS) {
        const uint32_t k = base + tid;
        const uint32_t va = (k < nvb) ? a.sc.blk_alive[k] : 0u;
        const uint32_t vs = (k < nvb) ? a.sc.blk_split[k] : 0u;
        sa[tid] = va;
        ss[tid] = vs;
        __syncthreads();
        for (uint32_t off = 1; off < CTL_THREADS; off <<= 1) { /* inclusive Hillis-Steele */
            uint32_t xa = 0, xs = 0;
            if (tid >= off) {
                xa = sa[tid - off];
                xs = ss[tid - off];
            }
            __syncthreads();
            sa[tid] += xa;
            ss[tid] += xs;
            __syncthreads();
        }
        if (k < nvb) {
            a.sc.blk_alive[k] = carry_a + sa[tid] - va; /* exclusive */
            a.sc.blk_split[k] = carry_s + ss[tid] - vs;
        }
        __syncthreads();
        if (tid == 0) {
            carry_a += sa[CTL_THREADS - 1];
            carry_s += ss[CTL_THREADS - 1];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t n_alive = carry_a, n_children = carry_s;
        ctl->n_alive = n_alive;
        ctl->n_children = n_children;
        ctl->n_dead = n - n_alive;
        ctl->n_next = n_alive + n_children;
        ctl->child_pos_base = n_alive;
        ctl->child_id_base = ctl->next_id;
        ctl->next_id += n_children;
        ctl->n_cur[a.parity ^ 1u] = n_alive + n_children;
        ctl->n_occ[a.parity] = 0;
        ctl->seg_cursor = 0;
        ctl->births += n_children;
        ctl->deaths += n - n_alive;
        ctl->updates += n;
        if (n_alive + n_children > ctl->cap) {
            ctl->errors |= DSIM_MODEL_CAPACITY;
            ctl->n_next = ctl->cap;
            ctl->n_cur[a.parity ^ 1u] = ctl->cap;
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * part 1c: move + split
 * ------------------------------------------------------------------------------------------ */
struct Choice { /* best_location state, lib.rs:911-916 */
    uint32_t target;
    double max_gain, t_cost, m;
};

/* Ordered acceptance scan of lib.rs:918-927 over the (up to) 32 candidates held one per lane, in
 * lane order: a candidate is accepted iff g > m with m as left by the previous acceptance. */
__device__ __forceinline__ void accept_scan(Choice &ch, bool valid, double g, double gain, double cost, uint32_t q,
                                            uint32_t lane) {
    uint32_t cursor = 0;
    for (;;) {
        const uint32_t mask = __ballot_sync(DSIM_FULL, valid && lane >= cursor && g > ch.m);
        if (mask == 0u) break;
        const int L = __ffs((int)mask) - 1;
        const double gain_l = __shfl_sync(DSIM_FULL, gain, L);
        const double cost_l = __shfl_sync(DSIM_FULL, cost, L);
        const uint32_t q_l = __shfl_sync(DSIM_FULL, q, L);
        ch.target = q_l;
        ch.max_gain = dsim_oyr_max(ch.max_gain, gain_l);
        ch.t_cost = cost_l;
        ch.m = ch.max_gain - ch.t_cost;
        cursor = (uint32_t)L + 1u;
    }
}

/* destination_quality, lib.rs:1034-1062 (has_enemies lib.rs:996-1017, expected_harvest 989-994) */
__device__ __forceinline__ double dest_quality(const StepArgs &a, uint32_t q, uint64_t culture, uint32_t size,
                                               bool same_patch) {
    const PatchView v = a.ps.view[q];
    bool enemy = false;
    if (v.nb > 0u) {
        enemy = !dsim_will_cooperate(culture, v.cult0, a.mc.threshold, a.mc.inclusive);
        for (uint32_t k = 1; k < v.nb && !enemy; ++k)
            enemy = !dsim_will_cooperate(culture, a.ps.band_cult[v.stor_off + k], a.mc.threshold, a.mc.inclusive);
    }
    if (enemy) return 0.0;
    const uint64_t pop_at_destination = (uint64_t)v.pop + (same_patch ? 0u : size);
    return v.quality / (double)pop_at_destination;
}

/* travel cost look-up in the reach row [r0, r1) (replaces travel_costs.get(), lib.rs:869) */
__device__ __forceinline__ bool reach_lookup(const ReachTable &rt, uint32_t r0, uint32_t r1, uint32_t q, double *cost) {
    uint32_t lo = r0, hi = r1;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (rt.dst[mid] < q)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo < r1 && rt.dst[lo] == q) {
        *cost = rt.cost[lo];
        return true;
    }
    return false;
}

/* The remembered-patch part of decide_on_moving (lib.rs:854-867): history entry n (0 = newest) is
 * kept with probability memory_n; kept entries with a travel cost become candidates whose expected
 * harvest is the remembered per-capita harvest.  With `virt`, entry 0 is (v0_patch, v0_harv) and
 * entry n >= 1 is ring entry n-1 (a child inheriting its parent's history, lib.rs:1882-1885). */
__device__ __forceinline__ void history_pass(const StepArgs &a, uint32_t step, uint32_t lane, uint64_t fid, uint32_t r0,
                                             uint32_t r1, uint32_t n_sensed, uint32_t hs, uint32_t ring_cnt, uint32_t nh,
                                             bool virt, uint32_t v0_patch, double v0_harv, Choice &ch) {
    const uint32_t hcap = a.hv.hcap;
    const size_t hbase = (size_t)hs * hcap;
    for (uint32_t n0 = 0; n0 < nh; n0 += 128u) {
        /* memory draws: entry n uses Philox block (n>>7)*32 + (n&31), word (n>>5)&3 */
        const dsim_u4 o = dsim_draw(stream_of(a.mc), step, fid, DSIM_P_MEM, (n0 >> 7) * 32u + lane);
#pragma unroll
        for (uint32_t w = 0; w < 4u; ++w) {
            const uint32_t nb0 = n0 + 32u * w;
            if (nb0 >= nh) break;
            const uint32_t n = nb0 + lane;
            bool keep = false;
            if (n < nh) keep = ((double)dsim_word(o, w) * (1.0 / 4294967296.0)) < a.mc.mem_table[n];
            if (!__any_sync(DSIM_FULL, keep)) continue;
            bool cand = false;
            double g = 0.0, gain = 0.0, cost = 0.0;
            uint32_t q = 0;
            if (keep) {
                if (virt && n == 0u) {
                    q = v0_patch;
                    gain = v0_harv;
                } else {
                    const uint32_t nn = virt ? n - 1u : n;
                    const uint32_t pos = (ring_cnt - 1u - nn) % hcap;
                    q = a.hv.hist_patch[hbase + pos];
                    gain = a.hv.hist_harv[hbase + pos];
                }
                if (reach_lookup(a.rt, r0, r1, q, &cost)) {
                    const double u = noise_draw(a.mc, step, fid, n_sensed + n);
                    g = gain * u - cost;
                    cand = true;
                }
            }
            accept_scan(ch, cand, g, gain, cost, q, lane);
        }
    }
}

/* number of set bits of a 256-family bitmap group that precede family i, plus the group's offset;
 * *flag receives bit i.  Warp-uniform i. */
__device__ __forceinline__ uint32_t rank_before(const uint32_t *bits, const uint32_t *blk_off, uint32_t i, uint32_t lane,
                                                bool *flag) {
    const uint32_t wi = (i >> 5) & 7u;
    uint32_t word = 0, v = 0;
    if (lane < 8u) {
        word = bits[(i >> 8) * 8u + lane];
        if (lane < wi)
            v = (uint32_t)__popc(word);
        else if (lane == wi)
            v = (uint32_t)__popc(word & ((1u << (i & 31u)) - 1u));
    }
    const uint32_t myword = __shfl_sync(DSIM_FULL, word, (int)wi);
    v += __shfl_xor_sync(DSIM_FULL, v, 4);
    v += __shfl_xor_sync(DSIM_FULL, v, 2);
    v += __shfl_xor_sync(DSIM_FULL, v, 1);
    v = __shfl_sync(DSIM_FULL, v, 0);
    *flag = ((myword >> (i & 31u)) & 1u) != 0u;
    return blk_off[i >> 8] + v;
}

__global__ void __launch_bounds__(MOVE_THREADS) k_move(StepArgs a) {
    DevCtl *ctl = a.ctl;
    const uint32_t n = ctl->n_cur[a.parity];
    const uint32_t step = ctl->t;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t cap = ctl->cap;
    const uint32_t hcap = a.hv.hcap, acap = a.hv.acap;

    for (uint32_t i = gwarp; i < n; i += nwarps) {
        bool alive, split;
        const uint32_t newpos = rank_before(a.sc.alive_bits, a.sc.blk_alive, i, lane, &alive);
        const uint32_t crank = rank_before(a.sc.split_bits, a.sc.blk_split, i, lane, &split);
        if (!alive) { /* dropped by `retain`, lib.rs:541: hand the heavy slot back */
            if (lane == 0) a.sc.dead_hs[i - newpos] = a.cur.hs[i];
            continue;
        }
        if (newpos >= cap) continue; /* capacity error already flagged by k_control */

        const uint64_t fid = a.cur.id[i];
        const uint64_t culture = a.cur.culture[i];
        double stored = a.cur.stored[i];
        const double last_harvest = a.cur.last_harvest[i];
        const uint32_t loc = a.cur.loc[i];
        uint32_t size = a.cur.size[i];
        const uint32_t till_adult = a.cur.till_adult[i];
        const uint32_t till_mut = a.cur.till_mut[i];
        uint32_t misc = a.cur.misc[i];
        const uint32_t hs = a.cur.hs[i];

        /* ---- decide_on_moving, lib.rs:830-889 ------------------------------------------------ */
        Choice ch;
        ch.target = loc;
        ch.max_gain = 0.0;
        ch.t_cost = 0.0;
        ch.m = 0.0;
        const uint32_t r0 = a.rt.off[loc], r1 = a.rt.off[loc + 1];
        const uint32_t n_near = a.rt.near_off[loc + 1] - a.rt.near_off[loc];
        /* sensed patches: the entries of the reach row that are in the 2-hop list, in ascending
         * node order; their noise index is their position in that list */
        for (uint32_t base = r0; base < r1; base += 32u) {
            const uint32_t e = base + lane;
            bool cand = false;
            double g = 0.0, gain = 0.0, cost = 0.0;
            uint32_t q = 0;
            if (e < r1) {
                const uint32_t nidx = a.rt.nidx[e];
                if (nidx != DSIM_NIDX_NONE) {
                    q = a.rt.dst[e];
                    cost = a.rt.cost[e];
                    gain = dest_quality(a, q, culture, size, q == loc);
                    g = gain * noise_draw(a.mc, step, fid, nidx) - cost;
                    cand = true;
                }
            }
            accept_scan(ch, cand, g, gain, cost, q, lane);
        }
        const uint32_t hist_cnt = a.hv.hist_cnt[hs];
        const uint32_t hist_valid = hist_cnt < hcap ? hist_cnt : hcap;
        const uint32_t nh = hist_valid < a.mc.n_mem ? hist_valid : a.mc.n_mem;
        history_pass(a, step, lane, fid, r0, r1, n_near, hs, hist_cnt, nh, false, 0u, 0.0, ch);

        /* ---- move bookkeeping, lib.rs:1826-1833 ---------------------------------------------- */
        const double harv0 = last_harvest / (double)size; /* pushed to the history */
        stored = stored + last_harvest;
        const uint32_t newloc = ch.target;
        stored = stored - ch.t_cost;

        /* ---- maybe_procreate + the child's move, lib.rs:1840-1855, 1869-1897 ------------------- */
        if (split) {
            uint32_t n_off = ((misc & 0xFFFFu) + 1u) & 0xFFFFu;
            misc = (misc & ~0xFFFFu) | n_off;
            const double take = stored * 2.0 / (double)size;
            stored = stored - take;
            size -= 2u;
            const uint64_t cid = ctl->child_id_base + crank;
            const uint32_t cpos = ctl->child_pos_base + crank;
            const uint32_t ft = ctl->free_top;
            const uint32_t chs = (crank < ft) ? a.sc.free_hs[ft - 1u - crank] : ctl->hs_hwm + (crank - ft);

            Choice cc;
            cc.target = newloc;
            cc.max_gain = 0.0;
            cc.t_cost = 0.0;
            cc.m = 0.0;
            const uint32_t c0 = a.rt.off[newloc], c1 = a.rt.off[newloc + 1];
            /* `&nearby[1..]` of the parent's OLD patch (lib.rs:1846), costs from the child's patch */
            const uint32_t nl0 = a.rt.near_off[loc], nl1 = a.rt.near_off[loc + 1];
            const uint32_t first = nl0 + ((nl1 > nl0) ? 1u : 0u);
            for (uint32_t base = first; base < nl1; base += 32u) {
                const uint32_t j = base + lane;
                bool cand = false;
                double g = 0.0, gain = 0.0, cost = 0.0;
                uint32_t q = 0;
                if (j < nl1) {
                    q = a.rt.near_dst[j];
                    if (reach_lookup(a.rt, c0, c1, q, &cost)) {
                        gain = dest_quality(a, q, culture, 2u, q == newloc);
                        g = gain * noise_draw(a.mc, step, cid, j - first) - cost;
                        cand = true;
                    }
                }
                accept_scan(cc, cand, g, gain, cost, q, lane);
            }
            /* child history = newest min(len, time_to_grow_adult) entries incl. the one just pushed */
            uint32_t child_len = hist_valid + 1u;
            if (child_len > a.mc.adult_reset) child_len = a.mc.adult_reset;
            if (child_len > hcap) child_len = hcap;
            const uint32_t nhc = child_len < a.mc.n_mem ? child_len : a.mc.n_mem;
            history_pass(a, step, lane, cid, c0, c1, nl1 - first, hs, hist_cnt, nhc, true, loc, harv0, cc);

            if (cpos < cap && chs < ctl->hs_cap) {
                const size_t pb = (size_t)hs * hcap, cb = (size_t)chs * hcap;
                for (uint32_t k = lane; k < child_len; k += 32u) { /* chronological index k */
                    const uint32_t nn = child_len - 1u - k;        /* newest-first index      */
                    uint32_t hp;
                    double hh;
                    if (nn == 0u) {
                        hp = loc;
                        hh = harv0;
                    } else {
                        const uint32_t pos = (hist_cnt - 1u - (nn - 1u)) % hcap;
                        hp = a.hv.hist_patch[pb + pos];
                        hh = a.hv.hist_harv[pb + pos];
                    }
                    a.hv.hist_patch[cb + k] = hp;
                    a.hv.hist_harv[cb + k] = hh;
                }
                for (uint32_t k = lane; k < acap; k += 32u) { /* adaptation: family.adaptation, lib.rs:1893 */
                    a.hv.a_eco[(size_t)chs * acap + k] = a.hv.a_eco[(size_t)hs * acap + k];
                    a.hv.a_cnt[(size_t)chs * acap + k] = a.hv.a_cnt[(size_t)hs * acap + k];
                    a.hv.a_val[(size_t)chs * acap + k] = a.hv.a_val[(size_t)hs * acap + k];
                }
                if (lane == 0) {
                    a.hv.hist_cnt[chs] = child_len;
                    a.hv.decay[chs] = a.hv.decay[hs];
                    a.hv.parent[chs] = fid;
                    a.hv.birth_index[chs] = (uint16_t)n_off;
                    a.next.id[cpos] = cid;
                    a.next.culture[cpos] = culture;
                    a.next.stored[cpos] = take - cc.t_cost; /* descendant.stored_resources -= cost */
                    a.next.last_harvest[cpos] = 0.0;
                    a.next.loc[cpos] = cc.target;
                    a.next.size[cpos] = 2u;
                    a.next.till_adult[cpos] = a.mc.adult_reset;
                    a.next.till_mut[cpos] = 0u;
                    a.next.misc[cpos] = 0u;
                    a.next.hs[cpos] = chs;
                    const uint32_t slot = atomicAdd(&a.ps.cnt[cc.target], 1u);
                    if (slot == 0u) a.sc.occ[a.parity][atomicAdd(&ctl->n_occ[a.parity], 1u)] = cc.target;
                    a.sc.fslot[cpos] = slot;
                }
            } else if (lane == 0) {
                atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
            }
            __syncwarp();
        }

        /* every lane has read the family's record and history by now; without this barrier lane 0
         * could bump hist_cnt / overwrite the oldest ring entry under a slower lane */
        __syncwarp();
        if (lane == 0) {
            /* family.history.push((location, last_harvest / size)), lib.rs:1826-1829 */
            const uint32_t pos = hist_cnt % hcap;
            a.hv.hist_patch[(size_t)hs * hcap + pos] = loc;
            a.hv.hist_harv[(size_t)hs * hcap + pos] = harv0;
            a.hv.hist_cnt[hs] = hist_cnt + 1u;
            a.next.id[newpos] = fid;
            a.next.culture[newpos] = culture;
            a.next.stored[newpos] = stored;
            a.next.last_harvest[newpos] = 0.0;
            a.next.loc[newpos] = newloc;
            a.next.size[newpos] = size;
            a.next.till_adult[newpos] = till_adult;
            a.next.till_mut[newpos] = till_mut;
            a.next.misc[newpos] = misc;
            a.next.hs[newpos] = hs;
            const uint32_t slot = atomicAdd(&a.ps.cnt[newloc], 1u);
            if (slot == 0u) a.sc.occ[a.parity][atomicAdd(&ctl->n_occ[a.parity], 1u)] = newloc;
            a.sc.fslot[newpos] = slot;
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * part 2a/2b: counting sort by node (families_by_location, lib.rs:602-612)
 * ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_segments(StepArgs a) {
    DevCtl *ctl = a.ctl;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t n_occ = ctl->n_occ[a.parity];
    for (uint32_t k = tid; k < n_occ; k += nth) {
        const uint32_t p = a.sc.occ[a.parity][k];
        a.ps.seg[p] = atomicAdd(&ctl->seg_cursor, a.ps.cnt[p]);
    }
    /* heavy slots vacated this step go onto the free stack, above what the children did not pop */
    const uint32_t ft = ctl->free_top, births = ctl->n_children, dead = ctl->n_dead;
    const uint32_t popped = births < ft ? births : ft;
    const uint32_t base = ft - popped;
    for (uint32_t k = tid; k < dead; k += nth) a.sc.free_hs[base + k] = a.sc.dead_hs[k];
}

__global__ void __launch_bounds__(256) k_scatter(StepArgs a) {
    DevCtl *ctl = a.ctl;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t n = ctl->n_next;
    for (uint32_t j = tid; j < n; j += nth) a.sc.members[a.ps.seg[a.next.loc[j]] + a.sc.fslot[j]] = j;
    /* nodes occupied last step and empty now: census 0, no band cultures (lib.rs:522-525, 614) */
    const uint32_t n_prev = ctl->n_occ[a.parity ^ 1u];
    for (uint32_t k = tid; k < n_prev; k += nth) {
        const uint32_t p = a.sc.occ[a.parity ^ 1u][k];
        if (a.ps.cnt[p] == 0u) {
            a.ps.view[p].pop = 0u;
            a.ps.view[p].nb = 0u;
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * part 2c: one warp per occupied node
 * ------------------------------------------------------------------------------------------ */
struct PatchMem { /* per-warp working arrays: shared memory, or HBM scratch for crowded nodes */
    uint32_t *idx, *tidx, *size, *band, *bcnt, *btot;
    uint64_t *key, *fid, *tfid, *cult, *bcult;
    double *stored, *contrib, *lh, *beff, *bcoef;
    uint8_t *balive;
};

/* adaptation of family `hs` to ecoregion `eco` after one more `adaptation * 0.95` and `+= 0.05`
 * (lib.rs:1953-1954, src/ecology.rs:147-161), on the sparse lazily-decayed representation: a slot
 * stores (ecoregion, value, decay count at which the value was current); untracked ecoregions sit at
 * the floor, where the decay is idempotent. */
__device__ __forceinline__ double adapt_touch(const StepArgs &a, uint32_t hs, uint32_t eco, uint32_t *errors) {
    const uint32_t acap = a.hv.acap;
    const size_t base = (size_t)hs * acap;
    const double minimum = a.mc.min_adapt;
    const uint32_t d = a.hv.decay[hs];
    int found = -1, spare = -1;
    for (uint32_t k = 0; k < acap; ++k) {
        const uint32_t e = a.hv.a_eco[base + k];
        if (e == eco) {
            found = (int)k;
            break;
        }
        if (e == DSIM_ECO_EMPTY && spare < 0) spare = (int)k;
    }
    double v = minimum;
    int slot = found;
    if (found >= 0) {
        v = a.hv.a_val[base + found];
        for (uint32_t r = d - a.hv.a_cnt[base + found]; r > 0u && v != minimum; --r) v = dsim_adapt_decay(v, minimum);
    } else {
        slot = spare;
        if (slot < 0) { /* reclaim a slot whose value has decayed to the floor */
            for (uint32_t k = 0; k < acap && slot < 0; ++k) {
                double x = a.hv.a_val[base + k];
                for (uint32_t r = d - a.hv.a_cnt[base + k]; r > 0u && x != minimum; --r) x = dsim_adapt_decay(x, minimum);
                if (x == minimum) slot = (int)k;
            }
        }
    }
    v = dsim_adapt_decay(v, minimum); /* this step's `* 0.95` */
    v = v + 0.05;                     /* entries[ecoregion] += 0.05 */
    if (slot >= 0) {
        a.hv.a_eco[base + slot] = (uint16_t)eco;
        a.hv.a_cnt[base + slot] = d + 1u;
        a.hv.a_val[base + slot] = v;
    } else {
        *errors |= DSIM_MODEL_ADAPTATION_OVERFLOW;
    }
    a.hv.decay[hs] = d + 1u;
    return v;
}

__global__ void __launch_bounds__(PATCH_WARPS * 32) k_patch(StepArgs a) {
    __shared__ uint32_t s_u32[PATCH_WARPS][6][PATCH_M];
    __shared__ uint64_t s_u64[PATCH_WARPS][5][PATCH_M];
    __shared__ double s_f64[PATCH_WARPS][5][PATCH_M];
    __shared__ uint8_t s_u8[PATCH_WARPS][PATCH_M];

    DevCtl *ctl = a.ctl;
    const uint32_t step = ctl->t;
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t gwarp = blockIdx.x * PATCH_WARPS + wib, nwarps = gridDim.x * PATCH_WARPS;
    const uint32_t n_occ = ctl->n_occ[a.parity];
    const dsim_stream rs = stream_of(a.mc);
    const double season = a.mc.season, minimum_unused = a.mc.min_adapt;
    (void)minimum_unused;
    uint32_t errors = 0;

    for (uint32_t vw = gwarp; vw < n_occ; vw += nwarps) {
        const uint32_t p = a.sc.occ[a.parity][vw];
        const uint32_t n = a.ps.cnt[p];
        const uint32_t s0 = a.ps.seg[p];
        PatchMem M;
        if (n <= PATCH_M) {
            M.idx = s_u32[wib][0]; M.tidx = s_u32[wib][1]; M.size = s_u32[wib][2]; M.band = s_u32[wib][3];
            M.bcnt = s_u32[wib][4]; M.btot = s_u32[wib][5];
            M.key = s_u64[wib][0]; M.fid = s_u64[wib][1]; M.tfid = s_u64[wib][2]; M.cult = s_u64[wib][3];
            M.bcult = s_u64[wib][4];
            M.stored = s_f64[wib][0]; M.contrib = s_f64[wib][1]; M.lh = s_f64[wib][2]; M.beff = s_f64[wib][3];
            M.bcoef = s_f64[wib][4];
            M.balive = s_u8[wib];
        } else {
            M.idx = a.sc.x_idx + s0; M.tidx = a.sc.x_tidx + s0; M.size = a.sc.x_size + s0; M.band = a.sc.x_band + s0;
            M.bcnt = a.sc.x_bcnt + s0; M.btot = a.sc.x_btot + s0;
            M.key = a.sc.x_key + s0; M.fid = a.sc.x_fid + s0; M.tfid = a.sc.x_tfid + s0; M.cult = a.sc.x_cult + s0;
            M.bcult = a.sc.x_bcult + s0;
            M.stored = a.sc.x_stored + s0; M.contrib = a.sc.x_contrib + s0; M.lh = a.sc.x_lh + s0;
            M.beff = a.sc.x_beff + s0; M.bcoef = a.sc.x_bcoef + s0;
            M.balive = a.sc.x_balive + s0;
        }
        __syncwarp();

        /* ---- stochasticity::shuffle (lib.rs:1376): order by (keyed random number, id) ----------- */
        for (uint32_t j = lane; j < n; j += 32u) {
            const uint32_t i = a.sc.members[s0 + j];
            const uint64_t id = a.next.id[i];
            const dsim_u4 o = dsim_draw(rs, step, id, DSIM_P_SHUF, 0u);
            M.tidx[j] = i;
            M.tfid[j] = id;
            M.key[j] = ((uint64_t)o.y << 32) | o.x;
        }
        __syncwarp();
        for (uint32_t j = lane; j < n; j += 32u) {
            const uint64_t kj = M.key[j], ij = M.tfid[j];
            uint32_t r = 0;
            for (uint32_t k = 0; k < n; ++k) {
                const uint64_t kk = M.key[k], ik = M.tfid[k];
                r += (kk < kj || (kk == kj && ik < ij)) ? 1u : 0u;
            }
            M.idx[r] = M.tidx[j];
            M.fid[r] = ij;
        }
        __syncwarp();
        for (uint32_t r = lane; r < n; r += 32u) {
            const uint32_t i = M.idx[r];
            M.cult[r] = a.next.culture[i];
            M.size[r] = a.next.size[i];
            M.stored[r] = a.next.stored[i];
            M.band[r] = 0xFFFFFFFFu;
        }
        __syncwarp();

        /* ---- cooperate_or_fight, lib.rs:1380-1399; encounter_maybe_join, lib.rs:1209-1232 -------- */
        uint32_t nb = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint64_t cf = M.cult[i], fid = M.fid[i];
            uint32_t sz = M.size[i];
            double st = M.stored[i];
            int joined = -1;
            for (uint32_t B = 0; B < nb && joined < 0; ++B) {
                bool any_hostile = false;
                for (uint32_t j0 = 0; j0 < i; j0 += 32u) {
                    const uint32_t j = j0 + lane;
                    bool hostile = false;
                    if (j < i) hostile = (M.band[j] == B) && !dsim_will_cooperate(cf, M.cult[j], a.mc.threshold, a.mc.inclusive);
                    uint32_t mask = __ballot_sync(DSIM_FULL, hostile);
                    while (mask != 0u) { /* hostile members in join order */
                        const uint32_t jj = j0 + (uint32_t)(__ffs((int)mask) - 1);
                        mask &= mask - 1u;
                        any_hostile = true;
                        uint32_t so = M.size[jj];
                        double sto = M.stored[jj];
                        __syncwarp();
                        const uint32_t reduction = sz < so ? sz : so;
                        const dsim_u4 o = dsim_draw(rs, step, fid, DSIM_P_FIGHT, jj);
                        if (o.x < a.mc.deadliness) { /* fight_is_deadly, lib.rs:1219 */
                            so -= reduction;
                            if (so == 0u) {
                                st = st + sto;
                                sto = 0.0;
                            }
                        }
                        if (o.y < a.mc.deadliness) sz -= reduction; /* lib.rs:1226 */
                        if (lane == 0) {
                            M.size[jj] = so;
                            M.stored[jj] = sto;
                        }
                        __syncwarp();
                    }
                }
                if (!any_hostile) joined = (int)B;
            }
            if (joined < 0) joined = (int)nb++;
            if (lane == 0) {
                M.band[i] = (uint32_t)joined;
                M.size[i] = sz;
                M.stored[i] = st;
            }
            __syncwarp();
        }

        /* ---- band culture = culture of a random member; drop empty bands, lib.rs:1400-1416 -------- */
        for (uint32_t B = lane; B < nb; B += 32u) {
            uint32_t cnt = 0, tot = 0;
            for (uint32_t j = 0; j < n; ++j)
                if (M.band[j] == B) {
                    cnt += 1u;
                    tot += M.size[j];
                }
            const dsim_u4 o = dsim_draw(rs, step, (uint64_t)p, DSIM_P_PIVOT, B);
            const uint32_t pick = (uint32_t)__umul64hi(((uint64_t)o.y << 32) | o.x, (uint64_t)cnt);
            uint64_t c = 0;
            uint32_t seen = 0;
            for (uint32_t j = 0; j < n; ++j)
                if (M.band[j] == B) {
                    if (seen == pick) {
                        c = M.cult[j];
                        break;
                    }
                    ++seen;
                }
            M.bcult[B] = c;
            M.bcnt[B] = cnt;
            M.btot[B] = tot;
            M.balive[B] = tot > 0u ? 1 : 0;
        }
        __syncwarp();

        /* ---- storage' (lib.rs:622-633): cultures and sizes of the surviving bands, creation order -- */
        uint32_t nb_alive = 0;
        uint64_t cult0 = 0;
        for (uint32_t B0 = 0; B0 < nb; B0 += 32u) {
            const uint32_t B = B0 + lane;
            const bool al = (B < nb) && M.balive[B];
            const uint32_t mask = __ballot_sync(DSIM_FULL, al);
            if (al) {
                const uint32_t k = nb_alive + (uint32_t)__popc(mask & ((1u << lane) - 1u));
                a.ps.band_cult[s0 + k] = M.bcult[B];
                a.ps.band_size[s0 + k] = M.btot[B];
            }
            if (nb_alive == 0u && mask != 0u) cult0 = __shfl_sync(DSIM_FULL, al ? M.bcult[B] : 0ull, __ffs((int)mask) - 1);
            nb_alive += (uint32_t)__popc(mask);
        }

        /* ---- adjust_culture / mutate, lib.rs:1960-1970, 1754-1785 ------------------------------- */
        for (uint32_t r = lane; r < n; r += 32u) {
            const uint32_t B = M.band[r];
            if (!M.balive[B]) continue;
            const uint32_t i = M.idx[r];
            uint64_t c = M.bcult[B];
            uint32_t misc = a.next.misc[i];
            uint32_t clock = a.next.till_mut[i];
            const dsim_u4 o = dsim_draw(rs, step, M.fid[r], DSIM_P_MUT, 0u);
            if (!(misc & DSIM_MISC_CLOCK)) { /* None => draw, then re-enter, lib.rs:1763-1774 */
                misc |= DSIM_MISC_CLOCK;
                clock = dsim_time_till_next_mutation(dsim_u53(o.x, o.y), a.mc.log1m_rate);
            }
            if (clock == 0u) { /* Some(0): flip one of 64 bits, redraw, lib.rs:1775-1780, 1277-1285 */
                c ^= (uint64_t)1 << (o.z >> 26);
                const dsim_u4 o2 = dsim_draw(rs, step, M.fid[r], DSIM_P_MUT, 1u);
                clock = dsim_time_till_next_mutation(dsim_u53(o2.x, o2.y), a.mc.log1m_rate);
            } else {
                clock -= 1u;
            }
            M.cult[r] = c;
            a.next.misc[i] = misc;
            a.next.till_mut[i] = clock;
        }
        __syncwarp();

        /* ---- exploit_patch + recover per ecoregion, lib.rs:2021-2067, 2095-2109 ------------------ */
        const uint32_t e0 = a.ps.eco_off[p], e1 = a.ps.eco_off[p + 1];
        double quality = 0.0;
        for (uint32_t e = e0; e < e1; ++e) {
            const uint32_t eco = a.ps.eco_id[e];
            for (uint32_t r = lane; r < n; r += 32u) { /* raw_family_contribution, lib.rs:1947-1958 */
                double c = 0.0;
                if (M.balive[M.band[r]]) c = (double)M.size[r] * adapt_touch(a, a.next.hs[M.idx[r]], eco, &errors);
                M.contrib[r] = c;
            }
            __syncwarp();
            for (uint32_t B = lane; B < nb; B += 32u) { /* sum_effort in join order */
                double eff = 0.0;
                if (M.balive[B])
                    for (uint32_t j = 0; j < n; ++j)
                        if (M.band[j] == B) eff = eff + M.contrib[j];
                M.beff[B] = eff;
            }
            __syncwarp();
            double total_squared = 1.0;
            for (uint32_t B = 0; B < nb; ++B)
                if (M.balive[B]) total_squared = total_squared + M.beff[B] * M.beff[B];
            double res = a.ps.res[e];
            const double mx = a.ps.max_res[e];
            for (uint32_t B = 0; B < nb; ++B) {
                if (!M.balive[B]) continue;
                const double eff = M.beff[B];
                if (!(eff > 0.0)) errors |= DSIM_MODEL_EFFORT_NOT_POSITIVE;
                const double harvest = dsim_oyr_min(res * (eff * eff) / total_squared, a.mc.cap_harvest * season * eff);
                if (!(harvest > 0.0)) errors |= DSIM_MODEL_HARVEST_NOT_POSITIVE;
                const double coefficient = dsim_oyr_min(harvest / eff, 2.5);
                if (!(coefficient > 0.0)) errors |= DSIM_MODEL_COEFF_NOT_POSITIVE;
                if (lane == 0) M.bcoef[B] = coefficient;
                res = res - coefficient * eff;
                if (!(res >= 0.0)) errors |= DSIM_MODEL_RESOURCES_NEGATIVE;
            }
            __syncwarp();
            for (uint32_t r = lane; r < n; r += 32u) {
                const uint32_t B = M.band[r];
                if (M.balive[B]) M.lh[r] = M.bcoef[B] * season * M.contrib[r];
            }
            /* recover, lib.rs:2100-2108 */
            if (res < mx) {
                const dsim_u4 o = dsim_draw(rs, step, (uint64_t)p, DSIM_P_REGROW, e - e0);
                const double proportion = 2.0 * dsim_u53(o.x, o.y);
                res = dsim_oyr_min(mx, mx * a.mc.recovery * proportion + res);
            }
            if (lane == 0) a.ps.res[e] = res;
            quality = quality + (res + mx * a.mc.recovery); /* expected_harvest for the next step */
            __syncwarp();
        }

        /* ---- write back ----------------------------------------------------------------------- */
        uint32_t pop = 0;
        for (uint32_t r = lane; r < n; r += 32u) {
            const uint32_t i = M.idx[r];
            const uint32_t sz = M.size[r];
            a.next.size[i] = sz;
            a.next.stored[i] = M.stored[r];
            if (M.balive[M.band[r]]) {
                a.next.culture[i] = M.cult[r];
                if (e1 > e0) a.next.last_harvest[i] = M.lh[r];
            }
            pop += sz;
        }
        pop += __shfl_xor_sync(DSIM_FULL, pop, 16);
        pop += __shfl_xor_sync(DSIM_FULL, pop, 8);
        pop += __shfl_xor_sync(DSIM_FULL, pop, 4);
        pop += __shfl_xor_sync(DSIM_FULL, pop, 2);
        pop += __shfl_xor_sync(DSIM_FULL, pop, 1);
        if (lane == 0) {
            PatchView v;
            v.quality = quality;
            v.pop = pop;
            v.nb = nb_alive;
            v.cult0 = cult0;
            v.stor_off = s0;
            v.pad_ = 0;
            a.ps.view[p] = v;
            a.ps.cnt[p] = 0u;
        }
        __syncwarp();
    }
    if (errors) atomicOr(&ctl->errors, errors);
}

__global__ void k_advance(StepArgs a) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    DevCtl *ctl = a.ctl;
    const uint32_t ft = ctl->free_top, births = ctl->n_children, dead = ctl->n_dead;
    const uint32_t popped = births < ft ? births : ft;
    ctl->hs_hwm += births - popped;
    ctl->free_top = ft - popped + dead;
    if (ctl->hs_hwm > ctl->hs_cap) ctl->errors |= DSIM_MODEL_CAPACITY;
    ctl->t += 1u; /* s.t += 1, src/cli.rs:154 */
}

/* ------------------------------------------------------------------------------------------ *
 * census for freshly loaded states: view.pop and view.quality (lib.rs:522-525, 989-994)
 * ------------------------------------------------------------------------------------------ */
__global__ void k_init_views(PatchState ps, uint32_t n_nodes, double recovery) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t p = tid; p < n_nodes; p += nth) {
        double q = 0.0;
        for (uint32_t e = ps.eco_off[p]; e < ps.eco_off[p + 1]; ++e) q = q + (ps.res[e] + ps.max_res[e] * recovery);
        ps.view[p].quality = q;
    }
}

__global__ void k_clear_census(PatchState ps, uint32_t n_nodes) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t p = tid; p < n_nodes; p += nth) {
        ps.view[p].pop = 0u;
        ps.cnt[p] = 0u;
    }
}

__global__ void k_census(PatchState ps, FamCore f, uint32_t n, uint32_t *occ, DevCtl *ctl, uint32_t parity) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < n; i += nth) {
        const uint32_t p = f.loc[i];
        atomicAdd(&ps.view[p].pop, f.size[i]);
        /* remember the node as "occupied last step" so that the first step clears stale censuses */
        if (atomicAdd(&ps.cnt[p], 1u) == 0u) occ[atomicAdd(&ctl->n_occ[parity], 1u)] = p;
    }
}

__global__ void k_reset_cnt(PatchState ps, const uint32_t *occ, const DevCtl *ctl, uint32_t parity) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t n = ctl->n_occ[parity];
    for (uint32_t k = tid; k < n; k += nth) ps.cnt[occ[k]] = 0u;
}

/* ------------------------------------------------------------------------------------------ *
 * host-side launchers
 * ------------------------------------------------------------------------------------------ */
static uint32_t grid_for(const LaunchGeom &g, uint32_t per_sm) { return g.n_sm * per_sm; }

void dsim_launch_kernel(int which, const StepArgs &a, const LaunchGeom &g, cudaStream_t s) {
    switch (which) {
    case DSIM_K_LIFECYCLE: DSIM_LAUNCH(k_lifecycle, grid_for(g, 8), LIFE_BLOCK, 0, s, a); break;
    case DSIM_K_CONTROL: DSIM_LAUNCH(k_control, 1, CTL_THREADS, 0, s, a); break;
    case DSIM_K_MOVE: DSIM_LAUNCH(k_move, grid_for(g, 8), MOVE_THREADS, 0, s, a); break;
    case DSIM_K_SEGMENTS: DSIM_LAUNCH(k_segments, grid_for(g, 4), 256, 0, s, a); break;
    case DSIM_K_SCATTER: DSIM_LAUNCH(k_scatter, grid_for(g, 8), 256, 0, s, a); break;
    case DSIM_K_PATCH: DSIM_LAUNCH(k_patch, grid_for(g, 8), PATCH_WARPS * 32, 0, s, a); break;
    case DSIM_K_ADVANCE: DSIM_LAUNCH(k_advance, 1, 32, 0, s, a); break;
    default: break;
    }
}

void dsim_launch_init_views(const PatchState &ps, uint32_t n_nodes, double recovery, const LaunchGeom &g, cudaStream_t s) {
    DSIM_LAUNCH(k_init_views, grid_for(g, 4), 256, 0, s, ps, n_nodes, recovery);
}

void dsim_launch_census(const PatchState &ps, const FamCore &f, uint32_t n, uint32_t n_nodes, uint32_t *occ, DevCtl *ctl,
                        uint32_t parity, const LaunchGeom &g, cudaStream_t s) {
    DSIM_LAUNCH(k_clear_census, grid_for(g, 4), 256, 0, s, ps, n_nodes);
    DSIM_LAUNCH(k_census, grid_for(g, 4), 256, 0, s, ps, f, n, occ, ctl, parity);
    DSIM_LAUNCH(k_reset_cnt, grid_for(g, 4), 256, 0, s, ps, (const uint32_t *)occ, (const DevCtl *)ctl, parity);
}
