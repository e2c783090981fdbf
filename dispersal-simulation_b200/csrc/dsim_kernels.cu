/*
 * dsim_kernels.cu — the step loop of the dispersal model as sm_100a kernels.
 *
 * One season = `step` (lib.rs:477-496) = nine launches (k_children next to k_move, the four patch kernels concurrent), replayed as a CUDA graph:
 *
 *   part 1, update_each_family_step (lib.rs:511-552)
 *     k_lifecycle   consume / shrink / grow per family (lib.rs:527-539, 1912-1937); survivor and
 *                   split bitmaps, per-word and per-256 counts (the `retain` of lib.rs:541 becomes a scan)
 *                   + (last block to finish) exclusive scan of the per-256 counts, step bookkeeping
 *     k_move        one warp per tile of 4 families: decide_on_moving + best_location (lib.rs:830-929),
 *                   destination_quality (lib.rs:1034-1062), move bookkeeping (lib.rs:1826-1833), the
 *                   parent's half of maybe_procreate (lib.rs:1869-1897); writes the compacted next
 *                   family buffer; counts arrivals per node (histogram of the counting sort by patch)
 *     k_children    one warp per child: its record, history and adaptation, its own move
 *                   (lib.rs:1840-1855, 1877-1895); launched next to k_move, it picks a child up as soon as the parent's
 *                   tile has tagged its ChildTask
 *   part 2, distribute_each_patch_resources_step (lib.rs:591-655)
 *     k_segments    one thread per occupied node: claim a segment of `members`
 *     k_scatter     one thread per family: counting-sort scatter; clears the views of vacated nodes
 *     k_patch_one   a thread per node with a single family, k_patch_small eight lanes per node with 2..8 families,
 *     k_patch_mid   a warp per node with 9..32, k_patch a block per crowded node (the four run concurrently): cooperate_or_fight (lib.rs:1372-1417), adjust_culture / mutate
 *                   (lib.rs:1754-1785, 1960-1970), exploit_patch + recover (lib.rs:2021-2109), next step's
 *                   census, quality and band cultures (lib.rs:522-525, 989-994, 622-633)
 *     (the heavy-slot allocator bookkeeping and t += 1, src/cli.rs:154, are marked as due by k_scatter and applied by
 *     the tail of the next k_lifecycle or by the host's read of the control block: dsim_apply_advance)
 *
 * Floating point: doubles, IEEE operations only, compiled with -fmad=false so that every decision
 * (`g > m`, `res < max`, `stored - size*season < 0`) matches the CPU oracle bit for bit.
 */
#include "dsim_async.h"
#include "dsim_kernels.h"
#include "dsim_math.h"
#include "dsim_mg_sync.h"

#include <cstring>

#define LIFE_BLOCK 256
#ifndef LIFE_GRID_PER_SM
#define LIFE_GRID_PER_SM 8 /* every block ends with a ticket on one counter */
#endif
#ifndef CHILD_HR
#define CHILD_HR 3u /* k_children: rounds of 32 history entries handled together (3 x 32 >= the default child history of 91) */
#endif
#ifndef PATCH_WARPS
#define PATCH_WARPS 4 /* k_patch: one block of four warps per node */
#endif
#ifndef PATCH_MIN_BLOCKS
#define PATCH_MIN_BLOCKS 4
#endif
#ifndef PATCH_M
#define PATCH_M 512u /* families per node handled in shared memory (54 KB); larger nodes spill to HBM scratch */
#endif
#define PATCH_SMALL_MAX 8 /* nodes with at most this many families are handled by a group of 8 lanes each */

__device__ __forceinline__ dsim_stream stream_of(const ModelConst &mc) {
    dsim_stream s;
    s.k0 = mc.k0;
    s.k1 = mc.k1;
    return s;
}

/* resource_uncertainty (lib.rs:1317) for the candidate with noise index `i` of family `fid`: two draws per
 * Philox block; indices i and i + 32 of every run of 64 share a block (a lane of k_move takes both):
 * block = (i / 64) * 32 + i % 32, words (x,y) if i % 64 < 32, (z,w) otherwise. */
__device__ __forceinline__ uint32_t noise_block(uint32_t i) { return ((i >> 6) << 5) | (i & 31u); }
__device__ __forceinline__ double noise_pick(const dsim_u4 &o, uint32_t i) {
    return (i & 32u) ? dsim_u53(o.z, o.w) : dsim_u53(o.x, o.y);
}
__device__ __forceinline__ double noise_draw(const ModelConst &mc, uint32_t step, uint64_t fid, uint32_t i) {
    const dsim_u4 o = dsim_draw(stream_of(mc), step, fid, DSIM_P_NOISE, noise_block(i));
    return noise_pick(o, i);
}

/* ------------------------------------------------------------------------------------------ *
 * part 1a: lifecycle
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void control_tail(const StepArgs &a, uint32_t n, uint32_t nvb);

__global__ void __launch_bounds__(LIFE_BLOCK) k_lifecycle(StepArgs a) {
    dsim_pdl_wait();
    const uint32_t n = a.ctl->n_cur[a.parity];
    const uint32_t nvb = (n + LIFE_BLOCK - 1) / LIFE_BLOCK;
    __shared__ uint32_t s_alive[LIFE_BLOCK / 32], s_split[LIFE_BLOCK / 32];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const double season = a.mc.season;
    for (uint32_t vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
        const uint32_t i = vb * LIFE_BLOCK + threadIdx.x;
        bool alive = false, split = false;
        if (i < n) {
            uint32_t size = a.cur.place[i].size;
            double stored = a.cur.stock[i].stored;
            uint32_t till_adult = a.cur.clock[i].till_adult;
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
            a.cur.place[i].size = size;
            a.cur.stock[i].stored = stored;
            a.cur.clock[i].till_adult = till_adult;
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
        if (threadIdx.x <= LIFE_BLOCK / 32) { /* thread k: counts of the words before word k; thread 8: the group's totals */
            uint32_t sa = 0, ss = 0;
            for (uint32_t k = 0; k < threadIdx.x; ++k) {
                sa += s_alive[k];
                ss += s_split[k];
            }
            if (threadIdx.x < LIFE_BLOCK / 32) {
                a.sc.word_pre[vb * (LIFE_BLOCK / 32) + threadIdx.x] = sa | (ss << 16);
            } else {
                a.sc.blk_alive[vb] = sa;
                a.sc.blk_split[vb] = ss;
            }
        }
        __syncthreads();
    }
    /* the last block to get here does part 1b */
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&a.ctl->life_done, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    control_tail(a, n, nvb);
}

/* ------------------------------------------------------------------------------------------ *
 * part 1b: exclusive scan of the per-256 counts + bookkeeping, by the last block of k_lifecycle to finish
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void control_tail(const StepArgs &a, uint32_t n, uint32_t nvb) {
    __shared__ uint32_t w_a[LIFE_BLOCK / 32], w_s[LIFE_BLOCK / 32];
    DevCtl *ctl = a.ctl;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    /* thread k owns the entries [k * per, (k + 1) * per) */
    const uint32_t per = (nvb + LIFE_BLOCK - 1) / LIFE_BLOCK;
    const uint32_t k0 = tid * per < nvb ? tid * per : nvb, k1 = k0 + per < nvb ? k0 + per : nvb;
    uint32_t ta = 0, ts = 0;
    for (uint32_t k = k0; k < k1; ++k) {
        ta += DSIM_LDCG(a.sc.blk_alive + k);
        ts += DSIM_LDCG(a.sc.blk_split + k);
    }
    uint32_t ia = ta, is = ts; /* inclusive scan over the block */
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t oa = __shfl_up_sync(DSIM_FULL, ia, d), os = __shfl_up_sync(DSIM_FULL, is, d);
        if ((int)lane >= d) {
            ia += oa;
            is += os;
        }
    }
    if (lane == 31u) {
        w_a[w] = ia;
        w_s[w] = is;
    }
    __syncthreads();
    uint32_t ba = 0, bs = 0, tot_a = 0, tot_s = 0;
    for (uint32_t k = 0; k < LIFE_BLOCK / 32; ++k) {
        if (k < w) {
            ba += w_a[k];
            bs += w_s[k];
        }
        tot_a += w_a[k];
        tot_s += w_s[k];
    }
    uint32_t ea = ba + ia - ta, es = bs + is - ts; /* exclusive prefix of this thread's first entry */
    for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t va = DSIM_LDCG(a.sc.blk_alive + k), vs = DSIM_LDCG(a.sc.blk_split + k);
        a.sc.blk_alive[k] = ea;
        a.sc.blk_split[k] = es;
        ea += va;
        es += vs;
    }
    if (tid == 0) {
        dsim_apply_advance(ctl); /* the end of the previous step: before its n_children / n_dead are replaced */
        const uint32_t carry_a = tot_a, carry_s = tot_s;
        const uint32_t n_alive = carry_a, n_children = carry_s;
        ctl->n_alive = n_alive;
        ctl->n_children = n_children;
        ctl->n_dead = n - n_alive;
        ctl->n_next = n_alive + n_children;
        ctl->child_pos_base = n_alive;
        ctl->child_id_base = ctl->next_id;
        ctl->next_id += n_children;
        ctl->n_cur[a.parity ^ 1u] = n_alive + n_children;
        ctl->n_occ[a.occ_par] = 0;
        ctl->seg_cursor = 0;
        ctl->n_small = 0;
        ctl->n_one = 0;
        ctl->n_mid = 0;
        ctl->n_big = 0;
        ctl->n_huge = 0;
        ctl->n_over = 0;
        ctl->births += n_children;
        ctl->deaths += n - n_alive;
        ctl->updates += n;
        if (n_alive + n_children > ctl->cap) {
            ctl->errors |= DSIM_MODEL_CAPACITY;
            ctl->n_next = ctl->cap;
            ctl->n_cur[a.parity ^ 1u] = ctl->cap;
        }
        ctl->life_done = 0;
        ctl->move_next = 0; /* k_move hands its tiles out from here */
        ctl->serial += 1u;  /* tags this step's ChildTasks */
    }
}

/* ------------------------------------------------------------------------------------------ *
 * part 1c: move + split — shared pieces
 * ------------------------------------------------------------------------------------------ */
struct Choice { /* best_location state, lib.rs:911-916 */
    uint32_t target;
    double max_gain, t_cost, m;
};

#define DSIM_NEG_INF (dsim_bits_to_f64(0xFFF0000000000000ull))

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *v, uint32_t lo, uint32_t hi, uint32_t q) {
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (v[mid] < q)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* destination_quality, lib.rs:1034-1062 (has_enemies lib.rs:996-1017, expected_harvest 989-994), in two
 * steps so that a caller can request the band cultures of several candidates before it uses any. */
struct BandsRef {
    uint64_t cult0;    /* first band culture stored for the node */
    uint32_t stor_off; /* where all of them are */
};
__device__ __forceinline__ BandsRef bands_fetch(const StepArgs &a, const PatchView &v, uint32_t q) {
    BandsRef r;
    r.cult0 = 0;
    r.stor_off = 0;
    if (v.nb > 0u) { /* candidates occupied last step */
        const PatchBands b = a.ps.bands[q];
        r.cult0 = b.cult0;
        r.stor_off = b.stor_off;
    }
    return r;
}
__device__ __forceinline__ double quality_eval(const StepArgs &a, const PatchView &v, const BandsRef &b, uint64_t culture, uint32_t size,
                                               bool same_patch) {
    bool enemy = false;
    if (v.nb > 0u) {
        enemy = !dsim_will_cooperate(culture, b.cult0, a.mc.threshold, a.mc.inclusive);
        if (v.nb > 1u && !enemy) {
            const uint64_t *others = (b.stor_off & DSIM_STOR_HALO) ? a.mg.halo_cult[(b.stor_off >> 30) & 1u] + (b.stor_off & 0x3FFFFFFFu)
                                                                    : a.ps.band_cult + b.stor_off;
            for (uint32_t k = 1; k < v.nb && !enemy; ++k)
                enemy = !dsim_will_cooperate(culture, others[k], a.mc.threshold, a.mc.inclusive);
        }
    }
    if (enemy) return 0.0;
    const uint64_t pop_at_destination = (uint64_t)v.pop + (same_patch ? 0u : size);
    return v.quality / (double)pop_at_destination;
}
__device__ __forceinline__ double view_quality(const StepArgs &a, const PatchView &v, uint32_t q, uint64_t culture, uint32_t size,
                                               bool same_patch) {
    const BandsRef b = bands_fetch(a, v, q);
    return quality_eval(a, v, b, culture, size, same_patch);
}

/* ---- remembered-patch sampling, lib.rs:854-866 (canonical stream v2, ModelConst) ---------------------------
 * Head: history entry n < mem_tail0 is kept iff U_n < memory_n with the 40-bit uniform U_n = (B_n 2^32 + W_n) 2^-40:
 * one Philox block of the MEM stream holds the top bytes B_n of 16 consecutive entries; they decide all but 1/256 of
 * the entries, the others need the word W_n of the MEM2 stream.  The byte comparisons are done four at a time on
 * packed words (flags in bit 7 of every byte).
 * Tail: the entries n >= mem_tail0 (memory_n < 2^-8, 0.07 kept entries expected out of 500) are sampled by skipping:
 * one uniform finds the next kept entry in the table of survival products, so a family costs one Philox block for its
 * whole tail instead of one per 16 entries. */
__device__ __forceinline__ uint32_t swar_ltu4(uint32_t x, uint32_t y) { /* bytes of x < bytes of y, unsigned */
    const uint32_t d = (x | 0x80808080u) - (y & 0x7F7F7F7Fu); /* bit 7: low 7 bits of x >= low 7 bits of y; no borrows */
    return ((~x & y) | (~(x ^ y) & ~d)) & 0x80808080u;
}
__device__ __forceinline__ uint32_t swar_eq4(uint32_t x, uint32_t y) { /* bytes of x == bytes of y */
    const uint32_t z = x ^ y;
    return ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);
}
__device__ __forceinline__ uint32_t swar_pack4(uint32_t m) { return ((m >> 7) * 0x10204080u) >> 28; } /* bit 7 of byte k -> bit k */

__device__ __forceinline__ bool mem_resolve(const ModelConst &mc, uint32_t step, uint64_t fid, uint32_t n) {
    const dsim_u4 o = dsim_draw(stream_of(mc), step, fid, DSIM_P_MEM2, n >> 2);
    return dsim_word(o, n & 3u) < mc.mem_lo32[n];
}
/* head entries [16 jb, 16 jb + 16) of family `fid` (jb < mem_head, 16 jb < nh): the mask of the kept ones */
__device__ __forceinline__ uint32_t mem_block(const ModelConst &mc, uint32_t step, uint64_t fid, uint32_t jb, uint32_t nh) {
    const dsim_u4 o = dsim_draw(stream_of(mc), step, fid, DSIM_P_MEM, jb);
    const uint4 h = *reinterpret_cast<const uint4 *>(mc.mem_hi8 + 16u * jb);
    uint32_t lt = swar_pack4(swar_ltu4(o.x, h.x)) | (swar_pack4(swar_ltu4(o.y, h.y)) << 4) | (swar_pack4(swar_ltu4(o.z, h.z)) << 8) |
                  (swar_pack4(swar_ltu4(o.w, h.w)) << 12);
    const uint32_t eq = swar_pack4(swar_eq4(o.x, h.x)) | (swar_pack4(swar_eq4(o.y, h.y)) << 4) | (swar_pack4(swar_eq4(o.z, h.z)) << 8) |
                        (swar_pack4(swar_eq4(o.w, h.w)) << 12);
    const uint32_t n0 = 16u * jb;
    if (n0 < mc.mem_always) lt |= (mc.mem_always - n0 >= 16u) ? 0xFFFFu : ((1u << (mc.mem_always - n0)) - 1u);
    const uint32_t valid = (nh - n0 >= 16u) ? 0xFFFFu : ((1u << (nh - n0)) - 1u);
    lt &= valid;
    uint32_t undecided = eq & valid & ~lt; /* top byte equal to the threshold byte (1/256 per entry): the low word decides */
    while (undecided) {
        const uint32_t b = (uint32_t)__ffs((int)undecided) - 1u;
        undecided &= undecided - 1u;
        if (mem_resolve(mc, step, fid, n0 + b)) lt |= 1u << b;
    }
    return lt;
}
/* The tail of a family with L tail entries, from the state (j0, base, d) = (first undecided tail entry, survival
 * product before it, uniforms used so far): appends kept entries (absolute indices) to out[0 .. cap) and returns
 * their number; *more = the list filled up before the tail was exhausted (call again with the updated state). */
struct TailState {
    uint32_t j0, d;
    double base;
};
__device__ __forceinline__ uint32_t mem_tail_run(const ModelConst &mc, uint32_t step, uint64_t fid, uint32_t L, TailState &ts, uint32_t *out,
                                                 uint32_t cap, bool *more) {
    uint32_t n_out = 0;
    *more = false;
    while (ts.j0 < L) {
        if (n_out == cap) {
            *more = true;
            break;
        }
        const dsim_u4 o = dsim_draw(stream_of(mc), step, fid, DSIM_P_MEMT, ts.d >> 1);
        const double u = (ts.d & 1u) ? dsim_u53(o.z, o.w) : dsim_u53(o.x, o.y);
        const double thr = u * ts.base;
        if (mc.mem_tail[L - 1u] > thr) { /* no entry of the rest of the tail is kept (93 % of the first draws) */
            ts.j0 = L;
            break;
        }
        uint32_t lo = ts.j0, hi = L - 1u; /* smallest j >= j0 with mem_tail[j] <= thr; the table decreases */
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (mc.mem_tail[mid] <= thr)
                hi = mid;
            else
                lo = mid + 1u;
        }
        out[n_out++] = mc.mem_tail0 + lo;
        ts.base = mc.mem_tail[lo];
        ts.j0 = lo + 1u;
        ts.d += 1u;
    }
    return n_out;
}
/* one entry n < nh of a history of nh sampled entries (k_children: a lane per entry) */
__device__ __forceinline__ bool mem_keep(const ModelConst &mc, uint32_t step, uint64_t fid, uint32_t n, uint32_t nh) {
    if (n < mc.mem_always) return true;
    if (n < mc.mem_tail0) {
        const dsim_u4 o = dsim_draw(stream_of(mc), step, fid, DSIM_P_MEM, n >> 4);
        const uint32_t b = (dsim_word(o, (n >> 2) & 3u) >> (8u * (n & 3u))) & 0xFFu, h = mc.mem_hi8[n];
        if (b != h) return b < h;
        return mem_resolve(mc, step, fid, n);
    }
    TailState ts; /* replay the skipping process up to entry n */
    ts.j0 = 0u;
    ts.d = 0u;
    ts.base = 1.0;
    const uint32_t L = nh - mc.mem_tail0;
    for (;;) {
        uint32_t got;
        bool more;
        if (mem_tail_run(mc, step, fid, L, ts, &got, 1u, &more) == 0u) return false;
        if (got >= n) return got == n;
    }
}

/* ---- reach rows in place ------------------------------------------------------------------------- */
struct RowPtr { /* a reach row in place (HBM / L1) */
    const uint32_t *dst;
    const double *cost;
    const uint16_t *hash; /* the row's 32 x 8 bucket index, or nullptr for rows longer than DSIM_ROW_HASH_MAX */
    uint32_t len, n_sensed;
};

__device__ __forceinline__ RowPtr row_of(const ReachTable &rt, uint32_t node, const RowHdr &hd) {
    RowPtr row;
    row.dst = rt.dst + hd.start;
    row.cost = rt.cost + hd.start;
    row.hash = hd.len <= DSIM_ROW_HASH_MAX ? rt.hash + (size_t)node * 256u : nullptr;
    row.len = hd.len;
    row.n_sensed = hd.n_sensed;
    return row;
}

__device__ __forceinline__ uint32_t row_hash(uint32_t q) { return (q * 2654435761u) >> 19; } /* bucket = bits 8..12, tag = bits 0..7 */

/* position of node q in the row, or 0xFFFFFFFF (replaces travel_costs.get(), lib.rs:869), by binary search in
 * the two sorted segments of the row */
__device__ DSIM_NOINLINE uint32_t row_find_sorted(const uint32_t *dst, uint32_t n_sensed, uint32_t len, uint32_t q) {
    uint32_t k = lower_bound_u32(dst, 0u, n_sensed, q);
    if (k < n_sensed && dst[k] == q) return k;
    k = lower_bound_u32(dst, n_sensed, len, q);
    if (k < len && dst[k] == q) return k;
    return 0xFFFFFFFFu;
}

/* The row's bucket index (rows of at most DSIM_ROW_HASH_MAX entries): one 16-byte load reads the key's home
 * bucket of 8 slots (tag << 8 | position, 0xFFFF = empty).  bucket_probe returns the position of the first
 * slot with the key's tag (to be verified against the row), ROW_ABSENT when no slot has it and the bucket
 * has a free slot, ROW_SEARCH otherwise (full bucket: the key may have overflowed into the next one).
 * A failed verification (tag collision, 1/256 per slot) also falls back to the binary search. */
#define ROW_ABSENT 0xFFFFFFFFu
#define ROW_SEARCH 0xFFFFFFFEu
__device__ __forceinline__ const uint4 *bucket_of(const uint16_t *hash, uint32_t q) {
    return reinterpret_cast<const uint4 *>(hash + (row_hash(q) >> 8) * 8u);
}
__device__ __forceinline__ uint32_t bucket_probe(const uint4 &bk, uint32_t q) {
    const uint32_t t2 = (row_hash(q) & 0xFFu) * 0x01000100u; /* the tag in the high byte of both 16-bit slots of a word */
    const uint32_t f0 = swar_eq4(bk.x & 0xFF00FF00u, t2) & 0x80008000u, f1 = swar_eq4(bk.y & 0xFF00FF00u, t2) & 0x80008000u;
    const uint32_t f2 = swar_eq4(bk.z & 0xFF00FF00u, t2) & 0x80008000u, f3 = swar_eq4(bk.w & 0xFF00FF00u, t2) & 0x80008000u;
    if ((f0 | f1 | f2 | f3) != 0u) {
        /* first matching slot in slot order: low half of x, high half of x, low half of y, ... */
        const uint32_t w = f0 ? bk.x : f1 ? bk.y : f2 ? bk.z : bk.w;
        const uint32_t f = f0 ? f0 : f1 ? f1 : f2 ? f2 : f3;
        return (f & 0x8000u) ? (w & 0xFFu) : ((w >> 16) & 0xFFu);
    }
    const uint32_t e0 = ~bk.x, e1 = ~bk.y, e2 = ~bk.z, e3 = ~bk.w; /* an empty slot becomes a zero halfword */
    const uint32_t z = ((e0 - 0x00010001u) & ~e0) | ((e1 - 0x00010001u) & ~e1) | ((e2 - 0x00010001u) & ~e2) | ((e3 - 0x00010001u) & ~e3);
    return (z & 0x80008000u) != 0u ? ROW_ABSENT : ROW_SEARCH;
}

/* position of node q in the row, or ROW_ABSENT (replaces travel_costs.get(), lib.rs:869) */
__device__ __forceinline__ uint32_t row_find(const RowPtr &row, uint32_t q) {
    if (row.hash != nullptr) {
        const uint4 bk = *bucket_of(row.hash, q);
        const uint32_t pos = bucket_probe(bk, q);
        if (pos == ROW_ABSENT) return ROW_ABSENT;
        if (pos < row.len && row.dst[pos] == q) return pos;
    }
    return row_find_sorted(row.dst, row.n_sensed, row.len, q);
}

/* Ordered acceptance scan of lib.rs:918-927 over the (up to) 32 candidates held one per lane, in lane
 * order: a candidate is accepted iff g > m with m as left by the previous acceptance.  `ch` is
 * warp-uniform.  An acceptance raises m to at least the accepted g (m' = max(max_gain, gain) - cost >=
 * gain U - cost), so the candidates before the accepted lane, which failed against the smaller m, and
 * the accepted lane itself drop out of the next ballot without a cursor. */
__device__ __forceinline__ void accept_scan_warp(Choice &ch, bool valid, double g, double gain, double cost, uint32_t q) {
    for (;;) {
        const uint32_t mask = __ballot_sync(DSIM_FULL, valid && g > ch.m);
        if (mask == 0u) break;
        const int L = __ffs((int)mask) - 1;
        const double m_before = ch.m;
        ch.target = __shfl_sync(DSIM_FULL, q, L);
        ch.max_gain = dsim_oyr_max(ch.max_gain, __shfl_sync(DSIM_FULL, gain, L));
        ch.t_cost = __shfl_sync(DSIM_FULL, cost, L);
        ch.m = ch.max_gain - ch.t_cost;
        if (!(ch.m > m_before)) break; /* cannot happen for finite non-negative gains; guards against a livelock on NaN/inf input */
    }
}

/* several GPUs: a family whose new node belongs to a neighbour's band is listed for k_mg_pack right where it is written */
__device__ __forceinline__ void mg_list_emigrant(const StepArgs &a, uint32_t pos, uint32_t newloc) {
    const uint32_t side = newloc < a.mg.own_lo ? 0u : 1u;
    const uint32_t e = atomicAdd(&a.mg.counters[MG_N_EMIG0 + side], 1u);
    if (e < a.mg.ecap)
        a.mg.emig_list[side * a.mg.ecap + e] = pos;
    else
        atomicOr(&a.ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
}

/* ------------------------------------------------------------------------------------------ *
 * part 1c: move, one WARP per tile of MOVE_WF families, work items spread over the lanes.
 *
 * The decision of one family is a short ORDERED scan (lib.rs:918-927) over ~60 sensed patches and ~18
 * remembered patches, but everything that feeds the scan is independent per candidate: 8 Philox blocks
 * of memory draws, one Philox block per two sensed patches plus a 16-byte view gather each, and per
 * remembered patch a ring read, a reach-row look-up and one Philox block.  Measured on B200 (profiles/):
 * a thread per family serialises all of that behind its own memory round trips (330 us at 85 k families);
 * a block per 32 families with block-wide phases idles at its barriers (270 us); a warp per tile that
 * walks its families one after the other is bound by the ~28 dependent memory round trips of a tile
 * (129 us, round 1).  Here a warp owns MOVE_WF families from their records to their new records, needs
 * no block barrier, and keeps the memory system busy for the whole tile at once:
 *   - the reach rows of all the tile's families (destinations and costs of the sensed patches) are
 *     fetched by the TMA engine — one bulk copy per row column, completion on an mbarrier — while the
 *     warp computes the memory draws: (family, 16 entries) items and the tails dealt to the lanes;
 *   - the candidate views of all families are then gathered with 16-byte cp.async into shared memory
 *     while the warp lists the kept history entries;
 *   - sensed patches: 64 per pass from shared memory (lane L takes entries L and L+32, which share one
 *     Philox block), the ordered acceptance a ballot loop with one iteration per acceptance;
 *   - remembered patches: the kept entries of ALL the tile's families form one list, 32 per pass whatever
 *     family they belong to; the acceptance runs per family present in the pass.
 * Tiles are handed out through an atomic counter (no tail of idle warps at 4 tiles per warp).
 * ------------------------------------------------------------------------------------------ */
#ifndef MOVE_WF
#define MOVE_WF 4u        /* families per warp tile (a power of two, 2 ... 8) */
#endif
#ifndef MOVE_LCAP
#define MOVE_LCAP 128u    /* kept history entries of a tile listed at a time (expected MOVE_WF * 18) */
#endif
#ifndef MOVE_TAILQ
#define MOVE_TAILQ 4u     /* kept tail entries per family listed at a time (expected 0.07) */
#endif
#define MOVE_SSTAGE 64u   /* sensed entries per family staged in shared memory: the first pass */
#define MOVE_HB_MAX 16u   /* head blocks of 16 history entries (ModelConst.mem_head <= 16) */
#ifndef MOVE_WARPS
#define MOVE_WARPS 4u
#endif
#ifndef MOVE_MIN_BLOCKS
#define MOVE_MIN_BLOCKS 6 /* 8.8 KB of shared memory per warp: 24 warps per SM, 80 registers per thread */
#endif

struct __align__(16) MoveTile { /* per warp */
    /* filled by asynchronous copies */
    PatchView view[MOVE_WF][MOVE_SSTAGE];  /* candidate views of the first MOVE_SSTAGE sensed patches (cp.async) */
    double r_cost[MOVE_WF][MOVE_SSTAGE];   /* their travel costs and node ids: the head of the reach row (TMA bulk copy) */
    uint32_t r_dst[MOVE_WF][MOVE_SSTAGE];
    uint64_t fid[MOVE_WF], culture[MOVE_WF];
    double max_gain[MOVE_WF], t_cost[MOVE_WF], m[MOVE_WF]; /* best_location state of each family (lib.rs:911-916) */
    double stored[MOVE_WF], last_harvest[MOVE_WF], tbase[MOVE_WF];
    uint32_t target[MOVE_WF];
    uint32_t loc[MOVE_WF], size[MOVE_WF], hs[MOVE_WF], hist_cnt[MOVE_WF], nh[MOVE_WF];
    uint32_t row_start[MOVE_WF], row_len[MOVE_WF], ns[MOVE_WF], n_near[MOVE_WF];
    uint32_t idx[MOVE_WF], newpos[MOVE_WF], crank[MOVE_WF], misc[MOVE_WF], newest[MOVE_WF];
    uint32_t till_adult[MOVE_WF], till_mut[MOVE_WF];
    uint32_t tj0[MOVE_WF], td[MOVE_WF], ntail[MOVE_WF], tmore[MOVE_WF]; /* tail sampling state (mem_tail_run) */
    uint32_t src[MOVE_WF];                 /* the family of the tile whose staged row and views this one uses (same node) */
    uint32_t tail[MOVE_WF][MOVE_TAILQ];    /* kept tail entries */
    uint32_t list[MOVE_LCAP];              /* kept entries of the tile: family << 24 | entry, family-major, ascending */
    uint16_t hmask[MOVE_WF][MOVE_HB_MAX];  /* kept head entries, 16 per block */
    dsim_mbar bar;                         /* completion of the bulk copies of a tile */
};

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(DSIM_FULL, v, (unsigned)d);
        if ((int)lane >= d) v += o;
    }
    return v;
}

__global__ void __launch_bounds__(MOVE_WARPS * 32, MOVE_MIN_BLOCKS) k_move(StepArgs a) {
    __shared__ MoveTile tiles[MOVE_WARPS];
    MoveTile &S = tiles[threadIdx.x >> 5];
    DevCtl *ctl = a.ctl;
    dsim_pdl_wait();
    const uint32_t n = ctl->n_cur[a.parity];
    const uint32_t step = ctl->t;
    const uint32_t serial = ctl->serial;
    const uint32_t cap = ctl->cap;
    const uint32_t hcap = a.hv.hcap;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ntiles = (n + MOVE_WF - 1u) / MOVE_WF;
    const dsim_stream rs = stream_of(a.mc);
    const uint32_t HB = a.mc.mem_head;          /* head blocks; item j == HB of a family is its tail */
    const uint32_t n_items = MOVE_WF * (HB + 1u);
    uint32_t phase = 0;                         /* parity of the mbarrier phase the next wait completes */
    uint32_t pend_pos = 0xFFFFFFFFu, pend_slot = 0; /* arrival rank of this lane's last family, not yet written */
    /* staging of the remembered-patch look-ups, in the arrays of the sensed phase (dead by then): ring entries (node,
     * harvest), index buckets | row entries (cost, node) | noise draws */
    static_assert(28u * MOVE_LCAP <= sizeof(S.view) && 12u * MOVE_LCAP <= sizeof(S.r_cost) && 8u * MOVE_LCAP <= sizeof(S.r_dst) &&
                      MOVE_LCAP % 32u == 0u,
                  "MOVE_LCAP entries must fit the staging arrays");
    unsigned char *const st_view = reinterpret_cast<unsigned char *>(&S.view[0][0]);
    uint32_t *const rq = reinterpret_cast<uint32_t *>(st_view);
    double *const rgain = reinterpret_cast<double *>(st_view + 4u * MOVE_LCAP);
    uint4 *const rbk = reinterpret_cast<uint4 *>(st_view + 12u * MOVE_LCAP);
    double *const rc = &S.r_cost[0][0];
    uint32_t *const rdq = reinterpret_cast<uint32_t *>(&S.r_cost[0][0] + MOVE_LCAP);
    double *const ru = reinterpret_cast<double *>(&S.r_dst[0][0]);

    if (lane == 0) dsim_mbar_init(&S.bar, 1u);
    __syncwarp();
#ifndef MOVE_HINTS
#define MOVE_HINTS 3
#endif
    const unsigned long long pol_keep = dsim_policy_evict_last(), pol_stream = dsim_policy_evict_first();
    (void)pol_keep;
    (void)pol_stream;

    /* tiles are handed out by a counter: the claim for the next tile is in flight while this one is worked on */
    uint32_t tile = 0;
    if (lane == 0) tile = atomicAdd(&ctl->move_next, 1u);
    tile = __shfl_sync(DSIM_FULL, tile, 0);
    while (tile < ntiles) {
        uint32_t next_tile = 0;
        if (lane == 0) next_tile = atomicAdd(&ctl->move_next, 1u);

        /* ---- prologue: lane f < MOVE_WF owns family f of the tile ---------------------------------- *
         * Every load that depends only on the family's position is requested before any is used (also for
         * families that turn out to be dead: their records are still valid memory), so the prologue costs
         * three memory round trips: members -> record and bitmaps -> row header and history length.
         * (Fetching the next tile's prologue with cp.async while this tile is worked on was measured and
         * brought nothing: the other resident warps already cover these round trips.) */
        uint32_t nh = 0, state = 0, nst = 0; /* state: bit 0 = decides (survivor), bit 1 = splits; nst: staged sensed entries, padded to 8 */
        if (lane < MOVE_WF) {
            /* Families are visited in the patch-grouped order that part 2 of the previous step left in
             * `members`: co-resident families (same reach row, same candidate views) share a tile or sit in
             * neighbouring ones.  The order never influences results. */
            const uint32_t t = tile * MOVE_WF + lane;
            uint32_t ns = 0;
            if (t < n) {
                const uint32_t i = a.sc.members[t];
                const uint32_t c = i >> 5, lt = (1u << (i & 31u)) - 1u;
                const uint32_t alive_mask = a.sc.alive_bits[c], split_mask = a.sc.split_bits[c];
                const uint32_t base_alive = a.sc.blk_alive[c >> 3], base_split = a.sc.blk_split[c >> 3];
                /* the 56-byte record: three 128-bit loads and one 64-bit load */
                const FamPlace pl = a.cur.place[i];
                const FamKey ky = a.cur.key[i];
                const FamStock sk = a.cur.stock[i];
                const FamClock ck = a.cur.clock[i];
                const uint32_t hs = pl.hs, loc = pl.loc, size = pl.size, misc = pl.misc;
                const uint64_t fid = ky.id, culture = ky.culture;
                const double stored = sk.stored, last_harvest = sk.last_harvest;
                const uint32_t till_adult = ck.till_adult, till_mut = ck.till_mut;
                /* ranks from the bitmaps (retain lib.rs:541, children lib.rs:551) */
                const uint32_t pre = a.sc.word_pre[c], va = pre & 0xFFFFu, vs = pre >> 16;
                const RowHdr hd = a.rt.hdr[loc];
                const uint32_t hist_cnt = a.hv.hist_cnt[hs];
                const uint32_t newpos = base_alive + va + (uint32_t)__popc(alive_mask & lt);
                if (!((alive_mask >> (i & 31u)) & 1u)) {
                    a.sc.dead_hs[i - newpos] = hs; /* dropped by `retain`, lib.rs:541: hand the heavy slot back */
                } else if (newpos < cap) {         /* else: capacity error already flagged by k_lifecycle's tail */
                    state = 1u | (((split_mask >> (i & 31u)) & 1u) << 1);
                    S.till_adult[lane] = till_adult;
                    S.till_mut[lane] = till_mut;
                    S.fid[lane] = fid;
                    S.culture[lane] = culture;
                    S.stored[lane] = stored;
                    S.last_harvest[lane] = last_harvest;
                    S.size[lane] = size;
                    S.misc[lane] = misc;
                    S.idx[lane] = i;
                    S.newpos[lane] = newpos;
                    S.crank[lane] = base_split + vs + (uint32_t)__popc(split_mask & lt);
                    const uint32_t hist_valid = hist_cnt < hcap ? hist_cnt : hcap;
                    nh = hist_valid < a.mc.n_mem ? hist_valid : a.mc.n_mem;
                    ns = hd.n_sensed;
                    nst = ((ns < MOVE_SSTAGE ? ns : MOVE_SSTAGE) + 7u) & ~7u; /* rows are padded to 8 entries */
                    S.loc[lane] = loc;
                    S.hs[lane] = hs;
                    S.hist_cnt[lane] = hist_cnt;
                    S.newest[lane] = (hist_cnt - 1u) % hcap; /* ring position of entry 0; entry n sits n positions before it */
                    S.row_start[lane] = hd.start;
                    S.row_len[lane] = hd.len;
                    S.n_near[lane] = hd.n_near;
                    /* best_location state, lib.rs:911-916 */
                    S.target[lane] = loc;
                    S.max_gain[lane] = 0.0;
                    S.t_cost[lane] = 0.0;
                    S.m[lane] = 0.0;
                }
            }
            S.ns[lane] = ns;
            S.nh[lane] = nh;
        }
        const uint32_t live_mask = __ballot_sync(DSIM_FULL, (state & 1u) != 0u);
        /* Co-resident families (the visiting order groups them) look at the same row and the same views: only the
         * first one of a tile stages them, the others read its copy. */
        {
            const uint32_t my_loc = (lane < MOVE_WF && (state & 1u)) ? S.loc[lane] : 0xFFFFFFFFu;
            uint32_t from = lane;
#pragma unroll
            for (uint32_t g = MOVE_WF - 1u; g-- > 0u;) { /* g = MOVE_WF-2 ... 0: the smallest g wins */
                const uint32_t lg = __shfl_sync(DSIM_FULL, my_loc, (int)g);
                if (g < lane && lg == my_loc && my_loc != 0xFFFFFFFFu) from = g;
            }
            if (lane < MOVE_WF) {
                S.src[lane] = from;
                if (from != lane) nst = 0u;
            }
        }
        /* ---- the reach rows of the tile, by bulk copy ---------------------------------------------- */
        uint32_t stage_entries = nst; /* sum over the tile's families */
#pragma unroll
        for (uint32_t d = 1; d < MOVE_WF; d <<= 1) stage_entries += __shfl_xor_sync(DSIM_FULL, stage_entries, (int)d);
        stage_entries = __shfl_sync(DSIM_FULL, stage_entries, 0);
        dsim_fence_proxy_async(); /* the previous tile's reads of the staging arrays come before this tile's copies into them */
        __syncwarp();
        if (stage_entries != 0u) {
            if (lane == 0) dsim_mbar_expect_tx(&S.bar, stage_entries * 12u);
            __syncwarp();
            const uint32_t f = lane >> 1;
            const uint32_t nst_f = __shfl_sync(DSIM_FULL, nst, (int)(f < MOVE_WF ? f : 0u));
            if (f < MOVE_WF && nst_f != 0u) {
#if MOVE_HINTS & 2
                if (lane & 1u)
                    dsim_bulk_g2s_hint(&S.r_cost[f][0], a.rt.cost + S.row_start[f], nst_f * 8u, &S.bar, pol_stream);
                else
                    dsim_bulk_g2s_hint(&S.r_dst[f][0], a.rt.dst + S.row_start[f], nst_f * 4u, &S.bar, pol_stream);
#else
                if (lane & 1u)
                    dsim_bulk_g2s(&S.r_cost[f][0], a.rt.cost + S.row_start[f], nst_f * 8u, &S.bar);
                else
                    dsim_bulk_g2s(&S.r_dst[f][0], a.rt.dst + S.row_start[f], nst_f * 4u, &S.bar);
#endif
            }
        }

        /* ---- memory draws, lib.rs:854-866: item (family f, block j) of the head or (f, HB) = the tail -- */
        for (uint32_t it = lane; it < n_items; it += 32u) {
            const uint32_t f = it / (HB + 1u), j = it - f * (HB + 1u);
            const uint32_t nhf = S.nh[f];
            if (j < HB) {
                uint32_t kept = 0;
                if (16u * j < nhf) kept = mem_block(a.mc, step, S.fid[f], j, nhf < a.mc.mem_tail0 ? nhf : a.mc.mem_tail0);
                S.hmask[f][j] = (uint16_t)kept;
            } else {
                TailState ts;
                ts.j0 = 0u;
                ts.d = 0u;
                ts.base = 1.0;
                uint32_t cnt = 0;
                bool more = false;
                if (nhf > a.mc.mem_tail0) cnt = mem_tail_run(a.mc, step, S.fid[f], nhf - a.mc.mem_tail0, ts, &S.tail[f][0], MOVE_TAILQ, &more);
                S.ntail[f] = cnt;
                S.tmore[f] = more ? 1u : 0u;
                S.tj0[f] = ts.j0;
                S.td[f] = ts.d;
                S.tbase[f] = ts.base;
            }
        }

        /* ---- the candidate views of the staged sensed patches: 16-byte gathers into shared memory --- */
        if (stage_entries != 0u) {
            dsim_mbar_wait(&S.bar, phase);
            phase ^= 1u;
#pragma unroll
            for (uint32_t f = 0; f < MOVE_WF; ++f) {
                if (!((live_mask >> f) & 1u) || S.src[f] != f) continue;
                const uint32_t ns = S.ns[f];
#if MOVE_HINTS & 1
                if (lane < ns) dsim_cp_async16_hint(&S.view[f][lane], &a.ps.view[S.r_dst[f][lane]], pol_keep);
                if (lane + 32u < ns && lane + 32u < MOVE_SSTAGE)
                    dsim_cp_async16_hint(&S.view[f][lane + 32u], &a.ps.view[S.r_dst[f][lane + 32u]], pol_keep);
#else
                if (lane < ns) dsim_cp_async16(&S.view[f][lane], &a.ps.view[S.r_dst[f][lane]]);
                if (lane + 32u < ns && lane + 32u < MOVE_SSTAGE) dsim_cp_async16(&S.view[f][lane + 32u], &a.ps.view[S.r_dst[f][lane + 32u]]);
#endif
            }
            dsim_cp_async_commit();
        }
        __syncwarp();

        /* ---- decide_on_moving, lib.rs:830-889 ------------------------------------------------------- *
         * sensed patches: the leading entries of the reach row, in ascending node order; the noise index of a
         * sensed patch is its position among the sensed patches that have a travel cost, i.e. in the row */
        dsim_cp_async_wait_all();
        __syncwarp();
        for (uint32_t f = 0; f < MOVE_WF; ++f) {
            if (!((live_mask >> f) & 1u)) continue;
            const uint32_t ns = S.ns[f];
            if (ns == 0u) continue;
            const uint64_t fid = S.fid[f], culture = S.culture[f];
            const uint32_t loc = S.loc[f], size = S.size[f];
            Choice ch;
            ch.target = S.target[f];
            ch.max_gain = S.max_gain[f];
            ch.t_cost = S.t_cost[f];
            ch.m = S.m[f];
            { /* first pass: entries 0..63 from shared memory.  Lane L takes entries L and 32 + L; out-of-range lanes read
               * entry 0 and are masked at the acceptance, so the loads need no branches */
                const bool vA = lane < ns, vB = lane + 32u < ns;
                const uint32_t iA = vA ? lane : 0u, iB = vB ? lane + 32u : 0u;
                const uint32_t sf = S.src[f];
                const uint32_t qA = S.r_dst[sf][iA], qB = S.r_dst[sf][iB];
                const double cA = S.r_cost[sf][iA], cB = S.r_cost[sf][iB];
                const PatchView pA = S.view[sf][iA], pB = S.view[sf][iB];
                /* entries e and e + 32 of a pass share one Philox block (noise_block / noise_pick) */
                const dsim_u4 o = dsim_draw(rs, step, fid, DSIM_P_NOISE, noise_block(lane));
                const double gainA = view_quality(a, pA, qA, culture, size, qA == loc);
                accept_scan_warp(ch, vA, gainA * dsim_u53(o.x, o.y) - cA, gainA, cA, qA);
                if (32u < ns) {
                    const double gainB = view_quality(a, pB, qB, culture, size, qB == loc);
                    accept_scan_warp(ch, vB, gainB * dsim_u53(o.z, o.w) - cB, gainB, cB, qB);
                }
            }
            if (ns > MOVE_SSTAGE) { /* longer lists of sensed patches: further passes straight from the row */
                const uint32_t *r_dst = a.rt.dst + S.row_start[f];
                const double *r_cost = a.rt.cost + S.row_start[f];
                for (uint32_t base = MOVE_SSTAGE; base < ns; base += 64u) {
                    const uint32_t eA = base + lane, eB = eA + 32u;
                    const bool vA = eA < ns, vB = eB < ns;
                    const uint32_t iA = vA ? eA : ns - 1u, iB = vB ? eB : ns - 1u;
                    const uint32_t qA = r_dst[iA], qB = r_dst[iB];
                    const double cA = r_cost[iA], cB = r_cost[iB];
                    const PatchView pA = a.ps.view[qA], pB = a.ps.view[qB];
                    const dsim_u4 o = dsim_draw(rs, step, fid, DSIM_P_NOISE, noise_block(eA));
                    const double gainA = view_quality(a, pA, qA, culture, size, qA == loc);
                    accept_scan_warp(ch, vA, gainA * dsim_u53(o.x, o.y) - cA, gainA, cA, qA);
                    if (base + 32u < ns) {
                        const double gainB = view_quality(a, pB, qB, culture, size, qB == loc);
                        accept_scan_warp(ch, vB, gainB * dsim_u53(o.z, o.w) - cB, gainB, cB, qB);
                    }
                }
            }
            if (lane == 0) {
                S.target[f] = ch.target;
                S.max_gain[f] = ch.max_gain;
                S.t_cost[f] = ch.t_cost;
                S.m[f] = ch.m;
            }
        }
        __syncwarp();

        /* ---- remembered patches, lib.rs:854-867: the kept entries of the whole tile, family-major ------ *
         * round 0 lists the kept head entries and the first MOVE_TAILQ kept tail entries of every family; further
         * rounds (a family with more kept tail entries than that: practically never) list tail entries only */
        for (uint32_t round = 0;; ++round) {
            for (uint32_t win = 0;; win += MOVE_LCAP) {
                /* positions by a scan over the items in (family, block) order; entries inside the window are listed */
                uint32_t run = 0;
                for (uint32_t it0 = 0; it0 < n_items; it0 += 32u) {
                    const uint32_t it = it0 + lane;
                    uint32_t f = 0, j = 0, mask = 0, cnt = 0;
                    if (it < n_items) {
                        f = it / (HB + 1u);
                        j = it - f * (HB + 1u);
                        if (!((live_mask >> f) & 1u)) {
                            cnt = 0u;
                        } else if (j < HB) {
                            mask = round == 0u ? (uint32_t)S.hmask[f][j] : 0u;
                            cnt = (uint32_t)__popc(mask);
                        } else {
                            cnt = S.ntail[f];
                        }
                    }
                    const uint32_t incl = warp_incl_scan_u32(cnt, lane);
                    uint32_t p = run + incl - cnt;
                    run += __shfl_sync(DSIM_FULL, incl, 31);
                    if (j < HB) {
                        while (mask) {
                            const uint32_t b = (uint32_t)__ffs((int)mask) - 1u;
                            mask &= mask - 1u;
                            if (p >= win && p < win + MOVE_LCAP) S.list[p - win] = (f << 24) | (16u * j + b);
                            ++p;
                        }
                    } else {
                        for (uint32_t k = 0; k < cnt; ++k, ++p)
                            if (p >= win && p < win + MOVE_LCAP) S.list[p - win] = (f << 24) | S.tail[f][k];
                    }
                }
                __syncwarp();
                const uint32_t total = run;
                if (total <= win) break;
                const uint32_t nwin = total - win < MOVE_LCAP ? total - win : MOVE_LCAP;
                /* The look-up of a remembered patch is a chain of three dependent loads: ring entry -> bucket of the
                 * reach row's index -> row entry (node to verify the bucket's answer, cost).  Every level is requested for
                 * ALL the listed entries at once — cp.async into the staging arrays, which the sensed phase no longer needs —
                 * so a tile pays three memory round trips here, not three per pass.  Lane L owns entries L, L + 32, ... */
                uint32_t verify = 0, search = 0; /* per pass of this lane: row entry requested / needs the binary search */
#pragma unroll
                for (uint32_t p = 0; p < MOVE_LCAP / 32u; ++p) {
                    const uint32_t i = 32u * p + lane;
                    if (i < nwin) {
                        const uint32_t item = S.list[i], f = item >> 24, nn = item & 0xFFFFFFu; /* entry nn of family f, 0 = newest */
                        const uint32_t newest = S.newest[f];
                        const uint32_t pos = newest >= nn ? newest - nn : newest + hcap - nn;
                        const size_t hb = (size_t)S.hs[f] * hcap;
#if MOVE_HINTS & 4
                        dsim_cp_async4_hint(&rq[i], &a.hv.hist_patch[hb + pos], pol_stream);
                        dsim_cp_async8_hint(&rgain[i], &a.hv.hist_harv[hb + pos], pol_stream);
#else
                        dsim_cp_async4(&rq[i], &a.hv.hist_patch[hb + pos]);
                        dsim_cp_async8(&rgain[i], &a.hv.hist_harv[hb + pos]);
#endif
                    }
                }
                dsim_cp_async_commit();
#pragma unroll
                for (uint32_t p = 0; p < MOVE_LCAP / 32u; ++p) { /* the noise draws, while the ring entries are in flight */
                    const uint32_t i = 32u * p + lane;
                    if (i < nwin) {
                        const uint32_t item = S.list[i], f = item >> 24, nn = item & 0xFFFFFFu;
                        ru[i] = noise_draw(a.mc, step, S.fid[f], S.n_near[f] + nn);
                    }
                }
                dsim_cp_async_wait_all();
#pragma unroll
                for (uint32_t p = 0; p < MOVE_LCAP / 32u; ++p) {
                    const uint32_t i = 32u * p + lane;
                    if (i < nwin) {
                        const uint32_t f = S.list[i] >> 24;
                        if (S.row_len[f] <= DSIM_ROW_HASH_MAX) dsim_cp_async16(&rbk[i], bucket_of(a.rt.hash + (size_t)S.loc[f] * 256u, rq[i]));
                    }
                }
                dsim_cp_async_commit();
                dsim_cp_async_wait_all();
#pragma unroll
                for (uint32_t p = 0; p < MOVE_LCAP / 32u; ++p) {
                    const uint32_t i = 32u * p + lane;
                    if (i < nwin) {
                        const uint32_t f = S.list[i] >> 24, r_len = S.row_len[f];
                        if (r_len <= DSIM_ROW_HASH_MAX) {
                            const uint32_t bp = bucket_probe(rbk[i], rq[i]);
                            if (bp == ROW_ABSENT) {
                                /* no travel cost from here: not a candidate (lib.rs:869) */
                            } else if (bp < r_len) {
                                dsim_cp_async4(&rdq[i], a.rt.dst + S.row_start[f] + bp);
                                dsim_cp_async8(&rc[i], a.rt.cost + S.row_start[f] + bp);
                                verify |= 1u << p;
                            } else {
                                search |= 1u << p;
                            }
                        } else {
                            search |= 1u << p;
                        }
                    }
                }
                dsim_cp_async_commit();
                dsim_cp_async_wait_all();
                for (uint32_t p0 = 0; p0 < nwin; p0 += 32u) {
                    const uint32_t p = p0 >> 5, i = p0 + lane;
                    const bool on = i < nwin;
                    const uint32_t item = S.list[on ? i : p0];
                    const uint32_t f = item >> 24;
                    bool cand = false;
                    double g = 0.0, gain = 0.0, cost = 0.0;
                    uint32_t q = 0;
                    if (on) {
                        q = rq[i];
                        gain = rgain[i];
                        bool do_search = ((search >> p) & 1u) != 0u;
                        if ((verify >> p) & 1u) {
                            if (rdq[i] == q) { /* the bucket's answer holds */
                                cost = rc[i];
                                cand = true;
                            } else {           /* tag collision (1/256 per slot) */
                                do_search = true;
                            }
                        }
                        if (do_search) {
                            const uint32_t *r_dst = a.rt.dst + S.row_start[f];
                            const uint32_t idx = row_find_sorted(r_dst, S.ns[f], S.row_len[f], q);
                            if (idx != ROW_ABSENT) {
                                cost = a.rt.cost[S.row_start[f] + idx];
                                cand = true;
                            }
                        }
                        if (cand) g = gain * ru[i] - cost;
                    }
                    /* ordered acceptance, one family of the pass after the other (the list is family-major) */
                    const uint32_t n_act = nwin - p0 < 32u ? nwin - p0 : 32u;
                    const uint32_t f_lo = __shfl_sync(DSIM_FULL, f, 0), f_hi = __shfl_sync(DSIM_FULL, f, (int)(n_act - 1u));
                    for (uint32_t ff = f_lo; ff <= f_hi; ++ff) {
                        Choice ch;
                        ch.target = S.target[ff];
                        ch.max_gain = S.max_gain[ff];
                        ch.t_cost = S.t_cost[ff];
                        ch.m = S.m[ff];
                        accept_scan_warp(ch, cand && f == ff, g, gain, cost, q);
                        if (lane == 0) {
                            S.target[ff] = ch.target;
                            S.max_gain[ff] = ch.max_gain;
                            S.t_cost[ff] = ch.t_cost;
                            S.m[ff] = ch.m;
                        }
                        __syncwarp();
                    }
                }
                if (win + MOVE_LCAP >= total) break;
            }
            /* families whose tail list filled up before their tail was exhausted go on from where they stopped */
            const bool cont = lane < MOVE_WF && ((live_mask >> lane) & 1u) != 0u && S.tmore[lane] != 0u;
            if (__ballot_sync(DSIM_FULL, cont) == 0u) break;
            if (lane < MOVE_WF) {
                uint32_t cnt = 0;
                bool more = false;
                if (cont) {
                    TailState ts;
                    ts.j0 = S.tj0[lane];
                    ts.d = S.td[lane];
                    ts.base = S.tbase[lane];
                    cnt = mem_tail_run(a.mc, step, S.fid[lane], S.nh[lane] - a.mc.mem_tail0, ts, &S.tail[lane][0], MOVE_TAILQ, &more);
                    S.tj0[lane] = ts.j0;
                    S.td[lane] = ts.d;
                    S.tbase[lane] = ts.base;
                }
                S.ntail[lane] = cnt;
                S.tmore[lane] = more ? 1u : 0u;
            }
            __syncwarp();
        }

        /* ---- epilogue: lane f ---------------------------------------------------------------------- */
        if (state & 1u) {
            const uint32_t loc = S.loc[lane], hs = S.hs[lane], hist_cnt = S.hist_cnt[lane];
            const uint32_t newpos = S.newpos[lane];
            const uint64_t fid = S.fid[lane], culture = S.culture[lane];
            const double last_harvest = S.last_harvest[lane];
            uint32_t misc = S.misc[lane];
            /* move bookkeeping, lib.rs:1826-1833 */
            uint32_t size = S.size[lane];
            const double harv0 = last_harvest / (double)size;
            double stored = S.stored[lane] + last_harvest;
            const uint32_t newloc = S.target[lane];
            stored = stored - S.t_cost[lane];
            {
                /* family.history.push((location, last_harvest / size)), lib.rs:1826-1829 */
                const uint32_t pos = hist_cnt % hcap;
                a.hv.hist_patch[(size_t)hs * hcap + pos] = loc;
                a.hv.hist_harv[(size_t)hs * hcap + pos] = harv0;
                a.hv.hist_cnt[hs] = hist_cnt + 1u;
            }
            /* maybe_procreate, lib.rs:1869-1897: the child moves in k_children */
            if (state & 2u) {
                const uint32_t n_off = ((misc & 0xFFFFu) + 1u) & 0xFFFFu;
                misc = (misc & ~0xFFFFu) | n_off;
                const double take = stored * 2.0 / (double)size;
                stored = stored - take;
                size -= 2u;
                ChildTask t;
                t.parent_id = fid;
                t.culture = culture;
                t.take = take;
                t.old_loc = loc;
                t.new_loc = newloc;
                t.parent_hs = hs;
                t.hist_cnt = hist_cnt + 1u;
                t.n_off = n_off;
                t.ready = 0u;
                a.sc.child_task[S.crank[lane]] = t;
                /* k_children runs next to this kernel: the task — and the history entry pushed above, which the child
                 * inherits — become visible before the tag that says so */
                __threadfence();
                dsim_store_release_u32(&a.sc.child_task[S.crank[lane]].ready, serial);
            }
            FamKey ky;
            ky.id = fid;
            ky.culture = culture;
            FamPlace pl;
            pl.loc = newloc;
            pl.size = size;
            pl.hs = hs;
            pl.misc = misc;
            FamStock sk;
            sk.stored = stored;
            sk.last_harvest = 0.0;
            FamClock ck;
            ck.till_adult = S.till_adult[lane];
            ck.till_mut = S.till_mut[lane];
            a.next.key[newpos] = ky;
            a.next.place[newpos] = pl;
            a.next.stock[newpos] = sk;
            a.next.clock[newpos] = ck;
            /* arrival at the new node = histogram of the counting sort by patch.  The rank the atomic returns is
             * written when the NEXT tile gets here (or after the last one): nothing of this tile waits for it. */
            if (pend_pos != 0xFFFFFFFFu) a.sc.fslot[pend_pos] = pend_slot;
            pend_pos = 0xFFFFFFFFu;
            if (!a.mg.enabled || (newloc >= a.mg.own_lo && newloc < a.mg.own_hi)) {
                pend_slot = atomicAdd(&a.ps.cnt[newloc], 1u);
                pend_pos = newpos;
            } else {
                a.sc.fslot[newpos] = 0xFFFFFFFFu; /* emigrant: counted where it arrives (k_mg_unpack) */
                mg_list_emigrant(a, newpos, newloc);
            }
        }
        __syncwarp();
        tile = __shfl_sync(DSIM_FULL, next_tile, 0);
    }
    if (pend_pos != 0xFFFFFFFFu) a.sc.fslot[pend_pos] = pend_slot;
}

/* part 1d: the children of this step (lib.rs:1840-1855, 1877-1895), one WARP per child: there are few
 * of them (about 1 % of the families), so the kernel is latency-bound and the candidates are spread
 * over the lanes. */
__global__ void __launch_bounds__(128) k_children(StepArgs a) {
    DevCtl *ctl = a.ctl;
    dsim_pdl_wait();
    const uint32_t step = ctl->t;
    const uint32_t n_children = ctl->n_children;
    const uint32_t cap = ctl->cap;
    const uint32_t hcap = a.hv.hcap, acap = a.hv.acap;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;

    /* Several GPUs: children get the ids the reference's append order (= parent order over ALL ranks) gives them, so a
     * child's rank is the number of splitting parents, on any rank, with a smaller id — counted below from the SPLIT
     * lists of all ranks.  With the peer-memory transport the lists were stored into this rank's receive area by the
     * other ranks' k_mg_split_ids while k_move ran; the kernel starts by waiting for their flags. */
    const uint64_t *split_lists = nullptr;
    const size_t split_stride = 1 + (size_t)a.mg.scap;
    if (a.mg.enabled) {
        const uint32_t round = a.mg.p2p ? *a.mg.seq : 0u;
        if (a.mg.p2p)
            mg_wait_flags(a.mg.flags, MG_FLAG_SPLIT0 + (round & 1u) * DSIM_MG_MAX_WORLD, a.mg.world, (1u << a.mg.world) - 1u, round, &ctl->errors);
        split_lists = a.mg.split_recv + (a.mg.p2p ? (size_t)(round & 1u) * a.mg.world * split_stride : 0u);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            uint64_t total = 0;
            for (uint32_t r = 0; r < a.mg.world; ++r) total += split_lists[r * split_stride];
            ctl->next_id = ctl->child_id_base + total; /* k_lifecycle's tail had added the local count only */
        }
    }

    const uint32_t serial = ctl->serial;
    for (uint32_t r = gwarp; r < n_children; r += nwarps) {
        /* the parent's k_move tile may still be running (this kernel is launched next to k_move and fills the SMs
         * its tail frees): wait for the task's tag.  Bounded: a producer that never comes raises a model error
         * instead of hanging the device. */
        {
            uint32_t spins = 0;
            while (dsim_load_acquire_u32(&a.sc.child_task[r].ready) != serial) {
                dsim_backoff();
                if (++spins > (1u << 24)) {
                    if (lane == 0) atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_COMM_TIMEOUT);
                    break;
                }
            }
            __syncwarp();
        }
        const ChildTask t = a.sc.child_task[r];
        uint32_t grank = r;
        if (a.mg.enabled) { /* the lanes stride over the gathered lists */
            const uint64_t id = a.mg.split_send[1 + r];
            uint32_t smaller = 0;
            for (uint32_t q = 0; q < a.mg.world; ++q) {
                const uint64_t *list = split_lists + q * split_stride;
                const uint32_t m = (uint32_t)list[0];
                /* four independent loads per lane and round: the scan is a chain of L2 round trips otherwise */
                for (uint32_t k0 = lane; k0 < m; k0 += 128u) {
                    uint64_t v[4];
#pragma unroll
                    for (uint32_t u = 0; u < 4u; ++u) v[u] = (k0 + 32u * u < m) ? list[1 + k0 + 32u * u] : ~0ull;
#pragma unroll
                    for (uint32_t u = 0; u < 4u; ++u) smaller += v[u] < id ? 1u : 0u;
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) smaller += __shfl_xor_sync(DSIM_FULL, smaller, d);
            grank = smaller;
        }
        const uint64_t cid = ctl->child_id_base + grank;
        const uint32_t cpos = ctl->child_pos_base + r;
        const uint32_t ft = ctl->free_top;
        const uint32_t chs = (r < ft) ? a.sc.free_hs[ft - 1u - r] : ctl->hs_hwm + (r - ft);
        if (cpos >= cap || chs >= ctl->hs_cap) {
            if (lane == 0) atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
            continue;
        }
        const RowHdr hd = a.rt.hdr[t.new_loc];
        const RowPtr row = row_of(a.rt, t.new_loc, hd);
        Choice cc;
        cc.target = t.new_loc;
        cc.max_gain = 0.0;
        cc.t_cost = 0.0;
        cc.m = 0.0;
        /* `&nearby[1..]` of the parent's OLD patch (lib.rs:1846), costs from the child's patch */
        const uint32_t nl0 = a.rt.near_off[t.old_loc], nl1 = a.rt.near_off[t.old_loc + 1];
        const uint32_t first = nl0 + ((nl1 > nl0) ? 1u : 0u);
        /* history = the parent's newest min(len, time_to_grow_adult) entries, the one pushed this step included */
        const uint32_t valid = t.hist_cnt < hcap ? t.hist_cnt : hcap;
        const uint32_t child_len = valid < a.mc.adult_reset ? valid : a.mc.adult_reset;
        const uint32_t nhc = child_len < a.mc.n_mem ? child_len : a.mc.n_mem;
        const size_t pb = (size_t)t.parent_hs * hcap, cb = (size_t)chs * hcap;

        /* The kernel is a chain of memory round trips (few children, a warp each), so everything that depends only on
         * the task is requested first and used last: the parent's newest entries (candidates AND the child's copy of
         * the history, CHILD_HR x 32 per chunk) and its adaptation slots. */
        uint32_t hp[CHILD_HR];
        double hh[CHILD_HR];
#pragma unroll
        for (uint32_t s = 0; s < CHILD_HR; ++s) {
            const uint32_t n = 32u * s + lane;
            hp[s] = 0u;
            hh[s] = 0.0;
            if (n < child_len) {
                const uint32_t pos = (t.hist_cnt - 1u - n) % hcap;
                hp[s] = a.hv.hist_patch[pb + pos];
                hh[s] = a.hv.hist_harv[pb + pos];
            }
        }
        uint16_t ad_eco = 0;
        uint32_t ad_cnt = 0;
        double ad_val = 0.0;
        if (lane < acap) {
            ad_eco = a.hv.a_eco[(size_t)t.parent_hs * acap + lane];
            ad_cnt = a.hv.a_cnt[(size_t)t.parent_hs * acap + lane];
            ad_val = a.hv.a_val[(size_t)t.parent_hs * acap + lane];
        }
        const uint32_t parent_decay = a.hv.decay[t.parent_hs];

        /* ---- sensed patches, two rounds of 32 at a time: entry -> {bucket, view, bands} -> {row entry, cost} ---- */
        uint32_t n_considered = 0; /* noise index = ordinal among the sensed patches that have a travel cost */
        for (uint32_t base = first; base < nl1; base += 64u) {
            uint32_t q[2], bp[2];
            bool on[2], cand[2];
            uint4 bk[2];
            PatchView pv[2];
            PatchBands pbn[2];
            uint32_t dq[2];
            double c[2];
#pragma unroll
            for (uint32_t s = 0; s < 2u; ++s) {
                const uint32_t jj = base + 32u * s + lane;
                on[s] = jj < nl1;
                q[s] = a.rt.near_dst[on[s] ? jj : nl1 - 1u];
            }
#pragma unroll
            for (uint32_t s = 0; s < 2u; ++s) {
                bk[s] = make_uint4(0u, 0u, 0u, 0u);
                if (row.hash != nullptr) bk[s] = *bucket_of(row.hash, q[s]);
                pv[s] = a.ps.view[q[s]];
                pbn[s] = a.ps.bands[q[s]];
            }
#pragma unroll
            for (uint32_t s = 0; s < 2u; ++s) {
                bp[s] = row.hash != nullptr ? bucket_probe(bk[s], q[s]) : ROW_SEARCH;
                const uint32_t at = bp[s] < row.len ? bp[s] : 0u;
                dq[s] = row.len ? row.dst[at] : 0xFFFFFFFFu;
                c[s] = row.len ? row.cost[at] : 0.0;
            }
#pragma unroll
            for (uint32_t s = 0; s < 2u; ++s) {
                cand[s] = false;
                if (on[s] && bp[s] != ROW_ABSENT) {
                    if (bp[s] < row.len && dq[s] == q[s]) {
                        cand[s] = true;
                    } else {
                        const uint32_t k = row_find_sorted(row.dst, row.n_sensed, row.len, q[s]);
                        if (k != ROW_ABSENT) {
                            cand[s] = true;
                            c[s] = row.cost[k];
                        }
                    }
                }
            }
#pragma unroll
            for (uint32_t s = 0; s < 2u; ++s) {
                if (s == 1u && base + 32u >= nl1) break;
                const uint32_t cmask = __ballot_sync(DSIM_FULL, cand[s]);
                double g = 0.0, gain = 0.0;
                if (cand[s]) {
                    BandsRef br;
                    br.cult0 = pbn[s].cult0;
                    br.stor_off = pbn[s].stor_off;
                    gain = quality_eval(a, pv[s], br, t.culture, 2u, q[s] == t.new_loc);
                    g = gain * noise_draw(a.mc, step, cid, n_considered + (uint32_t)__popc(cmask & ((1u << lane) - 1u))) - c[s];
                }
                n_considered += (uint32_t)__popc(cmask);
                accept_scan_warp(cc, cand[s], g, gain, cand[s] ? c[s] : 0.0, q[s]);
            }
        }

        /* ---- remembered patches and the copy of the history, CHILD_HR rounds of 32 entries at a time ---------- */
        for (uint32_t c0 = 0; c0 < child_len; c0 += 32u * CHILD_HR) {
            bool keep[CHILD_HR], cand[CHILD_HR];
            uint4 bk[CHILD_HR];
            uint32_t bp[CHILD_HR], dq[CHILD_HR];
            double c[CHILD_HR], u[CHILD_HR];
#pragma unroll
            for (uint32_t s = 0; s < CHILD_HR; ++s) {
                const uint32_t n = c0 + 32u * s + lane;
                keep[s] = n < nhc && mem_keep(a.mc, step, cid, n, nhc);
                bk[s] = make_uint4(0u, 0u, 0u, 0u);
                if (row.hash != nullptr && keep[s]) bk[s] = *bucket_of(row.hash, hp[s]);
            }
#pragma unroll
            for (uint32_t s = 0; s < CHILD_HR; ++s) { /* the draws, while the buckets are in flight */
                const uint32_t n = c0 + 32u * s + lane;
                u[s] = keep[s] ? noise_draw(a.mc, step, cid, (nl1 - first) + n) : 0.0;
            }
#pragma unroll
            for (uint32_t s = 0; s < CHILD_HR; ++s) {
                bp[s] = row.hash != nullptr ? bucket_probe(bk[s], hp[s]) : ROW_SEARCH;
                const uint32_t at = bp[s] < row.len ? bp[s] : 0u;
                dq[s] = 0xFFFFFFFFu;
                c[s] = 0.0;
                if (keep[s] && row.len) {
                    dq[s] = row.dst[at];
                    c[s] = row.cost[at];
                }
            }
#pragma unroll
            for (uint32_t s = 0; s < CHILD_HR; ++s) {
                cand[s] = false;
                if (keep[s] && bp[s] != ROW_ABSENT) {
                    if (bp[s] < row.len && dq[s] == hp[s]) {
                        cand[s] = true;
                    } else {
                        const uint32_t k = row_find_sorted(row.dst, row.n_sensed, row.len, hp[s]);
                        if (k != ROW_ABSENT) {
                            cand[s] = true;
                            c[s] = row.cost[k];
                        }
                    }
                }
            }
#pragma unroll
            for (uint32_t s = 0; s < CHILD_HR; ++s) {
                if (c0 + 32u * s >= nhc) break;
                const double cost = cand[s] ? c[s] : 0.0;
                accept_scan_warp(cc, cand[s], cand[s] ? hh[s] * u[s] - cost : 0.0, cand[s] ? hh[s] : 0.0, cost, cand[s] ? hp[s] : 0u);
            }
            /* the child's ring: chronological index k = newest-first index child_len - 1 - k */
#pragma unroll
            for (uint32_t s = 0; s < CHILD_HR; ++s) {
                const uint32_t n = c0 + 32u * s + lane;
                if (n < child_len) {
                    a.hv.hist_patch[cb + (child_len - 1u - n)] = hp[s];
                    a.hv.hist_harv[cb + (child_len - 1u - n)] = hh[s];
                }
            }
            if (c0 + 32u * CHILD_HR < child_len) { /* next chunk (longer child histories than the default 91) */
#pragma unroll
                for (uint32_t s = 0; s < CHILD_HR; ++s) {
                    const uint32_t n = c0 + 32u * CHILD_HR + 32u * s + lane;
                    hp[s] = 0u;
                    hh[s] = 0.0;
                    if (n < child_len) {
                        const uint32_t pos = (t.hist_cnt - 1u - n) % hcap;
                        hp[s] = a.hv.hist_patch[pb + pos];
                        hh[s] = a.hv.hist_harv[pb + pos];
                    }
                }
            }
        }
        /* adaptation: family.adaptation, lib.rs:1893 */
        if (lane < acap) {
            a.hv.a_eco[(size_t)chs * acap + lane] = ad_eco;
            a.hv.a_cnt[(size_t)chs * acap + lane] = ad_cnt;
            a.hv.a_val[(size_t)chs * acap + lane] = ad_val;
        }
        for (uint32_t k = 32u + lane; k < acap; k += 32u) {
            a.hv.a_eco[(size_t)chs * acap + k] = a.hv.a_eco[(size_t)t.parent_hs * acap + k];
            a.hv.a_cnt[(size_t)chs * acap + k] = a.hv.a_cnt[(size_t)t.parent_hs * acap + k];
            a.hv.a_val[(size_t)chs * acap + k] = a.hv.a_val[(size_t)t.parent_hs * acap + k];
        }
        if (lane == 0) {
            a.hv.hist_cnt[chs] = child_len;
            a.hv.decay[chs] = parent_decay;
            a.hv.parent[chs] = t.parent_id;
            a.hv.birth_index[chs] = (uint16_t)t.n_off;
            FamKey ky;
            ky.id = cid;
            ky.culture = t.culture;
            FamPlace pl;
            pl.loc = cc.target;
            pl.size = 2u;
            pl.hs = chs;
            pl.misc = 0u;
            FamStock sk;
            sk.stored = t.take - cc.t_cost; /* descendant.stored_resources -= cost */
            sk.last_harvest = 0.0;
            FamClock ck;
            ck.till_adult = a.mc.adult_reset;
            ck.till_mut = 0u;
            a.next.key[cpos] = ky;
            a.next.place[cpos] = pl;
            a.next.stock[cpos] = sk;
            a.next.clock[cpos] = ck;
            if (!a.mg.enabled || (cc.target >= a.mg.own_lo && cc.target < a.mg.own_hi)) {
                a.sc.fslot[cpos] = atomicAdd(&a.ps.cnt[cc.target], 1u);
            } else {
                a.sc.fslot[cpos] = 0xFFFFFFFFu;
                mg_list_emigrant(a, cpos, cc.target);
            }
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------ *
 * part 2a/2b: counting sort by node (families_by_location, lib.rs:602-612)
 * ------------------------------------------------------------------------------------------ */
/* warp-aggregated append: the lanes with `pred` get consecutive positions from *counter (one atomic per warp) */
__device__ __forceinline__ uint32_t warp_append(uint32_t *counter, bool pred, uint32_t lane) {
    const uint32_t m = __ballot_sync(DSIM_FULL, pred);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(DSIM_FULL, base, 0);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256) k_segments(StepArgs a) {
    /* One thread per family.  The first arrival at a node (rank 0 of the histogram that k_move, k_children and
     * k_mg_unpack counted) speaks for the node: it claims the node's segment of `members`, lists the node as occupied
     * and hands it to the patch kernel for its family count.  The six counters (all in one cache line of the control
     * block) are bumped once per BLOCK of 256 families: per warp they cost half of the kernel's time in same-address
     * atomics (profiles/r02_small_kernels.md). */
    DevCtl *ctl = a.ctl;
    dsim_pdl_wait();
    __shared__ uint32_t s_tot[8][7], s_base[8][7];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t n = ctl->n_next;
    const uint32_t n_round = (n + 255u) & ~255u; /* whole blocks take part in the barriers */
    for (uint32_t j0 = blockIdx.x * 256u; j0 < n_round; j0 += gridDim.x * 256u) {
        const uint32_t j = j0 + threadIdx.x;
        const bool first = j < n && a.sc.fslot[j] == 0u;
        uint32_t p = 0, c = 0;
        if (first) {
            p = a.next.place[j].loc;
            c = a.ps.cnt[p];
        }
        /* by family count: a thread, eight lanes, a warp or a block per node */
        const bool one = first && c == 1u && a.small_max >= 1u;
        const bool small = first && !one && c <= a.small_max;
        const bool mid = first && !one && !small && c <= a.mid_max;
        /* nodes beyond k_patch's shared-memory capacity are counted for the host, which adds k_patch_huge to the step
         * once it has seen one (a.big_max = PATCH_M from then on; until then they stay with k_patch and its HBM scratch) */
        if (first && c > PATCH_M) atomicAdd(&ctl->n_over, 1u);
        const bool huge = first && !one && !small && !mid && c > a.big_max;
        const bool big = first && !one && !small && !mid && !huge;
        /* segment starts: exclusive scan of the counts; positions in the five lists: rank among the block's threads */
        const uint32_t incl = warp_incl_scan_u32(c, lane);
        const uint32_t m_first = __ballot_sync(DSIM_FULL, first), m_one = __ballot_sync(DSIM_FULL, one);
        const uint32_t m_small = __ballot_sync(DSIM_FULL, small), m_mid = __ballot_sync(DSIM_FULL, mid), m_big = __ballot_sync(DSIM_FULL, big);
        const uint32_t m_huge = __ballot_sync(DSIM_FULL, huge);
        if (lane == 31u) {
            s_tot[w][0] = incl;
            s_tot[w][1] = (uint32_t)__popc(m_first);
            s_tot[w][2] = (uint32_t)__popc(m_one);
            s_tot[w][3] = (uint32_t)__popc(m_small);
            s_tot[w][4] = (uint32_t)__popc(m_mid);
            s_tot[w][5] = (uint32_t)__popc(m_big);
            s_tot[w][6] = (uint32_t)__popc(m_huge);
        }
        __syncthreads();
        if (threadIdx.x < 7u) { /* thread k: counter k — seven atomics in flight together, one memory round trip per block */
            uint32_t *counter = threadIdx.x == 0u ? &ctl->seg_cursor : threadIdx.x == 1u ? &ctl->n_occ[a.occ_par] : threadIdx.x == 2u ? &ctl->n_one
                                : threadIdx.x == 3u ? &ctl->n_small : threadIdx.x == 4u ? &ctl->n_mid : threadIdx.x == 5u ? &ctl->n_big : &ctl->n_huge;
            uint32_t tot = 0;
            for (uint32_t k = 0; k < 8u; ++k) tot += s_tot[k][threadIdx.x];
            uint32_t base = tot ? atomicAdd(counter, tot) : 0u;
            for (uint32_t k = 0; k < 8u; ++k) {
                s_base[k][threadIdx.x] = base;
                base += s_tot[k][threadIdx.x];
            }
        }
        __syncthreads();
        const uint32_t below = (1u << lane) - 1u;
        if (first) {
            a.ps.seg[p] = s_base[w][0] + incl - c;
            a.sc.occ[a.occ_par][s_base[w][1] + (uint32_t)__popc(m_first & below)] = p;
        }
        if (one) a.sc.plist_one[s_base[w][2] + (uint32_t)__popc(m_one & below)] = p;
        if (small) a.sc.plist_small[s_base[w][3] + (uint32_t)__popc(m_small & below)] = p;
        if (mid) a.sc.plist_mid[s_base[w][4] + (uint32_t)__popc(m_mid & below)] = p;
        if (big) a.sc.plist_big[s_base[w][5] + (uint32_t)__popc(m_big & below)] = p;
        if (huge) a.sc.plist_huge[s_base[w][6] + (uint32_t)__popc(m_huge & below)] = p;
    }
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (a.mg.enabled) return; /* the multi-GPU step recycles heavy slots in k_mg_finish (emigrants, immigrants) */
    /* heavy slots vacated this step go onto the free stack, above what the children did not pop */
    const uint32_t ft = ctl->free_top, births = ctl->n_children, dead = ctl->n_dead;
    const uint32_t popped = births < ft ? births : ft;
    const uint32_t base = ft - popped;
    for (uint32_t k = tid; k < dead; k += nth) a.sc.free_hs[base + k] = a.sc.dead_hs[k];
}

__global__ void __launch_bounds__(256) k_scatter(StepArgs a) {
    DevCtl *ctl = a.ctl;
    dsim_pdl_wait();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t n = ctl->n_next;
    for (uint32_t j = tid; j < n; j += nth) a.sc.members[a.ps.seg[a.next.place[j].loc] + a.sc.fslot[j]] = j;
    /* The end-of-step bookkeeping (heavy-slot allocator, `s.t += 1`, src/cli.rs:154) has no kernel of its own: it is
     * marked as due here — k_children, the last reader of the allocator fields, is done; the patch kernels still read
     * ctl->t, which does not change yet — and applied by the tail of the next k_lifecycle, or by the host when it
     * reads the control block (sync_ctl).  The multi-GPU step keeps its own k_mg_finish. */
    if (tid == 0 && !a.mg.enabled) ctl->advance_pending = 1u;
    /* nodes occupied last step and empty now: census 0, no band cultures (lib.rs:522-525, 614) */
    const uint32_t n_prev = ctl->n_occ[a.occ_par ^ 1u];
    for (uint32_t k = tid; k < n_prev; k += nth) {
        const uint32_t p = a.sc.occ[a.occ_par ^ 1u][k];
        if (a.ps.cnt[p] == 0u) {
            a.ps.view[p].pop = 0u;
            a.ps.view[p].nb = 0u;
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * part 2c: one block per crowded node
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
    /* the ecoregion ids of all slots, four per 8-byte load (acap is a multiple of 4), requested together: the scan
     * costs one memory round trip instead of one per slot */
    int found = -1, spare = -1;
    for (uint32_t k = 0; k < acap; k += 4u) {
        const uint2 w = *reinterpret_cast<const uint2 *>(a.hv.a_eco + base + k);
        const uint32_t e4[4] = {w.x & 0xFFFFu, w.x >> 16, w.y & 0xFFFFu, w.y >> 16};
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j) {
            if (e4[j] == eco && found < 0) found = (int)(k + j);
            if (e4[j] == DSIM_ECO_EMPTY && spare < 0) spare = (int)(k + j);
        }
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

/* marks, in the band column, a newcomer that is hostile to an earlier member: its band is decided in join order */
#define PATCH_TODO 0xFFFFFFFEu

/* one node of k_patch, worked on by the whole block; inlined twice, once with the working arrays in shared memory
 * (fixed offsets) and once with them in HBM scratch, so that neither copy carries seventeen live pointers.
 * s_ctl: two shared words (number of bands, population); X: shared scratch of the two steps that would otherwise cost
 * O(n^2) on a node of hundreds or thousands of families (the long-run regime of the continental grid: nodes of 1 400
 * families after 10^4 seasons) — the rank sort of the shuffle and the "hostile to an earlier member" test. */
#define PATCH_DMAX 64u /* distinct cultures of a node listed in shared memory (beyond: the quadratic test) */
#ifndef PATCH_ECAP
/* (band, culture) pairs of a node listed for the walk.  While a culture cooperates with itself, its members all end up in
 * one band (a later member finds the bands before it as hostile as the earlier one did, and that one's band still without
 * a hostile member, since whoever joined it since cooperates with the culture), so there are as many pairs as cultures;
 * only cooperation_threshold 0, exclusive, spreads a culture over bands: then the list overflows and every band is scanned. */
#define PATCH_ECAP 64u
#endif
#if defined(DSIM_EMU) /* the CPU debugging build counts the shortcuts of the walk, so that its tests can tell they were taken */
extern "C" unsigned long long dsim_walk_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define WALK_STAT_N(k, cnt) do { if (lane == 0u) __atomic_fetch_add(&dsim_walk_stats[k], (unsigned long long)(cnt), __ATOMIC_RELAXED); } while (0)
#else
#define WALK_STAT_N(k, cnt) do { } while (0)
#endif
#define WALK_STAT(k) WALK_STAT_N(k, 1)
struct PatchAux {
    uint32_t bstart[257]; /* rank sort: keys bucketed by their top byte */
    uint32_t bcur[256];
    uint64_t dcult[PATCH_DMAX]; /* distinct cultures in order of first appearance in the shuffled order ... */
    uint32_t dfirst[PATCH_DMAX]; /* ... and where they first appear */
    uint32_t dn;
    /* the walk of the hostile newcomers: (band, culture) pairs met so far with the position of the first such member,
     * and per band the number of members of size 0 whose stores are not +0.0 (a deadly draw still moves those) */
    uint64_t ecult[PATCH_ECAP];
    uint32_t eband[PATCH_ECAP];
    uint32_t efirst[PATCH_ECAP];
    uint32_t deadnz[PATCH_ECAP];
};
template <uint32_t NW> /* warps of the block that works on the node */
__device__ __forceinline__ void patch_node(const StepArgs &a, const PatchMem &M, uint32_t p, uint32_t n, uint32_t s0, uint32_t tid,
                                           uint32_t step, const dsim_stream rs, double season, uint32_t &errors, uint32_t *s_ctl, PatchAux &X) {
    const uint32_t lane = tid & 31u, nthr = NW * 32u;
    __syncthreads();

    /* ---- stochasticity::shuffle (lib.rs:1376): order by (keyed random number, id) ----------- */
    for (uint32_t j = tid; j < n; j += nthr) {
        const uint32_t i = a.sc.members[s0 + j];
        const uint64_t id = a.next.key[i].id;
        const dsim_u4 o = dsim_draw(rs, step, id, DSIM_P_SHUF, 0u);
        M.tidx[j] = i;
        M.tfid[j] = id;
        M.key[j] = ((uint64_t)o.y << 32) | o.x;
    }
    /* rank of (key, id): the keys are uniform 64-bit numbers, so they are first grouped by their top byte (a counting
     * sort into M.bcnt, free until the bands are counted) and a family is then compared with the n / 256 families of
     * its own group only — n^2 / 256 comparisons instead of n^2 */
    for (uint32_t b = tid; b < 257u; b += nthr) X.bstart[b] = 0u;
    __syncthreads();
    for (uint32_t j = tid; j < n; j += nthr) atomicAdd(&X.bstart[1u + (uint32_t)(M.key[j] >> 56)], 1u);
    __syncthreads();
    if (tid < 32u) { /* exclusive scan of the 256 counts by the first warp: eight per lane */
        uint32_t v[8], sum = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) {
            v[k] = X.bstart[1u + 8u * lane + k];
            sum += v[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(DSIM_FULL, incl, d);
            if ((int)lane >= d) incl += o;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) {
            X.bcur[8u * lane + k] = run;
            run += v[k];
            X.bstart[1u + 8u * lane + k] = run; /* bstart[b + 1] = end of group b, bstart[0] stays 0 */
        }
    }
    __syncthreads();
    for (uint32_t j = tid; j < n; j += nthr) M.bcnt[atomicAdd(&X.bcur[(uint32_t)(M.key[j] >> 56)], 1u)] = j;
    __syncthreads();
    for (uint32_t j = tid; j < n; j += nthr) {
        const uint64_t kj = M.key[j], ij = M.tfid[j];
        const uint32_t b = (uint32_t)(kj >> 56);
        uint32_t r = X.bstart[b];
        for (uint32_t k = X.bstart[b]; k < X.bstart[b + 1u]; ++k) {
            const uint32_t o = M.bcnt[k];
            const uint64_t kk = M.key[o], ik = M.tfid[o];
            r += (kk < kj || (kk == kj && ik < ij)) ? 1u : 0u;
        }
        M.idx[r] = M.tidx[j];
        M.fid[r] = ij;
    }
    __syncthreads();
    for (uint32_t r = tid; r < n; r += nthr) {
        const uint32_t i = M.idx[r];
        M.cult[r] = a.next.key[i].culture;
        M.size[r] = a.next.place[i].size;
        M.stored[r] = a.next.stock[i].stored;
    }
    __syncthreads();

    /* ---- cooperate_or_fight, lib.rs:1380-1399; encounter_maybe_join, lib.rs:1209-1232 --------
     * A newcomer that is hostile to none of the members before it finds no hostile member in the first band, joins
     * it and changes nothing else: those are settled here, all at once.  Only the others walk the bands. */
    /* "hostile to a member before it" = hostile to a culture that FIRST appears before it.  Band members take their band's
     * culture every step, so a node of a thousand families holds a handful of distinct cultures: the first warp lists
     * them with their first positions (32 members per round), and every member is then tested against the list — n d
     * tests instead of n^2 / 2.  A node with more than PATCH_DMAX cultures takes the quadratic test. */
    if (tid < 32u) {
        uint32_t d = 0;
        for (uint32_t i0 = 0; i0 < n && d <= PATCH_DMAX; i0 += 32u) {
            const bool in = i0 + lane < n;
            const uint64_t c = in ? M.cult[i0 + lane] : 0ull;
            bool fresh = in;
            for (uint32_t k = 0; k < d && k < PATCH_DMAX; ++k) fresh = fresh && X.dcult[k] != c;
            uint32_t m = __ballot_sync(DSIM_FULL, fresh);
            while (m != 0u) { /* the new cultures of this round, in member order; equal ones collapse onto the first */
                const int L = __ffs((int)m) - 1;
                const uint64_t cL = __shfl_sync(DSIM_FULL, c, L);
                if (d < PATCH_DMAX && lane == 0u) {
                    X.dcult[d] = cL;
                    X.dfirst[d] = i0 + (uint32_t)L;
                }
                ++d;
                m &= ~__ballot_sync(DSIM_FULL, fresh && c == cL);
            }
            __syncwarp();
        }
        if (lane == 0u) X.dn = d;
        for (uint32_t k = lane; k < PATCH_ECAP; k += 32u) X.deadnz[k] = 0u;
    }
    __syncthreads();
    {
        const uint32_t d = X.dn;
        for (uint32_t r = tid; r < n; r += nthr) {
            const uint64_t c = M.cult[r];
            bool h = false;
            if (d <= PATCH_DMAX) {
                for (uint32_t k = 0; k < d; ++k) h |= X.dfirst[k] < r && !dsim_will_cooperate(c, X.dcult[k], a.mc.threshold, a.mc.inclusive);
            } else {
                for (uint32_t j = 0; j < r; ++j) h |= !dsim_will_cooperate(c, M.cult[j], a.mc.threshold, a.mc.inclusive);
            }
            M.band[r] = h ? PATCH_TODO : 0u;
            if (!h && M.size[r] == 0u && dsim_f64_to_bits(M.stored[r]) != 0ull) atomicAdd(&X.deadnz[0], 1u);
        }
    }
    if (tid == 0) s_ctl[1] = 0u;
    __syncthreads();
    if (tid < 32u) { /* the first warp: hostile newcomers one after the other */
        /* A newcomer walks the bands in creation order and fights every hostile member of a band before it moves on
         * (lib.rs:1380-1399).  On a node of a thousand families that was a scan of all earlier members per band and
         * newcomer.  Two facts shorten it without changing a result: (1) whether a band holds a hostile member follows
         * from the (band, culture) pairs met so far — a few dozen, tested by the lanes in one round; a band without one
         * is joined without a scan; (2) a newcomer that is down to size 0 changes nothing in a fight (every reduction
         * is 0) unless the member is itself of size 0 and still holds stores, which deadnz counts per band: a hostile
         * band without such a member is passed without a scan, and a scan stops at the chunk where the newcomer died.
         * sm: the pair list is complete (it is given up, for the rest of the node, if it outgrows PATCH_ECAP). */
        uint32_t nb = 1, en = 0;
        bool sm = X.dn <= PATCH_DMAX && X.dn <= PATCH_ECAP;
        if (sm) { /* band 0 so far: the settled members; a culture has settled members iff its first member is settled */
            const uint32_t d = X.dn;
            for (uint32_t k0 = 0; k0 < d; k0 += 32u) {
                const uint32_t k = k0 + lane;
                const bool in0 = k < d && M.band[X.dfirst[k]] == 0u;
                const uint32_t m = __ballot_sync(DSIM_FULL, in0);
                if (in0) {
                    const uint32_t e = en + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    X.ecult[e] = X.dcult[k];
                    X.eband[e] = 0u;
                    X.efirst[e] = X.dfirst[k];
                }
                en += (uint32_t)__popc(m);
            }
            __syncwarp();
        }
        const bool reflexive = dsim_will_cooperate(0ull, 0ull, a.mc.threshold, a.mc.inclusive); /* a culture with itself */
        for (uint32_t i0 = 0; i0 < n; i0 += 32u) {
            bool mine_todo = i0 + lane < n && M.band[i0 + lane] == PATCH_TODO;
            if (sm && reflexive && __any_sync(DSIM_FULL, mine_todo)) {
                /* A newcomer whose culture band 0 already holds finds no hostile member there (whoever is in band 0 got
                 * in next to that culture), joins it without a fight and adds no pair: the 32 newcomers of a round are
                 * tested against the list at once, and only the others take their turns below.  Once a hostile culture
                 * has turned up on a node, every later member of the node's main culture is such a newcomer. */
                const uint64_t c = mine_todo ? M.cult[i0 + lane] : 0ull;
                bool home0 = false;
                for (uint32_t e = 0; e < en; ++e) home0 = home0 || (X.eband[e] == 0u && X.ecult[e] == c);
                home0 = home0 && mine_todo;
                if (home0) {
                    M.band[i0 + lane] = 0u;
                    if (M.size[i0 + lane] == 0u && dsim_f64_to_bits(M.stored[i0 + lane]) != 0ull) atomicAdd(&X.deadnz[0], 1u);
                }
                const uint32_t homed = __ballot_sync(DSIM_FULL, home0);
                WALK_STAT_N(7, __popc(homed));
                (void)homed;
                mine_todo = mine_todo && !home0;
                __syncwarp(); /* the band writes are visible to the turns below */
            }
            uint32_t todo = __ballot_sync(DSIM_FULL, mine_todo);
            while (todo != 0u) {
                const uint32_t i = i0 + (uint32_t)(__ffs((int)todo) - 1);
                todo &= todo - 1u;
                const uint64_t cf = M.cult[i], fid = M.fid[i];
                uint32_t sz = M.size[i];
                double st = M.stored[i];
                int joined = -1;
                /* the bands of the listed pairs this newcomer is hostile to, two pairs per lane (a settled member of
                 * band 0 that comes after the newcomer is not there yet for it: efirst) */
                uint32_t hband[PATCH_ECAP / 32u];
#pragma unroll
                for (uint32_t k = 0; k < PATCH_ECAP / 32u; ++k) {
                    const uint32_t e = lane + 32u * k;
                    hband[k] = 0xFFFFFFFFu;
                    if (sm && e < en && X.efirst[e] < i && !dsim_will_cooperate(cf, X.ecult[e], a.mc.threshold, a.mc.inclusive))
                        hband[k] = X.eband[e];
                }
                for (uint32_t B = 0; B < nb && joined < 0; ++B) {
                    bool any_hostile = false;
                    uint32_t dz = 1u;
                    if (sm) {
                        bool mine = false;
#pragma unroll
                        for (uint32_t k = 0; k < PATCH_ECAP / 32u; ++k) mine = mine || hband[k] == B;
                        if (!__any_sync(DSIM_FULL, mine)) {
                            WALK_STAT(0);
                            joined = (int)B;
                            break;
                        }
                        dz = X.deadnz[B];
                        if (sz == 0u && dz == 0u && dsim_f64_to_bits(st) != 0x8000000000000000ull) {
                            WALK_STAT(1);
                            continue;
                        }
                        WALK_STAT(2);
                    }
                    for (uint32_t j0 = 0; j0 < i; j0 += 32u) {
                        const uint32_t j = j0 + lane;
                        bool hostile = false;
                        if (j < i) hostile = (M.band[j] == B) && !dsim_will_cooperate(cf, M.cult[j], a.mc.threshold, a.mc.inclusive);
                        uint32_t mask = __ballot_sync(DSIM_FULL, hostile);
                        if (mask == 0u) continue;
                        any_hostile = true;
                        /* every hostile member's lane draws its own fight (the draws do not depend on the outcome of
                         * the earlier fights); only the sizes are then resolved one fight after the other, in join order */
                        uint32_t so_mine = 0;
                        double sto_mine = 0.0;
                        bool dx = false, dy = false;
                        if (hostile) {
                            so_mine = M.size[j];
                            sto_mine = M.stored[j];
                            const dsim_u4 o = dsim_draw(rs, step, fid, DSIM_P_FIGHT, j);
                            dx = o.x < a.mc.deadliness; /* fight_is_deadly, lib.rs:1219 */
                            dy = o.y < a.mc.deadliness; /* lib.rs:1226 */
                        }
                        const uint32_t mdx = __ballot_sync(DSIM_FULL, dx), mdy = __ballot_sync(DSIM_FULL, dy);
                        /* Fights that cannot change anything are not replayed: neither draw deadly; or only the newcomer's
                         * side deadly against a member of size 0 (reduction 0).  Once the newcomer itself is down to size 0
                         * every reduction is 0, and only a deadly draw against a size-0 member (whose stores then pass to the
                         * newcomer) still has an effect.  On a node of a thousand families most fights are of these kinds. */
                        const uint32_t mz0 = __ballot_sync(DSIM_FULL, hostile && so_mine == 0u);
                        mask &= mdx | (mdy & ~mz0);
                        bool pruned = false;
                        while (mask != 0u) {
                            if (sz == 0u && !pruned) {
                                mask &= mdx & __ballot_sync(DSIM_FULL, hostile && so_mine == 0u);
                                pruned = true;
                                if (mask == 0u) break;
                            }
                            const uint32_t l = (uint32_t)(__ffs((int)mask) - 1);
                            mask &= mask - 1u;
                            const uint32_t so = __shfl_sync(DSIM_FULL, so_mine, l);
                            const uint32_t reduction = sz < so ? sz : so;
                            if ((mdx >> l) & 1u) {
                                if (lane == l) so_mine = so - reduction;
                                if (so == reduction) { /* the member is wiped out: its stores go to the newcomer */
                                    const double got = __shfl_sync(DSIM_FULL, sto_mine, l);
                                    st = st + got;
                                    if (lane == l) sto_mine = 0.0;
                                    if (so == 0u && dsim_f64_to_bits(got) != 0ull) { /* one dead member with stores less */
                                        --dz;
                                        WALK_STAT(3);
                                    }
                                }
                            }
                            if ((mdy >> l) & 1u) sz -= reduction;
                        }
                        if (hostile) {
                            M.size[j] = so_mine;
                            M.stored[j] = sto_mine;
                        }
                        __syncwarp();
                        if (sm && sz == 0u && dz == 0u && dsim_f64_to_bits(st) != 0x8000000000000000ull) { /* the rest are no-ops */
                            if (j0 + 32u < i) WALK_STAT(4);
                            break;
                        }
                    }
                    if (sm) {
                        if (lane == 0u) X.deadnz[B] = dz;
                        __syncwarp();
                    } else if (!any_hostile) {
                        joined = (int)B;
                    }
                }
                if (joined < 0) joined = (int)nb++;
                if (sm) { /* the newcomer's pair, if it is a new one; its own remains, if it came through dead */
                    bool mine = false;
#pragma unroll
                    for (uint32_t k = 0; k < PATCH_ECAP / 32u; ++k) {
                        const uint32_t e = lane + 32u * k;
                        mine = mine || (e < en && X.eband[e] == (uint32_t)joined && X.ecult[e] == cf);
                    }
                    if (!__any_sync(DSIM_FULL, mine)) {
                        if (en < PATCH_ECAP) {
                            if (lane == 0u) {
                                X.ecult[en] = cf;
                                X.eband[en] = (uint32_t)joined;
                                X.efirst[en] = i;
                            }
                            ++en;
                        } else {
                            WALK_STAT(5);
                            sm = false;
                        }
                    }
                    if (sm && sz == 0u && dsim_f64_to_bits(st) != 0ull) {
                        WALK_STAT(6);
                        if (lane == 0u) X.deadnz[joined] += 1u;
                    }
                }
                if (lane == 0) {
                    M.band[i] = (uint32_t)joined;
                    M.size[i] = sz;
                    M.stored[i] = st;
                }
                __syncwarp();
            }
        }
        if (lane == 0) s_ctl[0] = nb;
    }
    __syncthreads();
    const uint32_t nb = s_ctl[0];

    /* ---- band culture = culture of a random member; drop empty bands, lib.rs:1400-1416 -------- */
    /* a WARP per band (a thread per band walked all n members twice while the rest of the block waited): the lanes take
     * 32 members at a time, ballots count the band's members and find the pick-th one in join order */
    for (uint32_t B = tid >> 5; B < nb; B += nthr >> 5) {
        uint32_t cnt = 0, tot = 0;
        for (uint32_t j0 = 0; j0 < n; j0 += 32u) {
            const uint32_t j = j0 + lane;
            const bool in = j < n && M.band[j] == B;
            cnt += (uint32_t)__popc(__ballot_sync(DSIM_FULL, in));
            tot += in ? M.size[j] : 0u;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(DSIM_FULL, tot, d);
        const dsim_u4 o = dsim_draw(rs, step, (uint64_t)p, DSIM_P_PIVOT, B);
        const uint32_t pick = (uint32_t)__umul64hi(((uint64_t)o.y << 32) | o.x, (uint64_t)cnt);
        uint64_t c = 0;
        uint32_t seen = 0;
        for (uint32_t j0 = 0; j0 < n; j0 += 32u) {
            const uint32_t j = j0 + lane;
            const bool in = j < n && M.band[j] == B;
            uint32_t m = __ballot_sync(DSIM_FULL, in);
            const uint32_t k = (uint32_t)__popc(m);
            if (seen + k > pick) { /* the (pick - seen)-th member of this round */
                for (uint32_t skip = pick - seen; skip > 0u; --skip) m &= m - 1u;
                c = __shfl_sync(DSIM_FULL, in ? M.cult[j] : 0ull, __ffs((int)m) - 1);
                break;
            }
            seen += k;
        }
        if (lane == 0u) {
            M.bcult[B] = c;
            M.bcnt[B] = cnt;
            M.btot[B] = tot;
            M.balive[B] = tot > 0u ? 1 : 0;
        }
    }
    __syncthreads();

    /* ---- storage' (lib.rs:622-633): cultures and sizes of the surviving bands, creation order -- */
    uint32_t nb_alive = 0;
    uint64_t cult0 = 0;
    if (tid < 32u)
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
    for (uint32_t r = tid; r < n; r += nthr) {
        const uint32_t B = M.band[r];
        if (!M.balive[B]) continue;
        const uint32_t i = M.idx[r];
        uint64_t c = M.bcult[B];
        uint32_t misc = a.next.place[i].misc;
        uint32_t clock = a.next.clock[i].till_mut;
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
        a.next.place[i].misc = misc;
        a.next.clock[i].till_mut = clock;
    }
    __syncthreads();

    /* ---- exploit_patch + recover per ecoregion, lib.rs:2021-2067, 2095-2109 ------------------ */
    const uint32_t e0 = a.ps.eco_off[p], e1 = a.ps.eco_off[p + 1];
    double quality = 0.0;
    for (uint32_t e = e0; e < e1; ++e) {
        const uint32_t eco = a.ps.eco_id[e];
        for (uint32_t r = tid; r < n; r += nthr) { /* raw_family_contribution, lib.rs:1947-1958 */
            double c = 0.0;
            if (M.balive[M.band[r]]) c = (double)M.size[r] * adapt_touch(a, a.next.place[M.idx[r]].hs, eco, &errors);
            M.contrib[r] = c;
        }
        __syncthreads();
        for (uint32_t B = tid >> 5; B < nb; B += nthr >> 5) { /* sum_effort in join order, a warp per band */
            double eff = 0.0;
            if (M.balive[B])
                for (uint32_t j0 = 0; j0 < n; j0 += 32u) {
                    const uint32_t j = j0 + lane;
                    const bool in = j < n && M.band[j] == B;
                    const double cj = in ? M.contrib[j] : 0.0;
                    uint32_t m = __ballot_sync(DSIM_FULL, in);
                    /* the band's members of this round, one sequential add each (lib.rs:2040-2047).  The adds are a chain
                     * whatever is done; what can be taken off the chain is the fetch of the next term: a round that is
                     * mostly this band's (the main band of a node of a thousand families) issues its 32 shuffles back to
                     * back and skips the terms of other bands by predicate */
                    if (__popc(m) >= 16) {
#pragma unroll
                        for (int l = 0; l < 32; ++l) {
                            const double v = __shfl_sync(DSIM_FULL, cj, l);
                            if ((m >> l) & 1u) eff = eff + v;
                        }
                    } else {
                        while (m != 0u) {
                            const int l = __ffs((int)m) - 1;
                            m &= m - 1u;
                            eff = eff + __shfl_sync(DSIM_FULL, cj, l);
                        }
                    }
                }
            if (lane == 0u) M.beff[B] = eff;
        }
        __syncthreads();
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
            if (tid == 0) M.bcoef[B] = coefficient;
            res = res - coefficient * eff;
            if (!(res >= 0.0)) errors |= DSIM_MODEL_RESOURCES_NEGATIVE;
        }
        __syncthreads();
        for (uint32_t r = tid; r < n; r += nthr) {
            const uint32_t B = M.band[r];
            if (M.balive[B]) M.lh[r] = M.bcoef[B] * season * M.contrib[r];
        }
        /* recover, lib.rs:2100-2108 */
        if (res < mx) {
            const dsim_u4 o = dsim_draw(rs, step, (uint64_t)p, DSIM_P_REGROW, e - e0);
            const double proportion = 2.0 * dsim_u53(o.x, o.y);
            res = dsim_oyr_min(mx, mx * a.mc.recovery * proportion + res);
        }
        quality = quality + (res + mx * a.mc.recovery); /* expected_harvest for the next step */
        __syncthreads(); /* every thread has read res[e] before it is replaced */
        if (tid == 0) a.ps.res[e] = res;
    }

    /* ---- write back ----------------------------------------------------------------------- */
    uint32_t pop = 0;
    for (uint32_t r = tid; r < n; r += nthr) {
        const uint32_t i = M.idx[r];
        const uint32_t sz = M.size[r];
        a.next.place[i].size = sz;
        a.next.stock[i].stored = M.stored[r];
        if (M.balive[M.band[r]]) {
            a.next.key[i].culture = M.cult[r];
            if (e1 > e0) a.next.stock[i].last_harvest = M.lh[r];
        }
        pop += sz;
    }
    pop += __shfl_xor_sync(DSIM_FULL, pop, 16);
    pop += __shfl_xor_sync(DSIM_FULL, pop, 8);
    pop += __shfl_xor_sync(DSIM_FULL, pop, 4);
    pop += __shfl_xor_sync(DSIM_FULL, pop, 2);
    pop += __shfl_xor_sync(DSIM_FULL, pop, 1);
    if (lane == 0 && pop != 0u) atomicAdd(&s_ctl[1], pop);
    __syncthreads();
    if (tid == 0) {
        PatchView v;
        v.quality = quality;
        v.pop = s_ctl[1];
        v.nb = nb_alive;
        a.ps.view[p] = v;
        PatchBands b;
        b.cult0 = cult0;
        b.stor_off = s0;
        b.pad_ = 0;
        a.ps.bands[p] = b;
        a.ps.cnt[p] = 0u;
    }
}

/* shared-memory working set of one node of up to MC families: 105 bytes per family */
#define PATCH_SMEM_BYTES_OF(MC) ((MC) * (6u * 4u + 5u * 8u + 5u * 8u + 1u))
#define PATCH_SMEM_BYTES PATCH_SMEM_BYTES_OF(PATCH_M)
#ifndef PATCH_HUGE_WARPS
#define PATCH_HUGE_WARPS 16u /* k_patch_huge: a block of sixteen warps per node with more than PATCH_M families ... */
#endif
#ifndef PATCH_HUGE_BLOCKS
#define PATCH_HUGE_BLOCKS 32u
#endif
#ifndef PATCH_HUGE_M
#define PATCH_HUGE_M 2048u   /* ... its working arrays in 210 KB of shared memory (one block per SM); HBM scratch beyond */
#endif

/* One block of NW warps per node of the list: crowded nodes are few, and a node of hundreds of families walks its member
 * arrays many times (every hostile newcomer meets every member, lib.rs:1380-1399) — from shared memory (up to MC
 * families), not from HBM scratch.  Two instances: k_patch (4 warps, 512 families, four blocks per SM) for nodes of
 * 33..512 families, k_patch_huge (16 warps, 2048 families, one block per SM) for the few nodes beyond — after 10^4 seasons
 * of the continental configuration 28 nodes hold 500..1 400 families each, and a block of four warps spent a millisecond
 * per step on the largest of them. */
template <uint32_t NW, uint32_t MC>
__device__ __forceinline__ void patch_block(const StepArgs &a, const uint32_t *plist, uint32_t n_list) {
    DSIM_DYN_SMEM(smem_raw);
    __shared__ uint32_t s_ctl[2];
    __shared__ PatchAux s_aux;
    uint64_t *const s_u64 = reinterpret_cast<uint64_t *>(smem_raw);
    double *const s_f64 = reinterpret_cast<double *>(smem_raw + 5u * 8u * MC);
    uint32_t *const s_u32 = reinterpret_cast<uint32_t *>(smem_raw + 10u * 8u * MC);
    uint8_t *const s_u8 = smem_raw + 10u * 8u * MC + 6u * 4u * MC;

    DevCtl *ctl = a.ctl;
    const uint32_t step = ctl->t;
    const uint32_t tid = threadIdx.x;
    const dsim_stream rs = stream_of(a.mc);
    const double season = a.mc.season;
    uint32_t errors = 0;

    for (uint32_t vw = blockIdx.x; vw < n_list; vw += gridDim.x) {
        const uint32_t p = plist[vw];
        const uint32_t n = a.ps.cnt[p];
        const uint32_t s0 = a.ps.seg[p];
        PatchMem M;
        if (n <= MC) {
            M.idx = s_u32; M.tidx = s_u32 + MC; M.size = s_u32 + 2u * MC; M.band = s_u32 + 3u * MC;
            M.bcnt = s_u32 + 4u * MC; M.btot = s_u32 + 5u * MC;
            M.key = s_u64; M.fid = s_u64 + MC; M.tfid = s_u64 + 2u * MC; M.cult = s_u64 + 3u * MC;
            M.bcult = s_u64 + 4u * MC;
            M.stored = s_f64; M.contrib = s_f64 + MC; M.lh = s_f64 + 2u * MC; M.beff = s_f64 + 3u * MC;
            M.bcoef = s_f64 + 4u * MC;
            M.balive = s_u8;
            patch_node<NW>(a, M, p, n, s0, tid, step, rs, season, errors, s_ctl, s_aux);
        } else {
            M.idx = a.sc.x_idx + s0; M.tidx = a.sc.x_tidx + s0; M.size = a.sc.x_size + s0; M.band = a.sc.x_band + s0;
            M.bcnt = a.sc.x_bcnt + s0; M.btot = a.sc.x_btot + s0;
            M.key = a.sc.x_key + s0; M.fid = a.sc.x_fid + s0; M.tfid = a.sc.x_tfid + s0; M.cult = a.sc.x_cult + s0;
            M.bcult = a.sc.x_bcult + s0;
            M.stored = a.sc.x_stored + s0; M.contrib = a.sc.x_contrib + s0; M.lh = a.sc.x_lh + s0;
            M.beff = a.sc.x_beff + s0; M.bcoef = a.sc.x_bcoef + s0;
            M.balive = a.sc.x_balive + s0;
            patch_node<NW>(a, M, p, n, s0, tid, step, rs, season, errors, s_ctl, s_aux);
        }
    }
    if (errors) atomicOr(&ctl->errors, errors);
}

__global__ void __launch_bounds__(PATCH_WARPS * 32, PATCH_MIN_BLOCKS) k_patch(StepArgs a) {
    dsim_pdl_wait();
    patch_block<PATCH_WARPS, PATCH_M>(a, a.sc.plist_big, a.ctl->n_big);
}
__global__ void __launch_bounds__(PATCH_HUGE_WARPS * 32, 1) k_patch_huge(StepArgs a) {
    dsim_pdl_wait();
    patch_block<PATCH_HUGE_WARPS, PATCH_HUGE_M>(a, a.sc.plist_huge, a.ctl->n_huge);
}

/* ------------------------------------------------------------------------------------------ *
 * part 2c': a GROUP of 8 lanes per sparsely occupied node (at most PATCH_SMALL_MAX = 8 families).
 * Same rules as k_patch.  With one to a handful of families per node (the common case on the
 * continental grid) a warp per node would leave most lanes idle, and a thread per node walks its
 * families one dependent memory round trip after the other; here lane j of the group owns the family
 * at position j of the shuffled order, so the records, adaptation slots and draws of all the families
 * of a node are in flight together and four nodes share a warp.  The groups of a warp are independent:
 * every collective names only the lanes of its group.
 * ------------------------------------------------------------------------------------------ */

#ifndef PATCH_SMALL_MIN_BLOCKS
#define PATCH_SMALL_MIN_BLOCKS 8
#endif
#define PATCH_MID_MAX 32 /* nodes with 9..32 families: the same kernel with a whole warp per node (k_patch_mid) */

template <uint32_t PG> struct Grp { /* PG lanes per node: 8 (four nodes per warp) or 32 */
    uint32_t gl, goff, gmask;
    __device__ __forceinline__ explicit Grp(uint32_t lane)
        : gl(lane % PG), goff(lane - lane % PG), gmask(PG == 32u ? 0xFFFFFFFFu : (((1u << (PG & 31u)) - 1u) << (lane - lane % PG))) {}
    template <typename T> __device__ __forceinline__ T get(T v, uint32_t src) const { return __shfl_sync(gmask, v, (int)src, (int)PG); }
    __device__ __forceinline__ uint32_t ballot(bool pred) const {
        const uint32_t b = __ballot_sync(gmask, pred) >> goff;
        return PG == 32u ? b : (b & ((1u << (PG & 31u)) - 1u));
    }
    __device__ __forceinline__ uint32_t sum(uint32_t v) const {
#pragma unroll
        for (uint32_t d = PG / 2u; d > 0u; d >>= 1) v += __shfl_xor_sync(gmask, v, (int)d, (int)PG);
        return v;
    }
};

template <uint32_t PG>
__device__ __forceinline__ void patch_groups(const StepArgs &a, const uint32_t *plist, uint32_t n_list) {
    DevCtl *ctl = a.ctl;
    const uint32_t step = ctl->t;
    const dsim_stream rs = stream_of(a.mc);
    const double season = a.mc.season;
    const Grp<PG> G(threadIdx.x & 31u);
    const uint32_t gl = G.gl;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / PG, ngroups = (gridDim.x * blockDim.x) / PG;
    uint32_t errors = 0;

    for (uint32_t v = group; v < n_list; v += ngroups) {
        const uint32_t p = plist[v];
        const uint32_t n = a.ps.cnt[p];
        const uint32_t s0 = a.ps.seg[p];
        const uint32_t e0 = a.ps.eco_off[p], e1 = a.ps.eco_off[p + 1];
        const bool mine = gl < n;

        /* ---- the family at position gl of the segment: everything part 2 reads of its record ------ */
        uint32_t idx = 0, size = 0, hs = 0, misc = 0, clock = 0;
        uint64_t fid = 0, cult = 0;
        double stored = 0.0;
        if (mine) {
            idx = a.sc.members[s0 + gl];
            const FamKey ky = a.next.key[idx]; /* two 128-bit loads + two 64-bit halves */
            const FamPlace pl = a.next.place[idx];
            fid = ky.id;
            cult = ky.culture;
            size = pl.size;
            hs = pl.hs;
            misc = pl.misc;
            stored = a.next.stock[idx].stored;
            clock = a.next.clock[idx].till_mut;
        }
        /* ---- stochasticity::shuffle (lib.rs:1376): order by (keyed random number, id); lane r takes the
         * family of rank r ---------------------------------------------------------------------------- */
        if (n > 1u) {
            uint64_t key = 0;
            if (mine) {
                const dsim_u4 o = dsim_draw(rs, step, fid, DSIM_P_SHUF, 0u);
                key = ((uint64_t)o.y << 32) | o.x;
            }
            uint32_t rank = 0;
            for (uint32_t k = 0; k < n; ++k) {
                const uint64_t kk = G.get(key, k), ik = G.get(fid, k);
                rank += (kk < key || (kk == key && ik < fid)) ? 1u : 0u;
            }
            uint32_t src = gl;
            for (uint32_t k = 0; k < n; ++k)
                if (G.get(rank, k) == gl) src = k;
            idx = G.get(idx, src);
            fid = G.get(fid, src);
            cult = G.get(cult, src);
            size = G.get(size, src);
            stored = G.get(stored, src);
            hs = G.get(hs, src);
            misc = G.get(misc, src);
            clock = G.get(clock, src);
        }
        /* ---- cooperate_or_fight, lib.rs:1380-1399; encounter_maybe_join, lib.rs:1209-1232 ---------- */
        uint32_t band = 0xFFFFFFFFu, nb = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint64_t cf = G.get(cult, i), fi = G.get(fid, i);
            uint32_t sz = G.get(size, i);
            double st = G.get(stored, i);
            int joined = -1;
            for (uint32_t B = 0; B < nb && joined < 0; ++B) {
                const bool host = gl < i && band == B && !dsim_will_cooperate(cf, cult, a.mc.threshold, a.mc.inclusive);
                uint32_t hostile = G.ballot(host);
                if (hostile == 0u) {
                    joined = (int)B;
                    break;
                }
                /* every hostile member's lane draws its own fight; the sizes are then resolved in join order */
                bool dx = false, dy = false;
                if (host) {
                    const dsim_u4 o = dsim_draw(rs, step, fi, DSIM_P_FIGHT, gl);
                    dx = o.x < a.mc.deadliness; /* fight_is_deadly, lib.rs:1219 */
                    dy = o.y < a.mc.deadliness; /* lib.rs:1226 */
                }
                const uint32_t mdx = G.ballot(dx), mdy = G.ballot(dy);
                while (hostile != 0u) {
                    const uint32_t jj = (uint32_t)__ffs((int)hostile) - 1u;
                    hostile &= hostile - 1u;
                    const uint32_t so = G.get(size, jj);
                    const uint32_t reduction = sz < so ? sz : so;
                    if ((mdx >> jj) & 1u) {
                        if (gl == jj) size = so - reduction;
                        if (so == reduction) { /* the member is wiped out: its stores go to the newcomer */
                            st = st + G.get(stored, jj);
                            if (gl == jj) stored = 0.0;
                        }
                    }
                    if ((mdy >> jj) & 1u) sz -= reduction;
                }
            }
            if (joined < 0) joined = (int)nb++;
            if (gl == i) {
                band = (uint32_t)joined;
                size = sz;
                stored = st;
            }
        }
        /* ---- band culture = culture of a random member; empty bands are dropped, lib.rs:1400-1416.
         * Lane B holds band B (nb <= n <= PG). ------------------------------------------------------- */
        uint64_t bcult = 0;
        uint32_t btot = 0;
        uint64_t pivot = 0; /* lane B draws for band B */
        if (gl < nb) {
            const dsim_u4 o = dsim_draw(rs, step, (uint64_t)p, DSIM_P_PIVOT, gl);
            pivot = ((uint64_t)o.y << 32) | o.x;
        }
        for (uint32_t B = 0; B < nb; ++B) {
            const uint32_t members = G.ballot(mine && band == B);
            const uint32_t tot = G.sum((mine && band == B) ? size : 0u);
            uint32_t pick = (uint32_t)__umul64hi(G.get(pivot, B), (uint64_t)__popc(members));
            uint32_t m = members;
            while (pick > 0u) { /* the pick-th member in join order */
                m &= m - 1u;
                --pick;
            }
            const uint64_t c = G.get(cult, (uint32_t)__ffs((int)m) - 1u);
            if (gl == B) {
                bcult = c;
                btot = tot;
            }
        }
        /* storage' (lib.rs:622-633): cultures and sizes of the surviving bands, creation order */
        const uint32_t alive_bands = G.ballot(gl < nb && btot > 0u);
        if (gl < nb && btot > 0u) {
            const uint32_t k = (uint32_t)__popc(alive_bands & (((1u << gl) - 1u)));
            a.ps.band_cult[s0 + k] = bcult;
            a.ps.band_size[s0 + k] = btot;
        }
        const uint64_t cult0 = alive_bands ? G.get(bcult, (uint32_t)__ffs((int)alive_bands) - 1u) : 0ull;
        const bool balive = mine && ((alive_bands >> (mine ? band : 0u)) & 1u) != 0u;
        /* ---- adjust_culture / mutate, lib.rs:1960-1970, 1754-1785 --------------------------------- */
        {
            const uint64_t c_band = G.get(bcult, mine ? band : 0u);
            if (balive) {
                uint64_t c = c_band;
                const dsim_u4 o = dsim_draw(rs, step, fid, DSIM_P_MUT, 0u);
                if (!(misc & DSIM_MISC_CLOCK)) { /* None => draw, then re-enter, lib.rs:1763-1774 */
                    misc |= DSIM_MISC_CLOCK;
                    clock = dsim_time_till_next_mutation(dsim_u53(o.x, o.y), a.mc.log1m_rate);
                }
                if (clock == 0u) { /* Some(0): flip one of 64 bits, redraw, lib.rs:1775-1780, 1277-1285 */
                    c ^= (uint64_t)1 << (o.z >> 26);
                    const dsim_u4 o2 = dsim_draw(rs, step, fid, DSIM_P_MUT, 1u);
                    clock = dsim_time_till_next_mutation(dsim_u53(o2.x, o2.y), a.mc.log1m_rate);
                } else {
                    clock -= 1u;
                }
                cult = c;
                a.next.place[idx].misc = misc;
                a.next.clock[idx].till_mut = clock;
            }
        }
        /* ---- exploit_patch + recover per ecoregion, lib.rs:2021-2067, 2095-2109 ------------------- */
        double quality = 0.0, lh = 0.0;
        for (uint32_t e = e0; e < e1; ++e) {
            const uint32_t eco = a.ps.eco_id[e];
            double res = a.ps.res[e];
            const double mx = a.ps.max_res[e];
            /* raw_family_contribution, lib.rs:1947-1958 */
            const double contrib = balive ? (double)size * adapt_touch(a, hs, eco, &errors) : 0.0;
            double beff = 0.0; /* lane B: sum_effort of band B, in join order */
            for (uint32_t j = 0; j < n; ++j) {
                const double cj = G.get(contrib, j);
                const uint32_t bj = G.get(band, j);
                if (bj == gl) beff = beff + cj;
            }
            double total_squared = 1.0;
            for (uint32_t B = 0; B < nb; ++B) {
                const double eff = G.get(beff, B);
                if ((alive_bands >> B) & 1u) total_squared = total_squared + eff * eff;
            }
            double bcoef = 0.0;
            for (uint32_t B = 0; B < nb; ++B) {
                const double eff = G.get(beff, B);
                if (!((alive_bands >> B) & 1u)) continue;
                if (!(eff > 0.0)) errors |= DSIM_MODEL_EFFORT_NOT_POSITIVE;
                const double harvest = dsim_oyr_min(res * (eff * eff) / total_squared, a.mc.cap_harvest * season * eff);
                if (!(harvest > 0.0)) errors |= DSIM_MODEL_HARVEST_NOT_POSITIVE;
                const double coefficient = dsim_oyr_min(harvest / eff, 2.5);
                if (!(coefficient > 0.0)) errors |= DSIM_MODEL_COEFF_NOT_POSITIVE;
                if (gl == B) bcoef = coefficient;
                res = res - coefficient * eff;
                if (!(res >= 0.0)) errors |= DSIM_MODEL_RESOURCES_NEGATIVE;
            }
            const double my_coef = G.get(bcoef, mine ? band : 0u);
            if (balive) lh = my_coef * season * contrib;
            /* recover, lib.rs:2100-2108 */
            if (res < mx) {
                const dsim_u4 o = dsim_draw(rs, step, (uint64_t)p, DSIM_P_REGROW, e - e0);
                res = dsim_oyr_min(mx, mx * a.mc.recovery * (2.0 * dsim_u53(o.x, o.y)) + res);
            }
            if (gl == 0u) a.ps.res[e] = res;
            quality = quality + (res + mx * a.mc.recovery); /* expected_harvest for the next step */
        }
        /* ---- write back ----------------------------------------------------------------------------- */
        const uint32_t pop = G.sum(mine ? size : 0u);
        if (mine) {
            a.next.place[idx].size = size;
            a.next.stock[idx].stored = stored;
            if (balive) {
                a.next.key[idx].culture = cult;
                if (e1 > e0) a.next.stock[idx].last_harvest = lh;
            }
        }
        if (gl == 0u) {
            PatchView pv;
            pv.quality = quality;
            pv.pop = pop;
            pv.nb = (uint32_t)__popc(alive_bands);
            a.ps.view[p] = pv;
            PatchBands pb;
            pb.cult0 = cult0;
            pb.stor_off = s0;
            pb.pad_ = 0;
            a.ps.bands[p] = pb;
            a.ps.cnt[p] = 0u;
        }
    }
    if (errors) atomicOr(&ctl->errors, errors);
}

/* a family alone on its node: a thread per node; a kernel of its own so that its chain of memory round trips runs
 * next to the one of the eight-lane groups instead of ahead of it */
__global__ void __launch_bounds__(128, PATCH_SMALL_MIN_BLOCKS) k_patch_one(StepArgs a) {
    dsim_pdl_wait();
    patch_groups<1u>(a, a.sc.plist_one, a.ctl->n_one);
}
__global__ void __launch_bounds__(128, PATCH_SMALL_MIN_BLOCKS) k_patch_small(StepArgs a) {
    dsim_pdl_wait();
    patch_groups<8u>(a, a.sc.plist_small, a.ctl->n_small);
}
__global__ void __launch_bounds__(128, PATCH_SMALL_MIN_BLOCKS) k_patch_mid(StepArgs a) {
    dsim_pdl_wait();
    patch_groups<32u>(a, a.sc.plist_mid, a.ctl->n_mid);
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
        const FamPlace pl = f.place[i];
        const uint32_t p = pl.loc;
        atomicAdd(&ps.view[p].pop, pl.size);
        /* remember the node as "occupied last step" so that the first step clears stale censuses */
        if (atomicAdd(&ps.cnt[p], 1u) == 0u) occ[atomicAdd(&ctl->n_occ[parity], 1u)] = p;
    }
}

__global__ void k_iota(uint32_t *v, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
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
    /* every kernel of the step starts with dsim_pdl_wait(): with g.pdl they are launched programmatically dependent on
     * their predecessor (dsim_rt.h) */
#define STEP_LAUNCH(kernel, grid, block, smem)                   \
    do {                                                         \
        if ((g.pdl >> which) & 1u)                               \
            DSIM_LAUNCH_PDL(kernel, grid, block, smem, s, a);    \
        else                                                     \
            DSIM_LAUNCH(kernel, grid, block, smem, s, a);        \
    } while (0)
    switch (which) {
    case DSIM_K_LIFECYCLE: DSIM_LAUNCH(k_lifecycle, grid_for(g, LIFE_GRID_PER_SM), LIFE_BLOCK, 0, s, a); break; /* first of the step: a full dependency on what precedes */
    case DSIM_K_MOVE: STEP_LAUNCH(k_move, grid_for(g, MOVE_MIN_BLOCKS), MOVE_WARPS * 32, 0); break;
    case DSIM_K_CHILDREN: STEP_LAUNCH(k_children, grid_for(g, 2), 128, 0); break;
    case DSIM_K_SEGMENTS: STEP_LAUNCH(k_segments, grid_for(g, 8), 256, 0); break;
    case DSIM_K_SCATTER: STEP_LAUNCH(k_scatter, grid_for(g, 8), 256, 0); break;
    case DSIM_K_PATCH_ONE: STEP_LAUNCH(k_patch_one, grid_for(g, PATCH_SMALL_MIN_BLOCKS), 128, 0); break;
    case DSIM_K_PATCH_SMALL: STEP_LAUNCH(k_patch_small, grid_for(g, 2 * PATCH_SMALL_MIN_BLOCKS), 128, 0); break;
    case DSIM_K_PATCH_MID: STEP_LAUNCH(k_patch_mid, grid_for(g, PATCH_SMALL_MIN_BLOCKS), 128, 0); break;
    case DSIM_K_PATCH: STEP_LAUNCH(k_patch, grid_for(g, PATCH_MIN_BLOCKS), PATCH_WARPS * 32, PATCH_SMEM_BYTES); break;
    case DSIM_K_PATCH_HUGE: /* a block takes a whole SM's shared memory: few blocks (the nodes of this class are few), so that the other patch
                             * kernels keep most SMs while this one finds out whether it has anything to do */
        STEP_LAUNCH(k_patch_huge, PATCH_HUGE_BLOCKS, PATCH_HUGE_WARPS * 32, PATCH_SMEM_BYTES_OF(PATCH_HUGE_M));
        break;
    default: break;
    }
#undef STEP_LAUNCH
}

void dsim_launch_iota(uint32_t *v, uint32_t n, const LaunchGeom &g, cudaStream_t s) {
    DSIM_LAUNCH(k_iota, grid_for(g, 4), 256, 0, s, v, n);
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

int dsim_configure_kernels(LaunchGeom *g) {
    (void)g;
#if !defined(DSIM_EMU)
    cudaError_t e = cudaFuncSetAttribute(k_patch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PATCH_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_patch_huge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PATCH_SMEM_BYTES_OF(PATCH_HUGE_M));
    if (e != cudaSuccess) return (int)e;
#endif
    return 0;
}
