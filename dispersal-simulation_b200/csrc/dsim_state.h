/*
 * dsim_state.h — HBM layout of the simulation state.
 *
 * Families (lib.rs:178-229) are split into
 *   - a 56-byte CORE record in four packed sub-records (16 + 16 + 16 + 8 bytes, an array each), kept in the
 *     reference's vector order and double-buffered: part 1 of the step reads buffer `cur` and writes the survivors (compacted,
 *     lib.rs:541) and the children (appended, lib.rs:551) into buffer `next`;
 *   - HEAVY per-family state that never moves (history ring, sparse adaptation slots, lineage),
 *     addressed through the `hs` (heavy slot) column of the core record.
 * Patches (lib.rs:142-152) keep their resources in the ecoregion CSR of the graph plus one 16-byte
 * VIEW record (and a 16-byte BANDS record) per node that holds everything the move decision gathers for a candidate patch.
 */
#ifndef DSIM_STATE_H
#define DSIM_STATE_H

#include <stdint.h>

#include "../../include/dsim.h"

#define DSIM_ROW_HASH_MAX 192u
#define DSIM_NIDX_NONE 0xFFFFu   /* reach entry that is not in the 2-hop `nearby` list */
#define DSIM_ECO_EMPTY 0xFFFFu   /* free adaptation slot */
#define DSIM_STOR_HALO 0x80000000u /* PatchView.stor_off: bit 31 = cultures live in halo_cult[bit 30] (multi-GPU) */
#define DSIM_MISC_CLOCK 0x10000u /* misc bit 16: seasons_till_next_mutation is Some(..) */

/* The 56-byte core record as four packed sub-records, one array each (index = position in the family vector): a
 * kernel that gathers a family through `members` or a segment reads it with three 128-bit loads and one 64-bit load
 * (four sectors) instead of ten scattered 4/8-byte loads (ten sectors); streaming passes stay coalesced. */
struct __align__(16) FamKey {   /* what keys the family's draws and its relations */
    uint64_t id;            /* birth-order id; keys every random draw of the family              */
    uint64_t culture;       /* Culture.in_memory[0]                                             */
};
struct __align__(16) FamPlace {
    uint32_t loc;           /* location (node id)                                               */
    uint32_t size;          /* effective_size                                                   */
    uint32_t hs;            /* heavy slot                                                       */
    uint32_t misc;          /* bits 0..15 number_offspring, bit 16 mutation clock present       */
};
struct __align__(16) FamStock {
    double stored;          /* stored_resources                                                 */
    double last_harvest;
};
struct __align__(8) FamClock {
    uint32_t till_adult;    /* seasons_till_next_adult                                          */
    uint32_t till_mut;      /* seasons_till_next_mutation (valid iff misc & DSIM_MISC_CLOCK)    */
};
struct FamCore {
    FamKey *key;
    FamPlace *place;
    FamStock *stock;
    FamClock *clock;
};

struct FamHeavy {           /* index = heavy slot */
    uint32_t *hist_cnt;     /* pushes so far; entry n (0 = newest) sits at ring position (cnt-1-n) % hcap */
    uint32_t *hist_patch;   /* [slot * hcap + pos] */
    double *hist_harv;      /* [slot * hcap + pos] */
    uint16_t *a_eco;        /* [slot * acap + k] ecoregion of adaptation slot k or DSIM_ECO_EMPTY  */
    uint32_t *a_cnt;        /* [slot * acap + k] decay count at which a_val was current           */
    double *a_val;          /* [slot * acap + k]                                                  */
    uint32_t *decay;        /* number of `adaptation * 0.95` applications so far (lib.rs:1953)    */
    uint64_t *parent;       /* lineage (descendence, lib.rs:227,1880)                             */
    uint16_t *birth_index;
    uint32_t hcap, acap;
};

struct __align__(16) PatchView { /* what decide_on_moving gathers per candidate (lib.rs:1034-1062): 16 bytes */
    double quality;         /* expected_harvest: sum_e(res + max * recovery), lib.rs:989-994       */
    uint32_t pop;           /* census cc[node], lib.rs:522-525                                    */
    uint32_t nb;            /* number of band cultures stored for the node (storage[node].len())   */
};

struct __align__(16) PatchBands { /* read only for candidates with nb > 0 (has_enemies, lib.rs:996-1017) */
    uint64_t cult0;         /* first band culture, inline                                         */
    uint32_t stor_off;      /* all of them: band_cult[stor_off ... stor_off + nb - 1]              */
    uint32_t pad_;
};

struct PatchState {
    PatchView *view;        /* [P]                                                                */
    PatchBands *bands;      /* [P]                                                                */
    uint32_t *cnt;          /* [P] arrivals during part 1 (counting sort histogram), 0 between steps */
    uint32_t *seg;          /* [P] start of the node's segment in `members`                        */
    const uint32_t *eco_off;/* [P+1]                                                              */
    const uint16_t *eco_id; /* [n_eco]                                                            */
    double *res;            /* [n_eco] Patch.resources[eco].0                                      */
    const double *max_res;  /* [n_eco] Patch.resources[eco].1                                      */
    uint64_t *band_cult;    /* [cap] storage' cultures, segment-aligned                            */
    uint32_t *band_size;    /* [cap] storage' sizes                                                */
};

struct __align__(16) RowHdr { /* one per node: where its reach row lives */
    uint32_t start;          /* first entry, a multiple of 8                                       */
    uint32_t len;            /* entries incl. padding (multiple of 8)                               */
    uint32_t n_sensed;       /* leading entries that are also in the 2-hop `nearby` list            */
    uint32_t n_near;         /* length of the node's nearby list (lib.rs:1158-1178)                 */
};

struct ReachTable {          /* replaces bounded_dijkstra per family per step (lib.rs:843)         */
    const RowHdr *hdr;       /* [P]                                                                */
    /* row = [sensed entries, ascending dst][other reachable nodes, ascending dst][padding 0xFFFFFFFF] */
    const uint32_t *dst;
    const double *cost;      /* OYR                                                                */
    const uint16_t *hash;    /* [P * 256] per-row index for rows of at most DSIM_ROW_HASH_MAX entries (longer rows are
                                binary-searched): 32 buckets x 8 slots, bucket = hash bits 8..12, slot = tag (hash
                                bits 0..7) << 8 | entry position, 0xFFFF = empty; one 16-byte load reads a bucket  */
    const uint32_t *near_off;/* [P+1] 2-hop lists (lib.rs:1158-1178), ascending                    */
    const uint32_t *near_dst;
};

struct ChildTask { /* what k_children needs about a family that split (lib.rs:1869-1897) */
    uint64_t parent_id, culture;
    double take;             /* resources_they_take */
    uint32_t old_loc, new_loc, parent_hs;
    uint32_t hist_cnt;       /* parent's history pushes incl. this step's */
    uint32_t n_off;          /* the ":<k>" of lib.rs:1880 */
    uint32_t ready;          /* = DevCtl.serial of the step once the task is complete: k_children runs next to k_move and
                                picks a task up as soon as its parent has decided (stored last, after a fence) */
};

struct StepScratch {
    uint32_t *alive_bits;   /* [cap/32] bit i: family i survives `retain` (lib.rs:541)             */
    uint32_t *split_bits;   /* [cap/32] bit i: family i procreates (lib.rs:1870)                    */
    uint32_t *word_pre;     /* [cap/32] survivors (bits 0..15) and splitting families (bits 16..31) in the words of
                               the same group of 256 families before word i                         */
    uint32_t *blk_alive;    /* [cap/256] per-256 counts, then exclusive offsets                    */
    uint32_t *blk_split;
    uint32_t *dead_hs;      /* [cap] heavy slots vacated this step                                 */
    uint32_t *free_hs;      /* [hcap_slots] free stack                                             */
    uint32_t *fslot;        /* [cap] arrival rank of family j at its new node                      */
    uint32_t *members;      /* [cap] family positions grouped by node                              */
    uint32_t *occ[2];       /* [cap] occupied nodes of this / the previous step                    */
    uint32_t *plist_one;    /* [cap] occupied nodes with exactly one family (a thread per node)              */
    uint32_t *plist_small;  /* [cap] occupied nodes with 2..8 families (eight lanes per node)                 */
    uint32_t *plist_mid;    /* [cap] occupied nodes with 9..32 families (a warp per node, registers only)    */
    uint32_t *plist_big;    /* [cap] nodes with 33..512 families (a block of four warps per node, shared memory)   */
    uint32_t *plist_huge;   /* [cap] the others (a block of sixteen warps per node, shared memory / HBM scratch)   */
    ChildTask *child_task;  /* [cap] indexed by child rank                                          */
    /* patch-kernel spill arrays for nodes with more families than fit in shared memory, indexed
     * by segment position */
    uint32_t *x_idx, *x_tidx, *x_size, *x_band, *x_bcnt, *x_btot;
    uint64_t *x_key, *x_fid, *x_tfid, *x_cult, *x_bcult;
    double *x_stored, *x_contrib, *x_lh, *x_beff, *x_bcoef;
    uint8_t *x_balive;
};

struct DevCtl {              /* one instance in device memory */
    uint32_t n_cur[2];       /* families in buffer 0 / 1                                           */
    uint32_t n_occ[2];       /* occupied nodes recorded in occ[0] / occ[1]                          */
    uint32_t n_alive, n_children, n_dead, n_next;
    uint32_t child_pos_base; /* = n_alive                                                         */
    uint32_t seg_cursor;
    uint32_t n_one, n_small, n_mid, n_big; /* lengths of plist_one / plist_small / plist_mid / plist_big         */
    uint32_t free_top, hs_hwm;
    uint32_t t;              /* State.t                                                            */
    uint32_t errors;
    uint32_t cap, hs_cap;
    uint32_t life_done;      /* blocks of k_lifecycle that have finished (the last one scans and resets it) */
    uint32_t move_next;      /* next tile of k_move to hand out (reset by the tail of k_lifecycle)           */
    uint32_t advance_pending;/* the end-of-step bookkeeping of the last step is still due (dsim_apply_advance)      */
    uint32_t n_huge, n_over; /* length of plist_huge; nodes with more families than k_patch holds in shared memory (this step) */
    uint32_t serial;         /* steps begun on this handle, never reset: tags the ChildTasks of a step (ChildTask.ready) */
    uint64_t child_id_base, next_id;
    uint64_t births, deaths, updates;
};

/* End-of-step bookkeeping: heavy slots popped by the children and pushed by the dead, `s.t += 1` (src/cli.rs:154).
 * Runs in the tail of the next k_lifecycle (device) or when the host reads the control block, whichever comes first. */
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void dsim_apply_advance(DevCtl *ctl) {
    if (!ctl->advance_pending) return;
    const uint32_t ft = ctl->free_top, births = ctl->n_children, dead = ctl->n_dead;
    const uint32_t popped = births < ft ? births : ft;
    ctl->hs_hwm += births - popped;
    ctl->free_top = ft - popped + dead;
    if (ctl->hs_hwm > ctl->hs_cap) ctl->errors |= (uint32_t)DSIM_MODEL_CAPACITY;
    ctl->t += 1u;
    ctl->advance_pending = 0u;
}

struct ModelConst {          /* parameters.rs:5-20 and derived constants, passed by value          */
    double recovery, season, cap_harvest, min_adapt, log1m_rate;
    uint32_t threshold, inclusive, deadliness;
    uint32_t adult_reset;    /* (15 / season) as u32 + 1, lib.rs:535,1878                          */
    uint32_t grow_reset;     /* (YEARS_BETWEEN_ADULTS / season) as u32, lib.rs:1910-1916            */
    uint32_t n_mem;          /* history entries the memory loop can still sample (lib.rs:856)       */
    uint32_t n_nodes;
    uint32_t k0, k1;         /* Philox key = seed                                                  */
    /* Remembered-patch sampling (lib.rs:847-866), canonical stream v2.
     * Head, n < mem_tail0 = 16 mem_head: entry n is kept iff U_n < memory_n with the 40-bit uniform
     * U_n = (B_n 2^32 + W_n) 2^-40, B_n a byte of the MEM stream and W_n a word of the MEM2 stream, i.e. iff
     * (B_n, W_n) < C_n = ceil(memory_n 2^40) lexicographically: B_n < mem_hi8[n], or B_n == mem_hi8[n] and
     * W_n < mem_lo32[n] (W_n is only drawn then).  Both tables are zero-padded to a multiple of 16 entries.
     * Tail, n >= mem_tail0 (memory_n < 2^-8): exact skip sampling against the survival products
     * mem_tail[j] = prod_{k <= j} (1 - memory_{mem_tail0 + k}), one uniform of the MEMT stream per kept entry + 1. */
    const uint8_t *mem_hi8;  /* C_n >> 32 (entries n < mem_always have C_n = 2^40 and are always kept) */
    const uint32_t *mem_lo32;/* C_n & 0xFFFFFFFF                                                   */
    const double *mem_tail;  /* [mem_tail_len]                                                     */
    uint32_t mem_always;     /* leading entries with memory_n >= 1                                  */
    uint32_t mem_head;       /* 16-entry blocks of the head (at most 16)                             */
    uint32_t mem_tail0, mem_tail_len;
};

/* ---- multi-GPU (band partition, dsim_mg_kernels.cu) ------------------------------------------------ */
struct MgInfo {              /* part of StepArgs */
    uint32_t enabled, rank, world;
    uint32_t own_lo, own_hi; /* this rank owns the nodes [own_lo, own_hi) and the families located there */
    /* what k_children needs to give its children their global ids: the ids of all families that split this step */
    const uint64_t *split_send;   /* [1 + scap] this rank's list (count, ids in local child order)                 */
    const uint64_t *split_recv;   /* [2][world][1 + scap] the lists of all ranks (second area: odd rounds, peer memory) */
    uint32_t scap, p2p;
    const uint32_t *seq;          /* peer-memory transport: the exchange round                                     */
    const uint32_t *flags;        /* ... and this rank's flag words                                                */
    uint32_t *counters;           /* MG_* counters (reset for the step by k_mg_split_ids)                           */
    uint32_t *emig_list;          /* [2][ecap] positions (in `next`) of the emigrants per side, appended by k_move / k_children */
    uint32_t ecap;
    const uint64_t *halo_cult[2]; /* band cultures of the halo views received from the lower / upper neighbour */
};

enum { MG_N_STAY = 0, MG_N_EMIG0, MG_N_EMIG1, MG_N_IMMIG, MG_N_FREED, MG_HALO_CULT0, MG_HALO_CULT1, MG_MAX_EMIG /* high-water mark of emigrants per side and step */, MG_N_COUNTERS = 8 };
#define DSIM_MG_MAX_WORLD 8u
#define MG_REC_HDR 72u /* emigrant record: 5 x u64 + 8 x u32, then adaptation slots, then history */
#if defined(__CUDACC__)
#define DSIM_HD_INLINE __host__ __device__ inline
#else
#define DSIM_HD_INLINE inline
#endif
DSIM_HD_INLINE uint32_t mg_pad8(uint32_t b) { return (b + 7u) & ~7u; }
DSIM_HD_INLINE uint32_t mg_rec_bytes(uint32_t acap, uint32_t hcap) {
    return (MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u) + acap * 8u + mg_pad8(hcap * 4u) + hcap * 8u + 15u) & ~15u;
}

struct MgBuffers {           /* device buffers of one rank; messages have a 16-byte header (count) */
    uint64_t *split_send;    /* [1 + scap]: count, ids of the local families that split this step       */
    uint64_t *split_recv;    /* [2][world][1 + scap]: the second area takes the odd rounds of the peer-memory transport */
    uint32_t *counters;      /* [MG_N_COUNTERS]                                                        */
    uint32_t *emig_list;     /* [2][ecap] positions (in `next`) of the emigrants per side              */
    uint32_t *emig_hs;       /* [2 * ecap] heavy slots given up by emigrants this step                 */
    uint32_t *fslot2;        /* [cap] arrival rank, indexed like the re-packed family buffer           */
    unsigned char *emig_send[2], *emig_recv[2]; /* [16 + ecap * rec_bytes], side 0 = lower neighbour   */
    unsigned char *halo_send[2], *halo_recv[2]; /* [16 + hw * 32 + ccap * 8]                           */
    uint64_t *halo_cult[2];  /* [ccap]                                                                 */
    uint32_t has_peer[2];
    uint32_t scap, ecap, rec_bytes, hw, ccap;
    /* peer-memory transport (dsim_mg_ipc_connect): the exchanges are remote stores over NVLink into the neighbours'
     * receive buffers followed by a flag, instead of NCCL operations */
    uint32_t p2p;
    uint32_t fuse_finish;               /* k_mg_halo_pack's last block also does k_mg_finish's work (the library's own step) */
    uint32_t *seq;                      /* [1] exchange round of this rank, bumped at the start of every step        */
    uint32_t *done;                     /* [3] blocks of a push kernel that have finished, per channel               */
    uint32_t *flags;                    /* [MG_FLAG_SPLIT0 + 2 x MAX_WORLD] round of the last complete message per source (SPLIT: per round parity) */
    unsigned char *peer_emig_recv[2];   /* side s: the receive buffer of that neighbour that faces me                 */
    unsigned char *peer_halo_recv[2];
    uint64_t *peer_split_recv[DSIM_MG_MAX_WORLD]; /* rank r's split_recv                                               */
    uint32_t *peer_flags[DSIM_MG_MAX_WORLD];
};
enum { MG_FLAG_EMIG0 = 0, MG_FLAG_HALO0 = 2, MG_FLAG_SPLIT0 = 4 };

#endif
