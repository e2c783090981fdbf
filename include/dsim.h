/*
 * dsim.h — C ABI of the B200-native step loop of the dispersal model.
 *
 * Drop-in boundary for `step()` of Anaphory/dispersal-simulation,
 * supplement/dispersal_model_rust/src/lib.rs:477-496 (called from `run`, src/cli.rs:123-156).
 * The reference has no FFI of its own; every entry point below cites the reference item it
 * replaces.  All citations are relative to supplement/dispersal_model_rust/src/.
 *
 * Conventions
 *   - every function returns an int status: 0 = DSIM_OK, negative = error (see dsim_status);
 *     nothing aborts across the ABI (the reference is `panic = "abort"`, Cargo.toml:11);
 *   - plain pointers and sizes only; the caller owns every host array it passes in or receives;
 *   - one host thread per handle, a handle is not re-entrant;
 *   - the library needs a CUDA device (sm_100a).  There is no CPU fallback: dsim_create fails
 *     with DSIM_ERR_NO_DEVICE when no GPU is present.
 *
 * The same declarations, with the prefix `oracle_` instead of `dsim_`, are exported by the
 * CPU oracle (oracle/oracle.h) so that the parity tests drive both through one binding.
 */
#ifndef DSIM_H
#define DSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSIM_ABI_VERSION 1

/* Number of ecoregion ids a family can adapt to: ecology.rs:137,143 (`[f64; ATTESTED_ECOREGIONS + 1]`). */
#define DSIM_N_ECOREGIONS 304
/* `seasons_till_next_mutation: Option<Seasons>` (lib.rs:220) is carried as value + flag. */
#define DSIM_NO_PARENT UINT64_MAX

typedef enum dsim_status {
    DSIM_OK = 0,
    DSIM_ERR_INVALID = -1,      /* bad argument (NULL, out-of-range node/ecoregion id, unsorted CSR, ...) */
    DSIM_ERR_NO_DEVICE = -2,    /* no CUDA device / wrong architecture */
    DSIM_ERR_CUDA = -3,         /* a CUDA runtime call failed; see dsim_last_error */
    DSIM_ERR_CAPACITY = -4,     /* more families than the handle can grow to */
    DSIM_ERR_MODEL = -5,        /* one of the reference's asserts failed (lib.rs:2040,2049,2055,2063) */
    DSIM_ERR_COMM = -6          /* multi-GPU transport failure */
} dsim_status;

/* Sticky model-error bits reported in dsim_step_stats.model_errors. */
#define DSIM_MODEL_EFFORT_NOT_POSITIVE   0x01u /* assert!(group_effort > 0.)            lib.rs:2040 */
#define DSIM_MODEL_HARVEST_NOT_POSITIVE  0x02u /* assert!(group_harvest > 0)            lib.rs:2049 */
#define DSIM_MODEL_COEFF_NOT_POSITIVE    0x04u /* assert!(coefficient > 0)              lib.rs:2055 */
#define DSIM_MODEL_RESOURCES_NEGATIVE    0x08u /* assert!(*res >= 0)                    lib.rs:2063 */
#define DSIM_MODEL_ADAPTATION_OVERFLOW   0x10u /* more live adaptation entries than slots (not reachable
                                                  from states whose adaptation values are <= 1)        */
#define DSIM_MODEL_CAPACITY              0x20u /* family capacity exceeded inside a batch of steps     */
#define DSIM_MODEL_COMM_TIMEOUT          0x40u /* multi-GPU, peer-memory transport: a neighbour's message did not arrive */

/* Mirrors `Parameters` (parameters.rs:5-20) minus the graph, which is passed separately.
 * Defaults: parameters.rs:22-54, filled in by dsim_default_params. */
typedef struct dsim_params {
    double   resource_recovery_per_season;             /* 0.10   */
    double   culture_mutation_rate;                    /* 6e-3   */
    uint64_t culture_dimensionality;                   /* 20; the reference only runs with 1 or 64 (lib.rs:1589) and
                                                          always flips one of 64 bits (lib.rs:1279) */
    uint32_t cooperation_threshold;                    /* 6      */
    uint32_t fight_deadliness;                         /* 1<<31: P(death) = fight_deadliness / 2^32 (lib.rs:1326) */
    double   maximum_resources_one_adult_can_harvest;  /* 1e50   */
    double   evidence_needed;                          /* 0.1, unused by the step (dead parameter) */
    double   payoff_std;                               /* 0.1, unused */
    double   minimum_adaptation;                       /* 0.5    */
    double   enemy_discount;                           /* 0.5^0.2, unused by the live code path */
    double   season_length_in_years;                   /* 1/6    */
    /* `emergence::will_cooperate` is missing from the reference (call sites lib.rs:958,1010,1212).
     * 0: cooperate iff popcount(a^b) <  threshold (default); 1: <= threshold. */
    uint32_t cooperation_inclusive;
    uint32_t reserved_;
} dsim_params;

/* Mirrors `MovementGraph` (movementgraph.rs:8-9) in CSR form.
 * Node data = (h3, lon, lat, {ecoregion -> popcap}); edge weight = travel seconds. */
typedef struct dsim_graph {
    uint32_t        n_nodes;
    const uint64_t *h3;        /* [n_nodes] may be NULL */
    const double   *lon;       /* [n_nodes] may be NULL */
    const double   *lat;       /* [n_nodes] may be NULL */
    const uint64_t *eco_off;   /* [n_nodes+1] ecoregion CSR                               */
    const uint32_t *eco_id;    /* [eco_off[n]] ecoregion ids < DSIM_N_ECOREGIONS, ascending within a node */
    const double   *popcap;    /* [eco_off[n]] may be NULL (only used by initial states)  */
    const uint64_t *edge_off;  /* [n_nodes+1] out-edge CSR, petgraph out-edge order       */
    const uint32_t *edge_dst;  /* [edge_off[n]]                                           */
    const double   *edge_sec;  /* [edge_off[n]] travel time in seconds                    */
} dsim_graph;

/* Mirrors `Vec<Family>` (lib.rs:178-229, plus the two fields the reference uses but lost:
 * `adaptation` and `last_harvest`, lib.rs:1614-1615) as structure-of-arrays.  The vector order is
 * part of the contract (lib.rs:541,551): entry 0 is `families[0]`.
 *
 * `descendence: String` (lib.rs:227) is replaced by (id, parent_id, birth_index): the family named
 * "<parent>:<k>" has parent_id = id of <parent> and birth_index = k (lib.rs:1880).  Ids are assigned
 * sequentially in the order the reference appends families to its vector.
 *
 * history (lib.rs:218) is a CSR, oldest entry first; adaptation is a sparse CSR holding the
 * entries that differ from minimum_adaptation (all other entries equal the minimum). */
typedef struct dsim_families {
    uint64_t  n;
    uint64_t *id;
    uint64_t *parent_id;
    uint16_t *birth_index;
    uint64_t *culture;            /* Culture.in_memory[0] (lib.rs:253-258) */
    uint32_t *location;           /* NodeId */
    uint32_t *size;               /* effective_size */
    double   *stored;             /* stored_resources */
    double   *last_harvest;
    uint32_t *till_adult;         /* seasons_till_next_adult */
    uint32_t *till_mutation;      /* seasons_till_next_mutation, valid iff has_mutation_clock */
    uint8_t  *has_mutation_clock; /* 0 = None */
    uint16_t *n_offspring;        /* number_offspring */
    uint64_t *hist_off;           /* [n+1] */
    uint32_t *hist_patch;
    double   *hist_harvest;
    uint64_t *adapt_off;          /* [n+1] */
    uint32_t *adapt_eco;
    double   *adapt_val;
} dsim_families;

/* Optional knobs that are not model parameters. */
typedef struct dsim_options {
    uint64_t family_capacity;  /* 0 = choose (2x the families set, at least 1024); grows on demand between steps */
    uint32_t history_capacity; /* 0 = exact: every entry the memory loop of lib.rs:854-866 can still sample; shorter rings need DSIM_OPT_SHORT_HISTORY */
    uint32_t device;           /* CUDA device ordinal */
    uint32_t use_graphs;       /* 1 = replay the step as a CUDA graph (default), 0 = plain launches */
    uint32_t flags;            /* DSIM_OPT_* test knobs, 0 by default */
} dsim_options;
#define DSIM_OPT_PATCH_WARP_ONLY 0x2u /* every occupied node takes the block-per-node kernel (k_patch) */
#define DSIM_OPT_REACH_ON_DEVICE 0x4u /* build the reach table with the device kernels: the default since round 2 (falls back
                                        * to the host builder for graphs with reach sets of more than 384 or 2-hop sets of
                                        * more than 512 nodes); the flag is accepted and changes nothing */
#define DSIM_OPT_REACH_ON_HOST 0x8u   /* build the reach table with the host (OpenMP) builder */
#define DSIM_OPT_SHORT_HISTORY 0x20u   /* accept a history_capacity below the exact ring length (test knob: saves memory in cases whose
                                        * histories never outgrow the ring; otherwise dsim_create rejects it) */
#define DSIM_OPT_NO_PDL 0x10u         /* plain stream-ordered launches instead of programmatic dependent launches (A/B knob) */

typedef struct dsim_step_stats {
    uint64_t t;               /* State.t after the call (lib.rs:308; `s.t += 1`, cli.rs:154) */
    uint64_t live_families;   /* families.len() after the last step (extinction check cli.rs:138) */
    uint64_t births;          /* children appended during this call (lib.rs:551) */
    uint64_t deaths;          /* families dropped by `retain` during this call (lib.rs:541) */
    uint64_t family_updates;  /* sum over the call's steps of families.len() at step start */
    uint64_t occupied_patches;/* patches visited by part 2 of the last step (lib.rs:615) */
    uint32_t model_errors;    /* sticky DSIM_MODEL_* bits */
    uint32_t reserved_;
} dsim_step_stats;

typedef struct dsim_handle dsim_handle;

int  dsim_abi_version(void);
void dsim_default_params(dsim_params *p);           /* parameters.rs:22-54 */
void dsim_default_options(dsim_options *o);

/* Parameters + MovementGraph -> handle.  Builds the per-patch reach table that replaces the
 * per-family-per-step `bounded_dijkstra` (movementgraph.rs:64-115 via lib.rs:1133-1154) and the
 * `nearby_cache` of 2-hop neighbourhoods (lib.rs:1158-1178, cli.rs:126).  Patch resources start at 0. */
int dsim_create(const dsim_params *p, const dsim_graph *g, uint64_t seed,
                const dsim_options *opt, dsim_handle **out);
void dsim_destroy(dsim_handle *h);
const char *dsim_last_error(const dsim_handle *h);  /* h may be NULL: error of the last failed dsim_create */

/* State.patches (lib.rs:317; Patch.resources lib.rs:151): (current, maximum) per (node, ecoregion),
 * aligned with the ecoregion CSR of the graph. */
int dsim_set_patches(dsim_handle *h, const double *res, const double *max_res);
int dsim_get_patches(dsim_handle *h, double *res, double *max_res); /* either may be NULL */

/* State.families (lib.rs:301) and State.t (lib.rs:308). */
int dsim_set_families(dsim_handle *h, const dsim_families *f);
int dsim_family_counts(dsim_handle *h, uint64_t *n, uint64_t *n_hist, uint64_t *n_adapt); /* a NULL count is not computed */
int dsim_get_families(dsim_handle *h, dsim_families *f);  /* arrays sized by dsim_family_counts; NULL members are skipped */
int dsim_set_time(dsim_handle *h, uint64_t t);
int dsim_get_time(dsim_handle *h, uint64_t *t);

/* The map `step` returns and takes back (lib.rs:480-484,489-493,654): node -> {band culture -> (size, 0)}.
 * One entry per (node, culture), sorted by (node, culture).  Also what `print_population_by_location`
 * prints (lib.rs:1443-1464): `cultures_by_location` holds the same keys and sizes (lib.rs:628-632). */
int dsim_storage_count(dsim_handle *h, uint64_t *n);
int dsim_get_storage(dsim_handle *h, uint32_t *node, uint64_t *culture, uint64_t *size);
int dsim_set_storage(dsim_handle *h, uint64_t n, const uint32_t *node, const uint64_t *culture, const uint64_t *size);

/* `step` x n_steps plus `s.t += 1` (lib.rs:477-496, cli.rs:128-154).  Enqueues on an internal stream;
 * returns after enqueueing when stats == NULL, otherwise waits and fills *stats. */
int dsim_step(dsim_handle *h, uint32_t n_steps, dsim_step_stats *stats);
int dsim_sync(dsim_handle *h);
/* `observation::print_population_by_location` (lib.rs:1443-1464): the linguistic landscape after the last step — for
 * every occupied node its coordinates (`node.1`, `node.2` of the graph's node data) and the map band culture -> total
 * population of the bands with that culture (`cultures_by_location`, lib.rs:615-647).  Grouped per node, nodes in
 * ascending id, cultures ascending within a node: entries [off[k], off[k+1]) belong to node[k].
 * Two calls: dsim_observe_population_counts for the sizes, then dsim_observe_population with arrays of
 * n_nodes (node, lon, lat), n_nodes + 1 (off) and n_entries (culture, population) elements.  lon/lat are 0 when the
 * graph was given without coordinates.  Host buffers, filled on demand (every `log_every` steps, lib.rs:649-653). */
int dsim_observe_population_counts(dsim_handle *h, uint64_t *n_nodes, uint64_t *n_entries);
int dsim_observe_population(dsim_handle *h, uint32_t *node, double *lon, double *lat, uint64_t *off, uint64_t *culture,
                            uint64_t *population);

/* As dsim_step with stats, and *device_ms receives the time between two CUDA events recorded on the
 * internal stream immediately before the first and after the last kernel of the n_steps steps. */
int dsim_step_timed(dsim_handle *h, uint32_t n_steps, dsim_step_stats *stats, double *device_ms);

/* Debugging aid for the parity tests: part 1 = update_each_family_step (lib.rs:511-552),
 * part 2 = distribute_each_patch_resources_step (lib.rs:591-655) + t += 1. */
int dsim_step_part(dsim_handle *h, int part);

/* Reach table read-back (parity tests of the precomputation against movementgraph.rs:64-115).
 * nearby: the 2-hop set of lib.rs:1158-1178 in ascending node order.
 * reach : (destination, cost in OYR) of lib.rs:1133-1154 in ascending destination order.
 * The device builder counts every node's row at dsim_create and fills the rows when the table is first needed (state
 * upload, a step, these calls): on a handle that dsim_mg_configure has split into bands by then, only the rows of the
 * own band and its halo exist (per-rank memory and fill time ~ 1/ranks); the other rows read as empty. */
int dsim_reach_counts(dsim_handle *h, uint64_t *n_nearby, uint64_t *n_reach);
int dsim_get_reach(dsim_handle *h, uint64_t *near_off, uint32_t *near_dst,
                   uint64_t *reach_off, uint32_t *reach_dst, double *reach_cost);
/* Order-independent 64-bit digests of the reach table as it lies in HBM, computed on the device: out[0] row headers,
 * out[1] destinations (padding included), out[2] costs (bit patterns), out[3] per-row bucket indices, out[4] 2-hop
 * lists and their offsets.  Two handles on the same graph hold identical tables iff the digests agree (used to compare
 * the device-built table with the host-built one at sizes where dsim_get_reach would move gigabytes). */
int dsim_reach_digest(dsim_handle *h, uint64_t out[5]);

/* ---- multi-GPU: one handle per GPU / process, contiguous node ranges ("bands") per rank ------------------------
 * New functionality (the reference is a single process).  Rank r owns the nodes [bounds[r], bounds[r+1]) and the
 * families located there; every rank is created with the full graph.  dsim_set_patches / dsim_set_storage take full
 * arrays, dsim_set_families takes the full family vector and keeps the families of the own band.  Getters return
 * the local part (families: the own ones, in no particular order; patches / storage: only the own nodes are valid).
 * Results are bit-identical to a single-GPU run.  Every band must hold at least dsim_mg_halo_width() nodes.
 *
 * A step is four phases with one message exchange after each of the first three:
 *   LIFECYCLE -> SPLIT allgather -> MOVE -> EMIGRANTS neighbour exchange -> PATCH -> HALO neighbour exchange -> FINISH.
 * With dsim_mg_init_nccl the library moves the messages itself (NCCL on its stream) and dsim_step works as on one
 * GPU; without it the caller drives dsim_mg_phase and moves the channel buffers (tests: "virtual ranks"). */
enum { DSIM_MG_PHASE_LIFECYCLE = 0, DSIM_MG_PHASE_MOVE = 1, DSIM_MG_PHASE_PATCH = 2, DSIM_MG_PHASE_FINISH = 3,
       DSIM_MG_PHASE_HALO_PACK = 4, DSIM_MG_PHASE_HALO_UNPACK = 5 /* 4,5: initial halo after the state was set */ };
enum { DSIM_MG_CH_SPLIT = 0, DSIM_MG_CH_EMIGRANTS = 1, DSIM_MG_CH_HALO = 2 };
int dsim_mg_configure(dsim_handle *h, uint32_t rank, uint32_t world, const uint32_t *bounds /* [world + 1] */);
int dsim_mg_halo_width(dsim_handle *h, uint32_t *nodes);
int dsim_mg_phase(dsim_handle *h, int phase);
/* Device buffers of a channel.  SPLIT: send = this rank's list, recv = slot `peer` (a rank) of the gathered lists.
 * EMIGRANTS / HALO: peer 0 = lower neighbour (rank-1), 1 = upper neighbour (rank+1); what a rank sends to its upper
 * neighbour arrives in that neighbour's recv buffer of peer 0.  Messages have a fixed size (*bytes). */
int dsim_mg_buffer(dsim_handle *h, int channel, int peer, int recv, void **dev_ptr, uint64_t *bytes);
int dsim_mg_memcpy(dsim_handle *h, void *dst, const void *src, uint64_t bytes); /* device-to-device copy helper */
#define DSIM_MG_UNIQUE_ID_BYTES 128
int dsim_mg_unique_id(void *out /* DSIM_MG_UNIQUE_ID_BYTES, rank 0 creates and broadcasts it */);
int dsim_mg_init_nccl(dsim_handle *h, const void *unique_id);
/* Peer-memory transport (all ranks on one NVLink node, one process per rank): every rank exports CUDA IPC handles of
 * its receive buffers (DSIM_MG_IPC_BYTES), the caller gathers them from all ranks (rank-major) and hands them back;
 * from then on dsim_mg_step_enqueue moves the three messages of a step by remote stores into the neighbours' buffers
 * plus a flag (written by the packing kernels themselves, awaited by the consuming ones) instead of NCCL operations.  Results are unchanged.  It works with or without
 * dsim_mg_init_nccl: without a communicator dsim_mg_halo_sync also goes through peer memory (call it on every rank
 * once all ranks have connected) and dsim_mg_reduce_stats is not available (the caller sums the per-rank stats). */
#define DSIM_MG_IPC_BYTES 384 /* 6 handles of 64 bytes */
int dsim_mg_ipc_export(dsim_handle *h, void *out /* DSIM_MG_IPC_BYTES */);
int dsim_mg_ipc_connect(dsim_handle *h, const void *all /* world x DSIM_MG_IPC_BYTES */);
/* Sums the per-rank counters of a dsim_step_stats over all ranks (NCCL transport only). */
int dsim_mg_reduce_stats(dsim_handle *h, dsim_step_stats *stats);
/* Diagnostics: internal per-step counters; out8[7] = most emigrants per neighbour and step seen so far. */
int dsim_mg_debug_counters(dsim_handle *h, uint32_t *out8);
int dsim_mg_segment_times(dsim_handle *h, double *out7); /* mean ms of lifecycle, SPLIT, move, EMIGRANTS, patch, HALO, finish while profiling */
int dsim_mg_kernel_times(dsim_handle *h, double *out12); /* mean ms per launch while profiling: split_ids, sort_out, pack, unpack,
                                                           * halo_pack, halo_unpack, finish, lifecycle, move, children, segments, scatter */
int dsim_mg_halo_sync(dsim_handle *h); /* built-in transports: initial halo exchange after the state was set on every rank */

/* Kernel-level timing for bench.py: device time in ms of each kernel class accumulated since the
 * last reset, measured with CUDA events on the internal stream (only while enabled; this
 * serialises the step and is off by default).  names[i] are static strings. */
#define DSIM_MAX_KERNELS 16
typedef struct dsim_kernel_times {
    uint32_t    n;
    const char *name[DSIM_MAX_KERNELS];
    double      ms[DSIM_MAX_KERNELS];
    uint64_t    launches[DSIM_MAX_KERNELS];
} dsim_kernel_times;
int dsim_profile_enable(dsim_handle *h, int on);
int dsim_profile_read(dsim_handle *h, dsim_kernel_times *out, int reset);
/* Total kernels launched by this handle (directly or inside graph replays) since creation. */
int dsim_launch_count(dsim_handle *h, uint64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* DSIM_H */
