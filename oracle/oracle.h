/*
 * oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a CPU restatement of the reference algorithm
 * (supplement/dispersal_model_rust/src/lib.rs `step` and everything it calls).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product (dispersal-simulation_b200/) never includes, links or calls anything in this directory.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, golden vectors or fixtures for
 * this path (one stale doctest, lib.rs:2087-2092, contradicts the function it documents), it does
 * not compile at this revision, and no Rust toolchain exists in this environment.  The oracle is
 * pinned instead by hand-derived known-answer vectors (tests/test_oracle_kat.py) and
 * by the Random123 known-answer vectors for Philox4x32-10.
 *
 * The entry points mirror include/dsim.h one to one (prefix `oracle_`), so one Python binding
 * drives both sides.
 */
#ifndef DSIM_ORACLE_H
#define DSIM_ORACLE_H

#include "../include/dsim.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_handle oracle_handle;

int  oracle_abi_version(void);
void oracle_default_params(dsim_params *p);
void oracle_default_options(dsim_options *o);
int  oracle_create(const dsim_params *p, const dsim_graph *g, uint64_t seed,
                   const dsim_options *opt, oracle_handle **out);
void oracle_destroy(oracle_handle *h);
const char *oracle_last_error(const oracle_handle *h);

/* mode 0 ("reference-shaped", default): one bounded Dijkstra per family per decision, exactly where
 *        the reference runs it (lib.rs:843).
 * mode 1: the Dijkstra result is memoised per source node (same values, far fewer heap operations);
 *        used by the tests at sizes where mode 0 would take minutes. */
int oracle_set_mode(oracle_handle *h, int mode);

/* threads of the OpenMP loops (n <= 0: the runtime's default); returns the number that will be used */
int oracle_set_threads(int n);

int oracle_set_patches(oracle_handle *h, const double *res, const double *max_res);
int oracle_get_patches(oracle_handle *h, double *res, double *max_res);
int oracle_set_families(oracle_handle *h, const dsim_families *f);
int oracle_family_counts(oracle_handle *h, uint64_t *n, uint64_t *n_hist, uint64_t *n_adapt);
int oracle_get_families(oracle_handle *h, dsim_families *f);
int oracle_set_time(oracle_handle *h, uint64_t t);
int oracle_get_time(oracle_handle *h, uint64_t *t);
int oracle_storage_count(oracle_handle *h, uint64_t *n);
int oracle_get_storage(oracle_handle *h, uint32_t *node, uint64_t *culture, uint64_t *size);
int oracle_set_storage(oracle_handle *h, uint64_t n, const uint32_t *node, const uint64_t *culture, const uint64_t *size);
int oracle_step(oracle_handle *h, uint32_t n_steps, dsim_step_stats *stats);
int oracle_sync(oracle_handle *h);
int oracle_step_part(oracle_handle *h, int part);
int oracle_reach_counts(oracle_handle *h, uint64_t *n_nearby, uint64_t *n_reach);
int oracle_get_reach(oracle_handle *h, uint64_t *near_off, uint32_t *near_dst,
                     uint64_t *reach_off, uint32_t *reach_dst, double *reach_cost);

/* --- primitives exported for the known-answer tests ------------------------------------------ */
void     oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double   oracle_log(double x);                          /* the shared, FMA-free natural logarithm */
uint32_t oracle_f64_as_u32(double x);                   /* Rust `as u32`: truncate, saturate, NaN -> 0 */
/* lib.rs:1923-1937; returns has_shrunk */
int      oracle_use_resources_and_maybe_shrink(uint64_t *size, double *stored, double season);
/* lib.rs:1912-1920 */
void     oracle_maybe_grow(uint32_t *size, uint32_t *till_adult, double season);
/* lib.rs:2095-2109 with the random proportion 2*U supplied */
double   oracle_recover(double res, double max_res, double rate, double u);
/* lib.rs:907-929 with the noise draws supplied; returns the index of the chosen candidate or -1 */
int      oracle_best_location(int n, const double *gain, const double *cost, const double *u, double *t_cost);
/* lib.rs:2021-2064 for one ecoregion and bands given as (band id per family, ascending); the
 * families' adaptation values after the update are supplied.  Returns the patch resources left. */
double   oracle_exploit_one(int n, const uint32_t *size, const double *adapt, const int *band,
                            double res, double cap, double season, double *last_harvest);
/* ecology.rs:147-161 followed by lib.rs:1954 on one entry */
double   oracle_adapt_touch(double v, double minimum);
/* lib.rs:1270-1272 with U supplied */
uint32_t oracle_time_till_next_mutation(double u, double rate);
/* emergence::will_cooperate (missing from lib.rs; call sites lib.rs:958,1010,1212).  Pinned against the reference's own
 * analysis code, plotting/plot.py:100-101,138: `cultural_distance(c1, c2) < cooperation_threshold` (tests/golden/cooperation.npz) */
int      oracle_will_cooperate(uint64_t a, uint64_t b, uint32_t threshold, int inclusive);
/* constants of SURVEY.md appendix A as the oracle evaluates them */
void     oracle_constants(double season, uint32_t *adult_reset, uint32_t *grow_reset, double *range_s,
                          double *cost_per_s, double *decay, uint32_t *n_mem);

#ifdef __cplusplus
}
#endif
#endif
