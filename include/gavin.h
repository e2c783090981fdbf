/*
 * gavin.h — C ABI of the batched gavin2017processbased model (BASELINE.json configs[4]).
 *
 * Replaces, for a whole parameter sweep at once, the main loop of supplement/gavin2017processbased/model.py:333-378
 * with `Language.grow` / `distribute` / `dhondt` (language.py:18-99): every sweep member runs the sequence of
 * languages on its own copy of the grid state; a member is one warp of the kernel.  Inputs that the reference
 * computes from files that are not in its repository are passed in: the population capacity of every cell
 * (model.py:209-214, 271-297: alpha * precipitation ** beta * area / 1e6, evaluated by the caller so that `pow` is
 * the caller's) and the table of language population caps (popcaps.py:4-17).
 *
 * Cells: 2 * rows * cols, numbered m * rows * cols + i * cols + j (two interleaved sub-grids, model.py:62-124);
 * neighbours as in model.py:236-262.  Canonical orders and the random stream: oracle/gavin.py (header).
 */
#ifndef GAVIN_H
#define GAVIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gavin_config {
    uint32_t rows, cols;          /* of one sub-grid */
    uint32_t n_members;           /* sweep members = warps */
    uint32_t n_lang_caps;         /* entries of lang_caps */
    uint32_t max_languages;       /* stop after this many languages per member (0 = until the grid is full) */
    uint32_t device;              /* CUDA device ordinal */
} gavin_config;

typedef struct gavin_result {
    double device_ms;             /* kernel time (CUDA events) */
    uint64_t grow_calls;          /* Language.grow calls over all members */
    uint64_t cell_updates;        /* cells visited by grow over all members (the unit of the throughput figure) */
} gavin_result;

/* popcap [n_members][cells], growth_factor [n_members] (language.py:7), seed [n_members], lang_caps [n_lang_caps];
 * out: language [n_members][cells] (0 = none, ids from 1 in seeding order), population [n_members][cells],
 * n_languages [n_members].  Returns 0, or a negative DSIM_ERR_* code of include/dsim.h (no device: -2). */
int gavin_run(const gavin_config *cfg, const double *popcap, const double *growth_factor, const uint64_t *seed,
              const double *lang_caps, int32_t *language, int32_t *population, uint32_t *n_languages, gavin_result *res);
const char *gavin_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
