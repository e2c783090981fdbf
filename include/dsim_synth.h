/*
 * dsim_synth.h — synthetic hex grids and family populations (BASELINE.json configs).
 *
 * The reference reads its graph from `graph.bincode` (src/main.rs:7-18), built offline from GIS
 * rasters that are not in the repository (supplement/distances/), and starts from two families in
 * Alaska (lib.rs:1600-1631).  The benchmark configurations are synthetic up-scalings, generated
 * here once and handed, as identical arrays, to the GPU engine and to the CPU oracle.
 *
 * Host-only code (no CUDA), C ABI.
 */
#ifndef DSIM_SYNTH_H
#define DSIM_SYNTH_H

#include "dsim.h"

#ifdef __cplusplus
extern "C" {
#endif

/* W x H hexagonal grid in odd-r offset coordinates, node id = row * W + col.  Every node has
 * out-edges to the cells at hex distance 1 and 2 (<= 18, mimicking the H3 2-ring connectivity of
 * supplement/distances/main.py:336), sorted by target id, weight = 14400 s * ring * (1 + 0.25 u).
 * One ecoregion per node, id = row * 16 / H;  popcap = clip(20 exp(N(0,1)), 0.5, 400);
 * max_res = res = popcap / recovery (lib.rs:1568 with scale 1). */
int dsim_synth_grid_sizes(uint32_t W, uint32_t H, uint64_t *n_nodes, uint64_t *n_edges);
int dsim_synth_grid(uint32_t W, uint32_t H, uint64_t seed, double recovery,
                    uint64_t *h3, double *lon, double *lat,
                    uint64_t *eco_off, uint32_t *eco_id, double *popcap,
                    uint64_t *edge_off, uint32_t *edge_dst, double *edge_sec,
                    double *res, double *max_res);

/* n families with ids 0..n-1 in vector order.
 *   n_occupied == 0: location uniform over the W*H nodes (configs C1-C3);
 *   n_occupied  > 0: families share `n_occupied` randomly chosen nodes, mildly skewed so that a few
 *                    patches hold ~100 families (config C4, dense co-residence).
 * size ~ U{2..9}, stored ~ U[0,5), till_adult ~ U{0..17}, no mutation clock, adaptation at the
 * floor (empty CSR), culture = one of 64 seed cultures chosen by 8x8 spatial block with 0..3 bit
 * flips.  hist_fill entries of history per family (0 = empty): patches within two offset steps of
 * the family, remembered harvests ~ U[0,0.4).  `out` arrays are caller-allocated
 * (hist arrays: n * hist_fill; adapt_off: n+1; adapt arrays unused). */
int dsim_synth_families(uint32_t W, uint32_t H, uint64_t seed, uint64_t n, uint64_t n_occupied,
                        uint32_t hist_fill, dsim_families *out);

#ifdef __cplusplus
}
#endif
#endif
