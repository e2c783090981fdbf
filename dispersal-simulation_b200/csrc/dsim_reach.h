/*
 * dsim_reach.h — one-time precomputation that replaces the per-family-per-step sensing of the
 * reference: `many_nearby_locations` (bounded Dijkstra, lib.rs:1133-1154, src/movementgraph.rs:64-115)
 * and `nearby_locations` (2-hop set, lib.rs:1158-1178, cached in `nearby_cache`, src/cli.rs:126).
 */
#ifndef DSIM_REACH_H
#define DSIM_REACH_H

#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/dsim.h"

struct HostRowHdr { uint32_t start, len, n_sensed, n_near; }; /* = RowHdr of dsim_state.h */

struct HostReach {
    std::vector<HostRowHdr> hdr;    /* [n] rows padded to multiples of 8 entries */
    std::vector<uint32_t> dst;      /* [sensed, ascending][others, ascending][padding 0xFFFFFFFF] */
    std::vector<double> cost;       /* OYR; padding +inf                             */
    std::vector<uint16_t> hash;     /* [n * 256] bucketized index of each row (rows of <= 192 entries), see dsim_state.h */
    std::vector<uint32_t> near_off; /* [n+1] */
    std::vector<uint32_t> near_dst; /* ascending within a node */
    uint64_t n_reach = 0;           /* entries without padding */
};

/* 13 hash bits of node q: bucket = bits 8..12, tag = bits 0..7 (dsim_state.h) */
static inline uint32_t dsim_row_hash(uint32_t q) { return (q * 2654435761u) >> 19; }
#define DSIM_ROW_HASH_MAX 192u

int dsim_build_reach(const dsim_graph *g, HostReach *out, std::string *err);

/* The same table built on the device (dsim_options.flags & DSIM_OPT_REACH_ON_DEVICE): a warp per node, two passes.
 * Pass 1 (fill = 0) writes per node {reach entries, sensed entries, nearby entries, 0} to `lens` and
 * {unsupported, largest |dst - node|, -, -} to `stats`; the caller lays the rows out (offsets, padding) and pass 2
 * (fill = 1) writes rows, nearby lists and bucket indices.  `unsupported` is raised for a node whose reach set or
 * 2-hop neighbourhood does not fit the per-warp tables (DSIM_REACH_DEV_MAX / DSIM_NEAR_DEV_MAX entries): the caller
 * then falls back to dsim_build_reach. */
#define DSIM_REACH_DEV_MAX 384u
#define DSIM_NEAR_DEV_MAX 512u
struct ReachDevGraph {
    const uint64_t *edge_off;
    const uint32_t *edge_dst;
    const double *edge_sec;
    uint32_t n;
    uint32_t v_lo, v_hi; /* the pass runs for the start nodes [v_lo, v_hi) (a band of a multi-GPU run, or everything) */
};
struct ReachDevOut {
    uint32_t *lens;  /* [4 n] */
    uint32_t *stats; /* [4]   */
    const HostRowHdr *hdr;
    uint32_t *dst;
    double *cost;
    uint16_t *hash;
    const uint32_t *near_off;
    uint32_t *near_dst;
};
void dsim_reach_device_pass(int fill, const ReachDevGraph &g, const ReachDevOut &o, uint32_t n_sm, void *stream);
double dsim_reach_cost_per_second(void);

#endif
