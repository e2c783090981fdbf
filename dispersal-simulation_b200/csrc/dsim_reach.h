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

#endif
