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

struct HostReach {
    std::vector<uint32_t> off;      /* [n+1], rows padded to multiples of 8 entries */
    std::vector<uint32_t> dst;      /* ascending within a row; padding 0xFFFFFFFF   */
    std::vector<double> cost;       /* OYR; padding +inf                             */
    std::vector<uint16_t> nidx;     /* position in the node's nearby list or 0xFFFF  */
    std::vector<uint32_t> near_off; /* [n+1] */
    std::vector<uint32_t> near_dst; /* ascending within a node */
    uint64_t n_reach = 0;           /* entries without padding */
};

int dsim_build_reach(const dsim_graph *g, HostReach *out, std::string *err);

#endif
