/*
 * dsim_kernels.h — kernel argument block and host-side launchers (dsim_kernels.cu).
 */
#ifndef DSIM_KERNELS_H
#define DSIM_KERNELS_H

#include "../../include/dsim.h"
#include "dsim_rt.h"
#include "dsim_state.h"

struct StepArgs {
    FamCore cur, next; /* family buffers: part 1 reads `cur` and writes `next`; part 2 updates `next` */
    FamHeavy hv;
    PatchState ps;
    ReachTable rt;
    StepScratch sc;
    DevCtl *ctl;
    ModelConst mc;
    uint32_t parity;   /* index of `cur` (0/1); also selects occ[parity] / n_occ[parity] */
};

struct LaunchGeom {
    uint32_t n_sm; /* grids are multiples of the SM count (148 on B200); kernels are grid-stride */
};

enum {
    DSIM_K_LIFECYCLE = 0,
    DSIM_K_CONTROL,
    DSIM_K_MOVE,
    DSIM_K_SEGMENTS,
    DSIM_K_SCATTER,
    DSIM_K_PATCH,
    DSIM_K_ADVANCE,
    DSIM_K_COUNT
};

static const char *const dsim_kernel_names[DSIM_K_COUNT] = {"k_lifecycle", "k_control", "k_move",  "k_segments",
                                                            "k_scatter",   "k_patch",   "k_advance"};

void dsim_launch_kernel(int which, const StepArgs &a, const LaunchGeom &g, cudaStream_t s);
void dsim_launch_init_views(const PatchState &ps, uint32_t n_nodes, double recovery, const LaunchGeom &g, cudaStream_t s);
void dsim_launch_census(const PatchState &ps, const FamCore &f, uint32_t n, uint32_t n_nodes, uint32_t *occ, DevCtl *ctl,
                        uint32_t parity, const LaunchGeom &g, cudaStream_t s);

#endif
