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
    uint32_t parity;   /* index of `cur` (0/1): ctl->n_cur[parity] families */
    uint32_t occ_par;  /* occ[occ_par] / n_occ[occ_par] collect this step's occupied nodes, [occ_par ^ 1] hold last step's */
    MgInfo mg;
    uint32_t mid_max;   /* nodes with small_max < n <= mid_max families take k_patch_mid (a warp per node) */
    uint32_t small_max; /* nodes with <= small_max families take the thread-per-node kernel (<= PATCH_SMALL_MAX) */
    uint32_t big_max;   /* nodes with mid_max < n <= big_max families take k_patch, larger ones k_patch_huge */
};

struct LaunchGeom {
    uint32_t n_sm;               /* grids are multiples of the SM count (148 on B200); kernels are grid-stride */
    uint32_t pdl;                /* bit k: kernel DSIM_K_<k> of a step is launched programmatically dependent on its predecessor */
};

/* Programmatic dependent launches were measured on B200 (profiles/r02_step_ab.md): inside a replayed CUDA graph the
 * kernel-to-kernel gaps are already ~1 us, PDL without an early trigger changes nothing (0.1805 vs 0.1815 ms per step
 * at C2) and with griddepcontrol.launch_dependents at the top of every kernel it costs 5-10 % (parked blocks of the
 * dependent kernel take SM resources from the tail of the running one).  Off by default; DSIM_PDL_MASK selects kernels. */
#define DSIM_PDL_DEFAULT_MASK 0x00u

/* one-time function attributes (none at present) */
int dsim_configure_kernels(LaunchGeom *g);

enum {
    DSIM_K_LIFECYCLE = 0,
    DSIM_K_MOVE,
    DSIM_K_CHILDREN,
    DSIM_K_SEGMENTS,
    DSIM_K_SCATTER,
    DSIM_K_PATCH_ONE,
    DSIM_K_PATCH_SMALL,
    DSIM_K_PATCH_MID,
    DSIM_K_PATCH,
    DSIM_K_PATCH_HUGE,
    DSIM_K_COUNT
};

static const char *const dsim_kernel_names[DSIM_K_COUNT] = {"k_lifecycle", "k_move",  "k_children", "k_segments",
                                                            "k_scatter",   "k_patch_one", "k_patch_small", "k_patch_mid", "k_patch", "k_patch_huge"};

enum { DSIM_MGK_SPLIT_IDS = 0, DSIM_MGK_SORT_OUT, DSIM_MGK_PACK, DSIM_MGK_UNPACK, DSIM_MGK_HALO_PACK, DSIM_MGK_HALO_UNPACK, DSIM_MGK_FINISH,
       DSIM_MGK_BUMP };
void dsim_mg_launch(int which, const StepArgs &a, const MgBuffers &mb, const LaunchGeom &g, cudaStream_t s);

void dsim_launch_kernel(int which, const StepArgs &a, const LaunchGeom &g, cudaStream_t s);
void dsim_launch_iota(uint32_t *v, uint32_t n, const LaunchGeom &g, cudaStream_t s);
void dsim_launch_init_views(const PatchState &ps, uint32_t n_nodes, double recovery, const LaunchGeom &g, cudaStream_t s);
void dsim_launch_census(const PatchState &ps, const FamCore &f, uint32_t n, uint32_t n_nodes, uint32_t *occ, DevCtl *ctl,
                        uint32_t parity, const LaunchGeom &g, cudaStream_t s);

#endif
