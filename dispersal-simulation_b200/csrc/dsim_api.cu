/*
 * dsim_api.cu — the C ABI of include/dsim.h: state ownership, host<->HBM conversion, step driver.
 *
 * Mirrors what `run` (src/cli.rs:123-156) holds between steps: State {families, t, p, patches}
 * (lib.rs:292-318), the `storage` map returned by `step`, and the `nearby_cache`.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "dsim_kernels.h"
#include "dsim_math.h"
#include "dsim_reach.h"

extern "C" void dsim_mg_release(dsim_handle *h);
extern "C" int dsim_mg_step_enqueue(dsim_handle *h);

namespace {

std::string g_create_error;

struct DevBuf { /* one cudaMalloc'd array */
    void *p = nullptr;
    size_t bytes = 0;
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char buf_[512];                                                                              \
            std::snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(h, DSIM_ERR_CUDA, buf_);                                                         \
        }                                                                                                \
    } while (0)

} // namespace

/* page-locked host memory for std::vector (the staging buffers of the getters) */
template <typename T> struct PinnedAlloc {
    using value_type = T;
    PinnedAlloc() = default;
    template <class U> PinnedAlloc(const PinnedAlloc<U> &) {}
    T *allocate(size_t n) {
        T *p = nullptr;
        if (cudaMallocHost(&p, n * sizeof(T)) != cudaSuccess) throw std::bad_alloc();
        return p;
    }
    void deallocate(T *p, size_t) { cudaFreeHost(p); }
    template <class U> bool operator==(const PinnedAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const PinnedAlloc<U> &) const { return false; }
};
template <typename T> using pinned_vector = std::vector<T, PinnedAlloc<T>>;

struct dsim_handle {
    dsim_params p;
    dsim_options opt;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr, stream4 = nullptr, stream5 = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr, join3_ev = nullptr, join4_ev = nullptr, join5_ev = nullptr;
#if !defined(DSIM_EMU)
    cudaAccessPolicyWindow l2_window = {}; /* views + band records pinned in the persisting part of L2 (num_bytes 0 = off) */
#endif
    LaunchGeom geom;
    uint32_t n_nodes = 0, n_eco = 0;
    std::vector<uint32_t> eco_off_h; /* [n+1] */
    uint64_t n_near = 0, n_reach = 0, n_reach_padded = 0;
    uint32_t max_row = 0;
    uint32_t reach_span = 0; /* largest |node - reachable or 2-hop node|, in node ids */
    /* device-built reach table: dsim_create runs the counting pass (row lengths, reach_span); the rows are laid out and
     * filled when the table is first needed (ensure_reach) — for the nodes of the own band and its halo only when the
     * handle has been configured for several GPUs by then */
    bool reach_pending = false;
    std::vector<void *> reach_tmp;      /* the graph on the device, kept for the fill pass */
    std::vector<uint32_t> reach_lens;   /* [4 n] of the counting pass */
    const uint64_t *rb_eoff = nullptr;
    const uint32_t *rb_edst = nullptr;
    const double *rb_esec = nullptr;
    uint32_t reach_lo = 0, reach_hi = 0; /* start nodes whose rows exist */
    bool mg_enabled = false;
    uint32_t mg_lo = 0, mg_hi = 0, mg_occ_prev = 0;

    std::vector<void *> allocs; /* everything to cudaFree */
    ReachTable rt;
    PatchState ps;
    FamCore fam[2];
    FamHeavy hv;
    StepScratch sc;
    DevCtl *ctl = nullptr;
    DevCtl *ctl_h = nullptr; /* pinned mirror */
    uint8_t *mem_hi8_d = nullptr;
    uint32_t *mem_lo32_d = nullptr;
    double *mem_tail_d = nullptr;
    ModelConst mc;
    std::vector<void *> fam_allocs; /* family-capacity dependent arrays */
    uint64_t cap = 0, hs_cap = 0;
    uint32_t hcap = 0, acap = 0;
    uint32_t parity = 0;
    bool mid_step = false;
    uint32_t steps_since_check = 0;
    uint32_t check_interval = 1;
    /* k_patch_huge (a whole SM's shared memory per block) joins the step when the host has seen a node of more than 512
     * families at one of its looks at the control block; until then the step has one launch less */
    bool huge_enabled = false; /* steps that may be enqueued before the live count is read again (host_check) */
    uint64_t n_known = 0; /* families at the last host check */

#if !defined(DSIM_EMU)
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
#endif
    bool graphs_valid = false;
    bool profiling = false;
    cudaEvent_t ev[DSIM_K_COUNT + 1] = {};
    cudaEvent_t tev[2] = {};
    double k_ms[DSIM_K_COUNT] = {};
    uint64_t k_launches[DSIM_K_COUNT] = {};
    uint64_t total_launches = 0;

    /* host snapshot for the getters: the record columns and history lengths (snap_valid) and, on demand, the
     * adaptation slots replayed to the current decay count (snap_adapt_valid); page-locked, so that the downloads are
     * plain DMA transfers */
    bool snap_valid = false, snap_adapt_valid = false;
    pinned_vector<uint64_t> s_id, s_culture;
    pinned_vector<double> s_stored, s_lh;
    pinned_vector<uint32_t> s_loc, s_size, s_till_adult, s_till_mut, s_misc, s_hs, s_hist_cnt;
    std::vector<uint32_t> s_hist_len;
    pinned_vector<uint16_t> s_a_eco;
    pinned_vector<double> s_a_val;
    pinned_vector<uint32_t> s_a_n;

    std::vector<uint32_t> storage_nodes; /* nodes named by dsim_set_storage */
    std::vector<double> lon_h, lat_h;    /* node coordinates (movementgraph.rs:8) for the observer */
    /* dsim_observe_population_counts keeps what it collected for the dsim_observe_population that follows */
    std::vector<uint32_t> obs_node;
    std::vector<uint64_t> obs_cult, obs_size;
    bool obs_valid = false;
    std::string err;
    /* grow-only staging for state uploads and downloads (dsim_set_families, the snapshots of the getters): a device
     * buffer for the caller's history / adaptation CSR arrays on their way into the rings, and page-locked host memory
     * for the packed core records — no cudaMalloc / cudaFree and no pageable copies on a repeated upload or download */
    std::vector<uint64_t> s_hoff;
    void *stage_dev = nullptr, *stage_host = nullptr;
    size_t stage_dev_bytes = 0, stage_host_bytes = 0;
};

namespace {

int fail(dsim_handle *h, int code, const char *msg) {
    if (h)
        h->err = msg;
    else
        g_create_error = msg;
    return code;
}

template <typename T> int dev_alloc(dsim_handle *h, T **out, size_t count, std::vector<void *> &list, bool zero = true) {
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(cudaMalloc(&p, bytes));
    if (zero) CK(cudaMemsetAsync(p, 0, bytes, h->stream));
    list.push_back(p);
    *out = (T *)p;
    return DSIM_OK;
}

template <typename T> int upload(dsim_handle *h, T *dst, const T *src, size_t count) {
    if (count) CK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return DSIM_OK;
}
template <typename T> int download(dsim_handle *h, T *dst, const T *src, size_t count) {
    if (count) CK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    return DSIM_OK;
}

/* index of the list of nodes occupied by part 2 of the last step */
uint32_t occ_prev_index(const dsim_handle *h) { return h->mg_enabled ? h->mg_occ_prev : (h->parity ^ 1u); }

int sync_ctl(dsim_handle *h) { /* device -> pinned host mirror */
    CK(cudaMemcpyAsync(h->ctl_h, h->ctl, sizeof(DevCtl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    /* the host sees the step as finished: bookkeeping the device would only apply at its next step is applied to the
     * mirror (every writer pushes the whole mirror back, so the two never disagree about what is still due) */
    dsim_apply_advance(h->ctl_h);
    return DSIM_OK;
}
int push_ctl(dsim_handle *h) {
    CK(cudaMemcpyAsync(h->ctl, h->ctl_h, sizeof(DevCtl), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

int stage_reserve(dsim_handle *h, bool device, size_t bytes) {
    void *&p = device ? h->stage_dev : h->stage_host;
    size_t &have = device ? h->stage_dev_bytes : h->stage_host_bytes;
    if (bytes <= have) return DSIM_OK;
    CK(cudaStreamSynchronize(h->stream));
    if (p) {
        if (device)
            cudaFree(p);
        else
            cudaFreeHost(p);
        p = nullptr;
        have = 0;
    }
    bytes += bytes / 8 + 4096;
    if (device) {
        if (cudaMalloc(&p, bytes) != cudaSuccess) return fail(h, DSIM_ERR_CUDA, "out of device memory (staging)");
    } else {
        if (cudaMallocHost(&p, bytes) != cudaSuccess) return fail(h, DSIM_ERR_CUDA, "out of page-locked host memory (staging)");
    }
    have = bytes;
    return DSIM_OK;
}
/* carves `count` elements of T out of a staging buffer at `*off` (256-byte aligned) */
template <typename T> T *stage_take(void *base, size_t *off, size_t count) {
    T *p = reinterpret_cast<T *>(static_cast<char *>(base) + *off);
    *off += (count * sizeof(T) + 255u) & ~(size_t)255u;
    return p;
}
template <typename T> size_t stage_size(size_t count) { return (count * sizeof(T) + 255u) & ~(size_t)255u; }

void free_list(std::vector<void *> &l) {
    for (void *p : l) cudaFree(p);
    l.clear();
}

void destroy_graphs(dsim_handle *h) {
#if !defined(DSIM_EMU)
    for (int k = 0; k < 2; ++k)
        if (h->gexec[k]) {
            cudaGraphExecDestroy(h->gexec[k]);
            h->gexec[k] = nullptr;
        }
#endif
    h->graphs_valid = false;
}

StepArgs make_args(dsim_handle *h, uint32_t parity) {
    StepArgs a;
    a.cur = h->fam[parity];
    a.next = h->fam[parity ^ 1u];
    a.hv = h->hv;
    a.ps = h->ps;
    a.rt = h->rt;
    a.sc = h->sc;
    a.ctl = h->ctl;
    a.mc = h->mc;
    a.parity = parity;
    a.occ_par = parity;
    std::memset(&a.mg, 0, sizeof a.mg);
    a.small_max = (h->opt.flags & DSIM_OPT_PATCH_WARP_ONLY) ? 0u : 8u;
    a.mid_max = (h->opt.flags & DSIM_OPT_PATCH_WARP_ONLY) ? 0u : 32u;
    /* k_patch's shared-memory capacity (PATCH_M); larger nodes take k_patch_huge once the step includes it */
    a.big_max = h->huge_enabled ? 512u : 0xFFFFFFFFu;
    return a;
}

/* family-capacity dependent allocations; preserves the contents of the first `keep` families / slots */
int alloc_families(dsim_handle *h, uint64_t new_cap) {
    if (new_cap > 0xFFFFFF00ull) return fail(h, DSIM_ERR_CAPACITY, "family capacity exceeds 2^32");
    new_cap = (new_cap + 255) & ~(uint64_t)255;
    std::vector<void *> fresh;
    FamCore nf[2];
    FamHeavy nh = h->hv;
    StepScratch ns;
    std::memset(&ns, 0, sizeof ns);
    const uint64_t old_cap = h->cap, old_hs = h->hs_cap;
    const uint64_t hs_cap = new_cap;
    int rc;
#define A(ptr, count)                                     \
    if ((rc = dev_alloc(h, &(ptr), (count), fresh)) != 0) { \
        free_list(fresh);                                 \
        return rc;                                        \
    }
    for (int b = 0; b < 2; ++b) {
        A(nf[b].key, new_cap) A(nf[b].place, new_cap) A(nf[b].stock, new_cap) A(nf[b].clock, new_cap)
    }
    A(nh.hist_cnt, hs_cap) A(nh.hist_patch, hs_cap * h->hcap) A(nh.hist_harv, hs_cap * h->hcap)
    A(nh.a_eco, hs_cap * h->acap) A(nh.a_cnt, hs_cap * h->acap) A(nh.a_val, hs_cap * h->acap)
    A(nh.decay, hs_cap) A(nh.parent, hs_cap) A(nh.birth_index, hs_cap)
    nh.hcap = h->hcap;
    nh.acap = h->acap;
    A(ns.alive_bits, new_cap / 32 + 8) A(ns.split_bits, new_cap / 32 + 8) A(ns.word_pre, new_cap / 32 + 8)
    A(ns.blk_alive, new_cap / 256 + 1) A(ns.blk_split, new_cap / 256 + 1)
    A(ns.dead_hs, new_cap) A(ns.free_hs, hs_cap) A(ns.fslot, new_cap) A(ns.members, new_cap)
    A(ns.occ[0], new_cap) A(ns.occ[1], new_cap) A(ns.plist_one, new_cap) A(ns.plist_small, new_cap) A(ns.plist_mid, new_cap) A(ns.plist_big, new_cap) A(ns.plist_huge, new_cap)
    A(ns.child_task, new_cap)
    A(ns.x_idx, new_cap) A(ns.x_tidx, new_cap) A(ns.x_size, new_cap) A(ns.x_band, new_cap)
    A(ns.x_bcnt, new_cap) A(ns.x_btot, new_cap)
    A(ns.x_key, new_cap) A(ns.x_fid, new_cap) A(ns.x_tfid, new_cap) A(ns.x_cult, new_cap) A(ns.x_bcult, new_cap)
    A(ns.x_stored, new_cap) A(ns.x_contrib, new_cap) A(ns.x_lh, new_cap) A(ns.x_beff, new_cap) A(ns.x_bcoef, new_cap)
    A(ns.x_balive, new_cap)
    uint64_t *new_band_cult;
    uint32_t *new_band_size;
    A(new_band_cult, 2 * new_cap) A(new_band_size, 2 * new_cap)
#undef A
    /* a_eco: all slots free */
    if (cudaMemsetAsync(nh.a_eco, 0xFF, hs_cap * h->acap * sizeof(uint16_t), h->stream) != cudaSuccess) {
        free_list(fresh);
        return fail(h, DSIM_ERR_CUDA, "cudaMemsetAsync failed");
    }
    if (old_cap) { /* grow: carry the live state over */
#define CP(dst, src, count, T) \
    if (cudaMemcpyAsync((dst), (src), (size_t)(count) * sizeof(T), cudaMemcpyDeviceToDevice, h->stream) != cudaSuccess) { \
        free_list(fresh);      \
        return fail(h, DSIM_ERR_CUDA, "device copy failed while growing"); \
    }
        for (int b = 0; b < 2; ++b) {
            CP(nf[b].key, h->fam[b].key, old_cap, FamKey) CP(nf[b].place, h->fam[b].place, old_cap, FamPlace)
            CP(nf[b].stock, h->fam[b].stock, old_cap, FamStock) CP(nf[b].clock, h->fam[b].clock, old_cap, FamClock)
        }
        CP(nh.hist_cnt, h->hv.hist_cnt, old_hs, uint32_t) CP(nh.hist_patch, h->hv.hist_patch, old_hs * h->hcap, uint32_t)
        CP(nh.hist_harv, h->hv.hist_harv, old_hs * h->hcap, double) CP(nh.a_eco, h->hv.a_eco, old_hs * h->acap, uint16_t)
        CP(nh.a_cnt, h->hv.a_cnt, old_hs * h->acap, uint32_t) CP(nh.a_val, h->hv.a_val, old_hs * h->acap, double)
        CP(nh.decay, h->hv.decay, old_hs, uint32_t) CP(nh.parent, h->hv.parent, old_hs, uint64_t)
        CP(nh.birth_index, h->hv.birth_index, old_hs, uint16_t)
        CP(ns.free_hs, h->sc.free_hs, old_hs, uint32_t)
        CP(ns.occ[0], h->sc.occ[0], old_cap, uint32_t) CP(ns.occ[1], h->sc.occ[1], old_cap, uint32_t)
        CP(ns.members, h->sc.members, old_cap, uint32_t) /* the visiting order of the next k_move */
        /* band cultures: [0, old_cap) written by k_patch, [old_cap, 2 old_cap) uploaded by dsim_set_storage; the
         * views hold absolute offsets, so both halves keep their positions */
        CP(new_band_cult, h->ps.band_cult, 2 * old_cap, uint64_t) CP(new_band_size, h->ps.band_size, 2 * old_cap, uint32_t)
#undef CP
    }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
        free_list(fresh);
        return fail(h, DSIM_ERR_CUDA, "synchronize failed while allocating families");
    }
    free_list(h->fam_allocs);
    h->fam_allocs.swap(fresh);
    h->fam[0] = nf[0];
    h->fam[1] = nf[1];
    h->hv = nh;
    h->sc = ns;
    h->ps.band_cult = new_band_cult;
    h->ps.band_size = new_band_size;
    h->cap = new_cap;
    h->hs_cap = hs_cap;
    destroy_graphs(h);
    return DSIM_OK;
}

int ensure_capacity(dsim_handle *h, uint64_t n_now) {
    if (h->mg_enabled) return DSIM_OK; /* fixed capacity per rank (dsim_options.family_capacity) */
    /* default: keep 30 % head room and double; with a capacity chosen by the caller (dsim_options.family_capacity,
     * e.g. because twice the population would not fit the device) grow late and by a quarter */
    const bool tight = h->opt.family_capacity != 0;
    if (tight ? (n_now * 100 <= h->cap * 92) : (n_now * 10 <= h->cap * 7)) return DSIM_OK;
    uint64_t want = std::max<uint64_t>(tight ? n_now + n_now / 4 + 1024 : 2 * n_now + 1024, h->cap);
    if (want == h->cap) return DSIM_OK;
    int rc = sync_ctl(h);
    if (rc) return rc;
    rc = alloc_families(h, want);
    if (rc) return rc;
    h->ctl_h->cap = (uint32_t)h->cap;
    h->ctl_h->hs_cap = (uint32_t)h->hs_cap;
    return push_ctl(h);
}

#if !defined(DSIM_EMU)
/* the persisting-L2 window of the handle on every kernel node of a captured step */
void apply_l2_window(dsim_handle *h, cudaGraph_t g) {
    if (h->l2_window.num_bytes == 0) return;
    size_t n_nodes = 0;
    if (cudaGraphGetNodes(g, nullptr, &n_nodes) != cudaSuccess || n_nodes == 0) return;
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    if (cudaGraphGetNodes(g, nodes.data(), &n_nodes) != cudaSuccess) return;
    cudaKernelNodeAttrValue av{};
    av.accessPolicyWindow = h->l2_window;
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nd, &t) == cudaSuccess && t == cudaGraphNodeTypeKernel)
            if (cudaGraphKernelNodeSetAttribute(nd, cudaKernelNodeAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
    }
}

/* The four patch kernels work on disjoint sets of nodes: they fork from the main stream onto three side streams and
 * join it again (also inside a stream capture, where the events become graph edges), so they run concurrently. */
int launch_patch_kernels(dsim_handle *h, const StepArgs &a) {
    CK(cudaEventRecord(h->fork_ev, h->stream));
    CK(cudaStreamWaitEvent(h->stream2, h->fork_ev, 0));
    CK(cudaStreamWaitEvent(h->stream3, h->fork_ev, 0));
    CK(cudaStreamWaitEvent(h->stream4, h->fork_ev, 0));
    CK(cudaStreamWaitEvent(h->stream5, h->fork_ev, 0));
    if (h->huge_enabled) dsim_launch_kernel(DSIM_K_PATCH_HUGE, a, h->geom, h->stream5); /* the few largest nodes are the longest pole: first */
    dsim_launch_kernel(DSIM_K_PATCH_SMALL, a, h->geom, h->stream);
    dsim_launch_kernel(DSIM_K_PATCH_ONE, a, h->geom, h->stream4);
    dsim_launch_kernel(DSIM_K_PATCH_MID, a, h->geom, h->stream3);
    dsim_launch_kernel(DSIM_K_PATCH, a, h->geom, h->stream2);
    CK(cudaEventRecord(h->join_ev, h->stream2));
    CK(cudaEventRecord(h->join3_ev, h->stream3));
    CK(cudaEventRecord(h->join4_ev, h->stream4));
    CK(cudaEventRecord(h->join5_ev, h->stream5));
    CK(cudaStreamWaitEvent(h->stream, h->join5_ev, 0));
    CK(cudaStreamWaitEvent(h->stream, h->join_ev, 0));
    CK(cudaStreamWaitEvent(h->stream, h->join3_ev, 0));
    CK(cudaStreamWaitEvent(h->stream, h->join4_ev, 0));
    for (int k = DSIM_K_PATCH_ONE; k <= (h->huge_enabled ? DSIM_K_PATCH_HUGE : DSIM_K_PATCH); ++k) h->k_launches[k] += 1;
    h->total_launches += h->huge_enabled ? 5 : 4;
    return DSIM_OK;
}

int build_graphs(dsim_handle *h) {
    /* one graph per buffer parity */
    for (uint32_t par = 0; par < 2; ++par) {
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const StepArgs a = make_args(h, par);
        /* k_children runs NEXT TO k_move (second stream): it waits, task by task, for the tag a parent's tile stores when
         * it has decided, and fills the SM resources k_move's tail frees; both are joined before k_segments */
        dsim_launch_kernel(DSIM_K_LIFECYCLE, a, h->geom, h->stream);
        CK(cudaEventRecord(h->fork_ev, h->stream));
        CK(cudaStreamWaitEvent(h->stream2, h->fork_ev, 0));
        dsim_launch_kernel(DSIM_K_MOVE, a, h->geom, h->stream);
        dsim_launch_kernel(DSIM_K_CHILDREN, a, h->geom, h->stream2);
        CK(cudaEventRecord(h->join_ev, h->stream2));
        CK(cudaStreamWaitEvent(h->stream, h->join_ev, 0));
        dsim_launch_kernel(DSIM_K_SEGMENTS, a, h->geom, h->stream);
        dsim_launch_kernel(DSIM_K_SCATTER, a, h->geom, h->stream);
        const uint64_t tl = h->total_launches; /* counted per replay, not per capture */
        uint64_t kl[DSIM_K_COUNT];
        std::memcpy(kl, h->k_launches, sizeof kl);
        int rc = launch_patch_kernels(h, a);
        h->total_launches = tl;
        std::memcpy(h->k_launches, kl, sizeof kl);
        if (rc) return rc;
        CK(cudaStreamEndCapture(h->stream, &g));
        apply_l2_window(h, g);
        CK(cudaGraphInstantiate(&h->gexec[par], g, 0));
        CK(cudaGraphDestroy(g));
    }
    h->graphs_valid = true;
    return DSIM_OK;
}
#endif

int launch_range(dsim_handle *h, int k0, int k1) { /* kernels [k0, k1) of one step, plain launches */
    const StepArgs a = make_args(h, h->parity);
    for (int k = k0; k < k1; ++k) {
        if (h->profiling) CK(cudaEventRecord(h->ev[k], h->stream));
        if (k == DSIM_K_PATCH_HUGE && !h->huge_enabled) continue; /* not part of the step yet (see dsim_handle::huge_enabled) */
        dsim_launch_kernel(k, a, h->geom, h->stream);
        h->k_launches[k] += 1;
        h->total_launches += 1;
    }
    if (h->profiling) {
        CK(cudaEventRecord(h->ev[k1], h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int k = k0; k < k1; ++k) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]));
            h->k_ms[k] += ms;
        }
    }
    CK(cudaGetLastError());
    return DSIM_OK;
}

int launch_step(dsim_handle *h) {
#if !defined(DSIM_EMU)
    if (h->opt.use_graphs && !h->profiling) {
        if (!h->graphs_valid) {
            int rc = build_graphs(h);
            if (rc) return rc;
        }
        CK(cudaGraphLaunch(h->gexec[h->parity], h->stream));
        const int n_k = h->huge_enabled ? DSIM_K_COUNT : DSIM_K_COUNT - 1; /* k_patch_huge is the last one */
        for (int k = 0; k < n_k; ++k) h->k_launches[k] += 1;
        h->total_launches += n_k;
        h->parity ^= 1u;
        return DSIM_OK;
    }
#endif
    int rc = launch_range(h, 0, DSIM_K_COUNT);
    if (rc) return rc;
    h->parity ^= 1u;
    return DSIM_OK;
}

uint32_t compute_acap(double minimum) {
    /* number of `* 0.95` applications that take the largest reachable value (1.0, the fixed point of
     * v -> 0.95 v + 0.05) down to the floor; a slot is live for at most that many applications */
    double v = 1.0;
    uint32_t k = 0;
    while (v != minimum && k < DSIM_N_ECOREGIONS) {
        v = dsim_adapt_decay(v, minimum);
        ++k;
    }
    uint32_t acap = k + 2;
    if (acap > DSIM_N_ECOREGIONS) acap = DSIM_N_ECOREGIONS;
    return (acap + 3u) & ~3u;
}

/* refresh the list of "occupied last step" nodes and the census from the current family buffer */
int recensus(dsim_handle *h) {
    int rc = sync_ctl(h);
    if (rc) return rc;
    const uint32_t prev = occ_prev_index(h);
    h->ctl_h->n_occ[prev] = 0;
    rc = push_ctl(h);
    if (rc) return rc;
    dsim_launch_census(h->ps, h->fam[h->parity], h->ctl_h->n_cur[h->parity], h->n_nodes, h->sc.occ[prev], h->ctl, prev,
                       h->geom, h->stream);
    h->total_launches += 3;
    CK(cudaGetLastError());
    return DSIM_OK;
}

int take_snapshot(dsim_handle *h, bool want_adapt);

/* The reach table built by the device kernels of dsim_reach.cu (DSIM_OPT_REACH_ON_DEVICE).  *built stays false when a
 * node does not fit the per-warp tables; the caller then builds the table on the host. */
int build_reach_device(dsim_handle *h, const dsim_graph *g, bool *built) {
    /* counting pass over all nodes; the fill pass follows in ensure_reach */
    *built = false;
    const uint32_t n = g->n_nodes;
    const uint64_t n_edge = g->edge_off[n];
    std::vector<void *> &tmp = h->reach_tmp;
    uint64_t *d_eoff;
    uint32_t *d_edst, *d_lens, *d_stats;
    double *d_esec;
    std::vector<void *> tmp2;
    int rc;
    if ((rc = dev_alloc(h, &d_eoff, (size_t)n + 1, tmp, false)) || (rc = dev_alloc(h, &d_edst, n_edge, tmp, false)) ||
        (rc = dev_alloc(h, &d_esec, n_edge, tmp, false)) || (rc = dev_alloc(h, &d_lens, (size_t)n * 4, tmp2)) ||
        (rc = dev_alloc(h, &d_stats, 4, tmp2)) || (rc = upload(h, d_eoff, g->edge_off, (size_t)n + 1)) ||
        (rc = upload(h, d_edst, g->edge_dst, n_edge)) || (rc = upload(h, d_esec, g->edge_sec, n_edge))) {
        free_list(tmp);
        free_list(tmp2);
        return rc;
    }
    ReachDevGraph dg{d_eoff, d_edst, d_esec, n, 0u, n};
    ReachDevOut out{d_lens, d_stats, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    dsim_reach_device_pass(0, dg, out, h->geom.n_sm, (void *)h->stream);
    h->reach_lens.assign((size_t)n * 4, 0u);
    uint32_t stats[4] = {0, 0, 0, 0};
    if ((rc = download(h, h->reach_lens.data(), d_lens, (size_t)n * 4)) || (rc = download(h, stats, d_stats, 4)) ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
        free_list(tmp);
        free_list(tmp2);
        return rc ? rc : fail(h, DSIM_ERR_CUDA, "reach table: the counting pass failed");
    }
    free_list(tmp2);
    if (stats[0] != 0u) { /* a reach set or 2-hop neighbourhood larger than the per-warp tables */
        free_list(tmp);
        h->reach_lens.clear();
        return DSIM_OK;
    }
    for (uint32_t v = 0; v < n; ++v)
        if (h->reach_lens[4 * (size_t)v + 2] >= 0xFFFFu) {
            free_list(tmp);
            return fail(h, DSIM_ERR_INVALID, "a node has 65535 or more nodes in its 2-hop neighbourhood");
        }
    h->rb_eoff = d_eoff;
    h->rb_edst = d_edst;
    h->rb_esec = d_esec;
    h->reach_span = std::max(h->reach_span, stats[1]);
    h->reach_pending = true;
    *built = true;
    return DSIM_OK;
}

/* Lays out and fills the rows of the device-built reach table the first time it is needed: all of them, or — on a handle
 * configured for several GPUs — those of the nodes [own_lo - halo, own_hi + halo), the only ones this rank reads
 * (its families sit on owned nodes; a child's costs are looked up in the row of its parent's new node, at most one
 * reach distance outside).  Memory and fill time per rank shrink with the number of bands. */
int ensure_reach(dsim_handle *h) {
    if (!h->reach_pending) return DSIM_OK;
    const uint32_t n = h->n_nodes;
    uint32_t lo = 0, hi = n;
    if (h->mg_enabled) {
        const uint32_t hw = 2u * h->reach_span;
        lo = h->mg_lo > hw ? h->mg_lo - hw : 0u;
        hi = h->mg_hi + hw < n ? h->mg_hi + hw : n;
    }
    const std::vector<uint32_t> &lens = h->reach_lens;
    int rc;
    /* layout: rows padded to multiples of 8 entries (16-byte aligned starts for every column), as dsim_build_reach */
    std::vector<HostRowHdr> hdr(n, HostRowHdr{0u, 0u, 0u, 0u});
    std::vector<uint32_t> near_off((size_t)n + 1, 0u);
    uint64_t ro = 0, no = 0, n_reach = 0;
    uint32_t max_row = 0;
    for (uint32_t v = 0; v < n; ++v) {
        near_off[v] = (uint32_t)no;
        if (v < lo || v >= hi) continue;
        const uint32_t r_len = lens[4 * (size_t)v], n_sensed = lens[4 * (size_t)v + 1], n_near = lens[4 * (size_t)v + 2];
        const uint32_t padded = (r_len + 7u) & ~7u;
        if (ro > 0xFFFFFFF0ull || no > 0xFFFFFFF0ull) break;
        hdr[v] = HostRowHdr{(uint32_t)ro, padded, n_sensed, n_near};
        ro += padded;
        no += n_near;
        n_reach += r_len;
        max_row = std::max(max_row, padded);
    }
    if (ro > 0xFFFFFFF0ull || no > 0xFFFFFFF0ull) return fail(h, DSIM_ERR_CAPACITY, "reach table exceeds 2^32 entries");
    near_off[n] = (uint32_t)no;
    RowHdr *d_hdr;
    uint16_t *d_hash;
    uint32_t *d_dst, *d_noff, *d_ndst;
    double *d_cost;
    if ((rc = dev_alloc(h, &d_hdr, n, h->allocs, false)) || (rc = dev_alloc(h, &d_dst, ro + 8, h->allocs, false)) ||
        (rc = dev_alloc(h, &d_cost, ro + 8, h->allocs, false)) || (rc = dev_alloc(h, &d_hash, (size_t)(hi - lo) * 256u + 8, h->allocs, false)) ||
        (rc = dev_alloc(h, &d_noff, (size_t)n + 1, h->allocs, false)) || (rc = dev_alloc(h, &d_ndst, no + 8, h->allocs, false)) ||
        (rc = upload(h, (HostRowHdr *)d_hdr, hdr.data(), n)) || (rc = upload(h, d_noff, near_off.data(), (size_t)n + 1)))
        return rc;
    /* the bucket index is addressed by node id: the pointer is biased so that node `lo` maps to the first bucket row */
    uint16_t *hash_biased = reinterpret_cast<uint16_t *>(reinterpret_cast<uintptr_t>(d_hash) - (uintptr_t)lo * 256u * sizeof(uint16_t));
    ReachDevGraph dg{h->rb_eoff, h->rb_edst, h->rb_esec, n, lo, hi};
    uint32_t *d_stats;
    std::vector<void *> tmp2;
    if ((rc = dev_alloc(h, &d_stats, 4, tmp2))) return rc;
    ReachDevOut out{nullptr, d_stats, (const HostRowHdr *)d_hdr, d_dst, d_cost, hash_biased, d_noff, d_ndst};
    dsim_reach_device_pass(1, dg, out, h->geom.n_sm, (void *)h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
        free_list(tmp2);
        return fail(h, DSIM_ERR_CUDA, "reach table: the fill pass failed");
    }
    free_list(tmp2);
    free_list(h->reach_tmp);
    h->reach_lens.clear();
    h->reach_lens.shrink_to_fit();
    h->reach_pending = false;
    h->reach_lo = lo;
    h->reach_hi = hi;
    h->n_near = no;
    h->n_reach = n_reach;
    h->n_reach_padded = ro;
    h->max_row = max_row;
    h->rt.hash = hash_biased;
    h->rt.hdr = d_hdr;
    h->rt.dst = d_dst;
    h->rt.cost = d_cost;
    h->rt.near_off = d_noff;
    h->rt.near_dst = d_ndst;
    if (std::getenv("DSIM_DEBUG")) std::fprintf(stderr, "[dsim] reach table filled on the device for the nodes [%u, %u): %llu entries, %llu nearby, span %u\n", lo, hi,
                                                (unsigned long long)n_reach, (unsigned long long)no, h->reach_span);
    return DSIM_OK;
}

} // namespace

/* kernels used only by the getters / setters ----------------------------------------------------- */
__global__ void k_digest32(const uint32_t *w, uint64_t n, unsigned long long *acc) {
    /* sum over i of mix(i, w[i]) mod 2^64: independent of the order in which threads add */
    uint64_t s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (i << 32) ^ (i >> 32) ^ ((uint64_t)w[i] * 0x9E3779B97F4A7C15ull);
        x ^= x >> 29;
        x *= 0xBF58476D1CE4E5B9ull;
        x ^= x >> 32;
        s += x;
    }
    atomicAdd(acc, (unsigned long long)s);
}

__global__ void k_pack_adaptation(FamHeavy hv, const FamPlace *place, uint32_t n, double minimum, uint16_t *out_eco,
                                  double *out_val, uint32_t *out_n) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < n; i += nth) {
        const uint32_t hs = place[i].hs;
        const uint32_t d = hv.decay[hs];
        uint32_t m = 0;
        for (uint32_t k = 0; k < hv.acap; ++k) {
            const uint32_t e = hv.a_eco[(size_t)hs * hv.acap + k];
            if (e == DSIM_ECO_EMPTY) continue;
            double v = hv.a_val[(size_t)hs * hv.acap + k];
            for (uint32_t r = d - hv.a_cnt[(size_t)hs * hv.acap + k]; r > 0u && v != minimum; --r) v = dsim_adapt_decay(v, minimum);
            if (v == minimum) continue;
            out_eco[(size_t)i * hv.acap + m] = (uint16_t)e;
            out_val[(size_t)i * hv.acap + m] = v;
            ++m;
        }
        out_n[i] = m;
    }
}

/* dsim_set_families: history CSR (oldest first) -> ring, newest min(len, hcap) entries at positions 0.. */
__global__ void k_unpack_history(FamHeavy hv, const uint64_t *off, const uint32_t *patch, const double *harv, uint32_t n,
                                 uint32_t n_nodes, uint32_t *bad) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = gw; i < n; i += nw) {
        const uint64_t a = off[i], b = off[i + 1];
        const uint32_t keep = (b - a) < hv.hcap ? (uint32_t)(b - a) : hv.hcap;
        for (uint32_t k = lane; k < keep; k += 32u) {
            const uint32_t q = patch[b - keep + k];
            if (q >= n_nodes) atomicOr(bad, 1u);
            hv.hist_patch[(size_t)i * hv.hcap + k] = q;
            hv.hist_harv[(size_t)i * hv.hcap + k] = harv[b - keep + k];
        }
        if (lane == 0) hv.hist_cnt[i] = keep;
    }
}

/* dsim_set_families: sparse adaptation CSR -> slots (entries at the floor are not stored) */
__global__ void k_unpack_adaptation(FamHeavy hv, const uint64_t *off, const uint32_t *eco, const double *val, uint32_t n,
                                    double minimum, uint32_t *bad) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < n; i += nth) {
        uint32_t m = 0;
        for (uint64_t k = off[i]; k < off[i + 1]; ++k) {
            if (eco[k] >= DSIM_N_ECOREGIONS) {
                atomicOr(bad, 2u);
                continue;
            }
            if (val[k] == minimum) continue;
            if (m >= hv.acap) {
                atomicOr(bad, 4u);
                break;
            }
            hv.a_eco[(size_t)i * hv.acap + m] = (uint16_t)eco[k];
            hv.a_cnt[(size_t)i * hv.acap + m] = 0u;
            hv.a_val[(size_t)i * hv.acap + m] = val[k];
            ++m;
        }
    }
}

__global__ void k_pack_history(FamHeavy hv, const uint32_t *hs_of, const uint64_t *off, uint32_t n, uint32_t *out_patch,
                               double *out_harv) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = gw; i < n; i += nw) {
        const uint32_t hs = hs_of[i];
        const uint32_t cnt = hv.hist_cnt[hs];
        const uint32_t len = cnt < hv.hcap ? cnt : hv.hcap;
        for (uint32_t k = lane; k < len; k += 32u) { /* chronological k: newest-first index len-1-k */
            const uint32_t pos = (cnt - 1u - (len - 1u - k)) % hv.hcap;
            out_patch[off[i] + k] = hv.hist_patch[(size_t)hs * hv.hcap + pos];
            out_harv[off[i] + k] = hv.hist_harv[(size_t)hs * hv.hcap + pos];
        }
    }
}

__global__ void k_pack_storage(PatchState ps, const uint32_t *nodes, const uint64_t *off, uint32_t n, uint32_t *out_node,
                               uint64_t *out_cult, uint64_t *out_size, uint32_t *out_nb) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t k = tid; k < n; k += nth) {
        const uint32_t p = nodes[k];
        const PatchView v = ps.view[p];
        const uint32_t stor_off = ps.bands[p].stor_off;
        if (out_nb) out_nb[k] = v.nb;
        if (off)
            for (uint32_t b = 0; b < v.nb; ++b) {
                out_node[off[k] + b] = p;
                out_cult[off[k] + b] = ps.band_cult[stor_off + b];
                out_size[off[k] + b] = ps.band_size[stor_off + b];
            }
    }
}

__global__ void k_apply_storage(PatchState ps, const uint32_t *nodes, const uint32_t *first, const uint32_t *count, uint32_t n,
                                uint32_t base, uint32_t *occ, DevCtl *ctl, uint32_t prev) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t k = tid; k < n; k += nth) {
        const uint32_t p = nodes[k];
        ps.view[p].nb = count[k];
        ps.bands[p].stor_off = base + first[k];
        ps.bands[p].cult0 = ps.band_cult[base + first[k]];
        /* make sure the node is on the "occupied last step" list so that the next step clears it */
        if (ps.view[p].pop == 0u) occ[atomicAdd(&ctl->n_occ[prev], 1u)] = p;
    }
}

__global__ void k_clear_storage(PatchState ps, uint32_t n_nodes) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t p = tid; p < n_nodes; p += nth) ps.view[p].nb = 0u;
}

namespace {

int take_snapshot(dsim_handle *h, bool want_adapt) {
    int rc;
    if (!h->snap_valid) {
        if ((rc = sync_ctl(h))) return rc;
        const uint32_t buf = h->mid_step ? (h->parity ^ 1u) : h->parity;
        const uint32_t n = h->mid_step ? h->ctl_h->n_next : h->ctl_h->n_cur[h->parity];
        const FamCore &f = h->fam[buf];
        h->snap_adapt_valid = false;
        h->s_id.resize(n); h->s_culture.resize(n); h->s_stored.resize(n); h->s_lh.resize(n);
        h->s_loc.resize(n); h->s_size.resize(n); h->s_till_adult.resize(n); h->s_till_mut.resize(n);
        h->s_misc.resize(n); h->s_hs.resize(n); h->s_hist_len.resize(n); h->s_hist_cnt.resize(h->hs_cap);
        /* the packed sub-records come down as they are (into page-locked staging memory) and are split into the
         * caller's columns here */
        if ((rc = stage_reserve(h, false, stage_size<FamKey>(n) + stage_size<FamPlace>(n) + stage_size<FamStock>(n) + stage_size<FamClock>(n) +
                                              stage_size<uint32_t>(h->hs_cap))))
            return rc;
        size_t so_ = 0;
        FamKey *key = stage_take<FamKey>(h->stage_host, &so_, n);
        FamPlace *place = stage_take<FamPlace>(h->stage_host, &so_, n);
        FamStock *stock = stage_take<FamStock>(h->stage_host, &so_, n);
        FamClock *clock = stage_take<FamClock>(h->stage_host, &so_, n);
        uint32_t *hcnt = stage_take<uint32_t>(h->stage_host, &so_, h->hs_cap);
        if ((rc = download(h, key, (const FamKey *)f.key, n))) return rc;
        if ((rc = download(h, place, (const FamPlace *)f.place, n))) return rc;
        if ((rc = download(h, stock, (const FamStock *)f.stock, n))) return rc;
        if ((rc = download(h, clock, (const FamClock *)f.clock, n))) return rc;
        if ((rc = download(h, hcnt, (const uint32_t *)h->hv.hist_cnt, h->hs_cap))) return rc; /* history lengths */
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(h, DSIM_ERR_CUDA, "synchronize failed in snapshot");
        std::copy(hcnt, hcnt + h->hs_cap, h->s_hist_cnt.begin());
        for (uint32_t i = 0; i < n; ++i) {
            h->s_id[i] = key[i].id; h->s_culture[i] = key[i].culture;
            h->s_stored[i] = stock[i].stored; h->s_lh[i] = stock[i].last_harvest;
            h->s_loc[i] = place[i].loc; h->s_size[i] = place[i].size; h->s_hs[i] = place[i].hs; h->s_misc[i] = place[i].misc;
            h->s_till_adult[i] = clock[i].till_adult; h->s_till_mut[i] = clock[i].till_mut;
        }
        for (uint32_t i = 0; i < n; ++i) h->s_hist_len[i] = std::min(h->s_hist_cnt[h->s_hs[i]], h->hcap);
        h->snap_valid = true;
    }
    if (want_adapt && !h->snap_adapt_valid) {
        /* adaptation, replayed to the current decay count */
        const uint32_t buf = h->mid_step ? (h->parity ^ 1u) : h->parity;
        const FamCore &f = h->fam[buf];
        const uint32_t n = (uint32_t)h->s_id.size();
        h->s_a_n.resize(n);
        h->s_a_eco.resize((size_t)n * h->acap); h->s_a_val.resize((size_t)n * h->acap);
        std::vector<void *> tmp;
        uint16_t *d_eco;
        double *d_val;
        uint32_t *d_n;
        if ((rc = dev_alloc(h, &d_eco, (size_t)n * h->acap, tmp, false)) || (rc = dev_alloc(h, &d_val, (size_t)n * h->acap, tmp, false)) ||
            (rc = dev_alloc(h, &d_n, n, tmp, false))) {
            free_list(tmp);
            return rc;
        }
        if (n) {
            DSIM_LAUNCH(k_pack_adaptation, h->geom.n_sm * 4, 256, 0, h->stream, h->hv, (const FamPlace *)f.place, n, h->p.minimum_adaptation,
                        d_eco, d_val, d_n);
            h->total_launches += 1;
        }
        rc = download(h, h->s_a_eco.data(), d_eco, (size_t)n * h->acap);
        if (!rc) rc = download(h, h->s_a_val.data(), d_val, (size_t)n * h->acap);
        if (!rc) rc = download(h, h->s_a_n.data(), d_n, n);
        if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed in snapshot");
        free_list(tmp);
        if (rc) return rc;
        h->snap_adapt_valid = true;
    }
    return DSIM_OK;
}

} // namespace

extern "C" {

int dsim_abi_version(void) { return DSIM_ABI_VERSION; }

void dsim_default_params(dsim_params *p) { /* src/parameters.rs:22-54 */
    std::memset(p, 0, sizeof *p);
    p->resource_recovery_per_season = 0.10;
    p->culture_mutation_rate = 6e-3;
    p->culture_dimensionality = 20;
    p->cooperation_threshold = 6;
    p->maximum_resources_one_adult_can_harvest = 1e50;
    p->evidence_needed = 0.1;
    p->payoff_std = 0.1;
    p->minimum_adaptation = 0.5;
    p->fight_deadliness = 1u << 31;
    p->enemy_discount = std::pow(0.5, 0.2);
    p->season_length_in_years = 1. / 6.;
    p->cooperation_inclusive = 0;
}

void dsim_default_options(dsim_options *o) {
    std::memset(o, 0, sizeof *o);
    o->use_graphs = 1;
}

const char *dsim_last_error(const dsim_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void dsim_destroy(dsim_handle *h) {
    if (!h) return;
    if (h->stream) cudaStreamSynchronize(h->stream);
    dsim_mg_release(h);
    destroy_graphs(h);
    for (auto &e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : h->tev)
        if (e) cudaEventDestroy(e);
    free_list(h->fam_allocs);
    free_list(h->allocs);
    if (h->ctl_h) cudaFreeHost(h->ctl_h);
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    if (h->join_ev) cudaEventDestroy(h->join_ev);
    if (h->join3_ev) cudaEventDestroy(h->join3_ev);
    if (h->join4_ev) cudaEventDestroy(h->join4_ev);
    if (h->stream4) cudaStreamDestroy(h->stream4);
    if (h->join5_ev) cudaEventDestroy(h->join5_ev);
    if (h->stream5) cudaStreamDestroy(h->stream5);
    free_list(h->reach_tmp);
    if (h->stage_dev) cudaFree(h->stage_dev);
    if (h->stage_host) cudaFreeHost(h->stage_host);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int dsim_create(const dsim_params *p, const dsim_graph *g, uint64_t seed, const dsim_options *opt, dsim_handle **out) {
    if (!p || !g || !out || !g->eco_off || !g->eco_id || !g->edge_off || !g->edge_dst || !g->edge_sec)
        return fail(nullptr, DSIM_ERR_INVALID, "null argument");
    if (g->n_nodes == 0 || g->n_nodes > 0xFFFFFF00u) return fail(nullptr, DSIM_ERR_INVALID, "bad node count");
    if (!(p->season_length_in_years > 0.0)) return fail(nullptr, DSIM_ERR_INVALID, "season length must be positive");
    if (!(p->minimum_adaptation >= 0.0)) return fail(nullptr, DSIM_ERR_INVALID, "minimum_adaptation must be >= 0");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0)
        return fail(nullptr, DSIM_ERR_NO_DEVICE, "no CUDA device: the step loop has no CPU fallback");
    dsim_options o;
    if (opt)
        o = *opt;
    else
        dsim_default_options(&o);
    if ((int)o.device >= n_dev) return fail(nullptr, DSIM_ERR_NO_DEVICE, "device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaSetDevice((int)o.device) != cudaSuccess || cudaGetDeviceProperties(&prop, (int)o.device) != cudaSuccess)
        return fail(nullptr, DSIM_ERR_NO_DEVICE, "cannot select the CUDA device");
    if (prop.major != 10)
        return fail(nullptr, DSIM_ERR_NO_DEVICE, "the kernels are built for sm_100a (B200) only");

    const uint32_t n = g->n_nodes;
    const uint64_t n_eco = g->eco_off[n], n_edge = g->edge_off[n];
    if (n_eco > 0xFFFFFF00ull) return fail(nullptr, DSIM_ERR_INVALID, "too many ecoregion entries");
    for (uint32_t v = 0; v < n; ++v) {
        if (g->eco_off[v + 1] < g->eco_off[v] || g->edge_off[v + 1] < g->edge_off[v])
            return fail(nullptr, DSIM_ERR_INVALID, "CSR offsets must not decrease");
        for (uint64_t k = g->eco_off[v]; k < g->eco_off[v + 1]; ++k) {
            if (g->eco_id[k] >= DSIM_N_ECOREGIONS) return fail(nullptr, DSIM_ERR_INVALID, "ecoregion id out of range");
            if (k > g->eco_off[v] && g->eco_id[k] <= g->eco_id[k - 1])
                return fail(nullptr, DSIM_ERR_INVALID, "ecoregion ids must ascend within a node");
        }
    }
    for (uint64_t e = 0; e < n_edge; ++e)
        if (g->edge_dst[e] >= n) return fail(nullptr, DSIM_ERR_INVALID, "edge target out of range");

    dsim_handle *h = new dsim_handle;
    h->p = *p;
    h->opt = o;
    h->seed = seed;
    h->n_nodes = n;
    h->n_eco = (uint32_t)n_eco;
    if (g->lon) h->lon_h.assign(g->lon, g->lon + n);
    if (g->lat) h->lat_h.assign(g->lat, g->lat + n);
    std::memset(&h->geom, 0, sizeof h->geom);
    h->geom.n_sm = (uint32_t)std::max(1, prop.multiProcessorCount);
    h->geom.pdl = (h->opt.flags & DSIM_OPT_NO_PDL) ? 0u : DSIM_PDL_DEFAULT_MASK;
    if (const char *e = std::getenv("DSIM_PDL_MASK")) h->geom.pdl = (uint32_t)std::strtoul(e, nullptr, 0); /* developer A/B */
    std::memset(&h->rt, 0, sizeof h->rt);
    std::memset(&h->ps, 0, sizeof h->ps);
    std::memset(&h->hv, 0, sizeof h->hv);
    std::memset(&h->sc, 0, sizeof h->sc);
    std::memset(h->fam, 0, sizeof h->fam);
    int rc = DSIM_OK;
#define TRY(expr)                     \
    if ((rc = (expr)) != DSIM_OK) {   \
        g_create_error = h->err;      \
        dsim_destroy(h);              \
        return rc;                    \
    }
#define TRYCUDA(call)                                                                     \
    if ((call) != cudaSuccess) {                                                          \
        g_create_error = std::string(#call) + " failed: " + cudaGetErrorString(cudaGetLastError()); \
        dsim_destroy(h);                                                                  \
        return DSIM_ERR_CUDA;                                                             \
    }
    TRYCUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {
        /* the second stream carries kernels that run NEXT TO a bigger one on the main stream (k_children next to k_move):
         * lowest priority, so that their blocks only take what the main stream's kernel leaves or frees */
        int least = 0, greatest = 0;
        TRYCUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        TRYCUDA(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, least));
    }
    TRYCUDA(cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking));
    TRYCUDA(cudaStreamCreateWithFlags(&h->stream4, cudaStreamNonBlocking));
    TRYCUDA(cudaStreamCreateWithFlags(&h->stream5, cudaStreamNonBlocking));
    TRYCUDA(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&h->join3_ev, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&h->join4_ev, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&h->join5_ev, cudaEventDisableTiming));
    for (auto &e : h->ev) TRYCUDA(cudaEventCreate(&e));
    for (auto &e : h->tev) TRYCUDA(cudaEventCreate(&e));
    TRYCUDA(cudaMallocHost(&h->ctl_h, sizeof(DevCtl)));
    std::memset(h->ctl_h, 0, sizeof(DevCtl));
    TRY(dev_alloc(h, &h->ctl, 1, h->allocs));

    /* reach table (replaces bounded_dijkstra per family per step and the nearby_cache) */
    bool reach_built = false;
#if defined(DSIM_EMU)
    const bool reach_on_device = (o.flags & DSIM_OPT_REACH_ON_DEVICE) != 0; /* the CPU debugging build keeps the host builder */
#else
    const bool reach_on_device = !(o.flags & DSIM_OPT_REACH_ON_HOST);
#endif
    if (reach_on_device) TRY(build_reach_device(h, g, &reach_built));
    if (!reach_built) {
        HostReach hr;
        std::string err;
        rc = dsim_build_reach(g, &hr, &err);
        if (rc != DSIM_OK) {
            g_create_error = err;
            dsim_destroy(h);
            return rc;
        }
        h->n_near = hr.near_dst.size();
        h->n_reach = hr.n_reach;
        h->n_reach_padded = hr.dst.size();
        h->max_row = 0;
        for (uint32_t v = 0; v < n; ++v) {
            h->max_row = std::max(h->max_row, hr.hdr[v].len);
            for (uint32_t e = hr.hdr[v].start; e < hr.hdr[v].start + hr.hdr[v].len; ++e)
                if (hr.dst[e] != 0xFFFFFFFFu) h->reach_span = std::max(h->reach_span, hr.dst[e] > v ? hr.dst[e] - v : v - hr.dst[e]);
            for (uint32_t e = hr.near_off[v]; e < hr.near_off[v + 1]; ++e)
                h->reach_span = std::max(h->reach_span, hr.near_dst[e] > v ? hr.near_dst[e] - v : v - hr.near_dst[e]);
        }
        RowHdr *hdr;
        uint16_t *hash;
        uint32_t *dst, *near_off, *near_dst;
        double *cost;
        static_assert(sizeof(RowHdr) == sizeof(HostRowHdr), "row header layouts must agree");
        TRY(dev_alloc(h, &hdr, hr.hdr.size(), h->allocs, false));
        TRY(dev_alloc(h, &dst, hr.dst.size() + 8, h->allocs, false));
        TRY(dev_alloc(h, &cost, hr.cost.size() + 8, h->allocs, false));
        TRY(dev_alloc(h, &hash, hr.hash.size(), h->allocs, false));
        TRY(upload(h, hash, hr.hash.data(), hr.hash.size()));
        TRY(dev_alloc(h, &near_off, hr.near_off.size(), h->allocs, false));
        TRY(dev_alloc(h, &near_dst, hr.near_dst.size(), h->allocs, false));
        TRY(upload(h, (HostRowHdr *)hdr, hr.hdr.data(), hr.hdr.size()));
        TRY(upload(h, dst, hr.dst.data(), hr.dst.size()));
        TRY(upload(h, cost, hr.cost.data(), hr.cost.size()));
        TRY(upload(h, near_off, hr.near_off.data(), hr.near_off.size()));
        TRY(upload(h, near_dst, hr.near_dst.data(), hr.near_dst.size()));
        TRYCUDA(cudaStreamSynchronize(h->stream));
        h->rt.hash = hash;
        h->reach_lo = 0;
        h->reach_hi = n;
        h->rt.hdr = hdr; h->rt.dst = dst; h->rt.cost = cost;
        h->rt.near_off = near_off; h->rt.near_dst = near_dst;
    }
    /* patches */
    {
        h->eco_off_h.resize((size_t)n + 1);
        for (uint32_t v = 0; v <= n; ++v) h->eco_off_h[v] = (uint32_t)g->eco_off[v];
        std::vector<uint16_t> eco16(n_eco);
        for (uint64_t k = 0; k < n_eco; ++k) eco16[k] = (uint16_t)g->eco_id[k];
        uint32_t *eco_off;
        uint16_t *eco_id;
        double *mx;
        /* views and band records in ONE allocation: the window of the persisting-L2 policy below covers both */
        {
            static_assert(sizeof(PatchView) == 16 && sizeof(PatchBands) == 16, "16-byte patch records");
            PatchView *vb = nullptr;
            TRY(dev_alloc(h, &vb, 2 * (size_t)n, h->allocs));
            h->ps.view = vb;
            h->ps.bands = reinterpret_cast<PatchBands *>(vb + n);
        }
#if !defined(DSIM_EMU)
        /* Developer knob DSIM_L2_PERSIST=1, off by default.  Every move decision gathers ~61 of these 16-byte records at
         * random (a third of k_move's DRAM sectors), and between two steps more than an L2 worth of reach rows and
         * history rings streams through the cache: with the knob the 32 n bytes of views and bands are pinned in the
         * persisting carve-out of L2 (access policy window on every stream of the handle and on every kernel node of
         * its graphs).  Measured on B200 (profiles/r02_k_move_ab.md): k_move does not get faster (115.4 vs 115.4 us at
         * C2, 855 vs 867 at C3: the kernel is not bound by those sectors), and the 61 MB taken from the rest of L2
         * slow the patch kernels down (C3: k_patch_small 75 -> 132 us). */
        h->l2_window = {};
        if (std::getenv("DSIM_L2_PERSIST")) {
            cudaDeviceProp prop{};
            int dev = 0;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 &&
                prop.accessPolicyMaxWindowSize > 0) {
                const size_t want = 2 * (size_t)n * sizeof(PatchView);
                const size_t window = std::min<size_t>(want, (size_t)prop.accessPolicyMaxWindowSize);
                const size_t carve = std::min<size_t>(window, (size_t)prop.persistingL2CacheMaxSize);
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                    h->l2_window.base_ptr = h->ps.view;
                    h->l2_window.num_bytes = window;
                    h->l2_window.hitRatio = (float)std::min(1.0, (double)carve / (double)window);
                    h->l2_window.hitProp = cudaAccessPropertyPersisting;
                    h->l2_window.missProp = cudaAccessPropertyStreaming;
                    cudaStreamAttrValue av{};
                    av.accessPolicyWindow = h->l2_window;
                    for (cudaStream_t st : {h->stream, h->stream2, h->stream3, h->stream4})
                        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
                    if (std::getenv("DSIM_DEBUG"))
                        std::fprintf(stderr, "[dsim] persisting L2: window %zu MB, carve-out %zu MB (device maximum %d MB), hit ratio %.2f\n", window >> 20,
                                     carve >> 20, prop.persistingL2CacheMaxSize >> 20, h->l2_window.hitRatio);
                } else {
                    (void)cudaGetLastError();
                }
            }
        }
#endif
        TRY(dev_alloc(h, &h->ps.cnt, n, h->allocs));
        TRY(dev_alloc(h, &h->ps.seg, n, h->allocs));
        TRY(dev_alloc(h, &eco_off, (size_t)n + 1, h->allocs, false));
        TRY(dev_alloc(h, &eco_id, n_eco, h->allocs, false));
        TRY(dev_alloc(h, &h->ps.res, n_eco, h->allocs));
        TRY(dev_alloc(h, &mx, n_eco, h->allocs));
        TRY(upload(h, eco_off, h->eco_off_h.data(), (size_t)n + 1));
        TRY(upload(h, eco_id, eco16.data(), n_eco));
        TRYCUDA(cudaStreamSynchronize(h->stream));
        h->ps.eco_off = eco_off;
        h->ps.eco_id = eco_id;
        h->ps.max_res = mx;
    }
    /* model constants */
    {
        ModelConst &mc = h->mc;
        std::memset(&mc, 0, sizeof mc);
        mc.recovery = p->resource_recovery_per_season;
        mc.season = p->season_length_in_years;
        mc.cap_harvest = p->maximum_resources_one_adult_can_harvest;
        mc.min_adapt = p->minimum_adaptation;
        mc.log1m_rate = dsim_log(1. - p->culture_mutation_rate);
        mc.threshold = p->cooperation_threshold;
        mc.inclusive = p->cooperation_inclusive;
        mc.deadliness = p->fight_deadliness;
        mc.adult_reset = dsim_f64_as_u32(15. / mc.season) + 1u;                    /* lib.rs:535,1878 */
        mc.grow_reset = dsim_f64_as_u32((2.85 - 1.35) / (1. - 0.488) / mc.season); /* lib.rs:1910,1916 */
        mc.n_nodes = n;
        mc.k0 = (uint32_t)seed;
        mc.k1 = (uint32_t)(seed >> 32);
        /* memory after n decays, lib.rs:847-864; examined while memory >= f64::EPSILON */
        std::vector<double> mem;
        const double decay = std::pow(0.5, mc.season / 2.0);
        double memory = 1.0;
        while (!(memory < std::numeric_limits<double>::epsilon()) && mem.size() < (1u << 20)) {
            mem.push_back(memory);
            memory *= decay;
        }
        mc.n_mem = (uint32_t)mem.size();
        /* entry n is kept iff U_n < memory_n (lib.rs:859) with the 40-bit uniform U_n = X_n 2^-40, i.e. iff the
         * integer X_n < C_n = ceil(memory_n 2^40); the kernels compare the top byte and the low word separately */
        const size_t n_pad = (mem.size() + 15u) & ~(size_t)15u;
        std::vector<uint8_t> hi8(n_pad, 0);
        std::vector<uint32_t> lo32(n_pad, 0);
        mc.mem_always = 0;
        mc.mem_head = 0;
        for (size_t k = 0; k < mem.size(); ++k) {
            const double c = std::ceil(mem[k] * 1099511627776.0);
            if (c >= 1099511627776.0) { /* memory_n >= 1: always kept (only n = 0) */
                if (k != mc.mem_always) {
                    g_create_error = "memory table: entries with memory >= 1 must form a prefix";
                    dsim_destroy(h);
                    return DSIM_ERR_INVALID;
                }
                mc.mem_always = (uint32_t)k + 1u;
                hi8[k] = 255;
                lo32[k] = 0xFFFFFFFFu;
            } else {
                const uint64_t ci = (uint64_t)c;
                hi8[k] = (uint8_t)(ci >> 32);
                lo32[k] = (uint32_t)ci;
            }
            if (hi8[k] != 0) mc.mem_head = (uint32_t)(k / 16u) + 1u;
        }
        /* stream v2: at most 16 head blocks; the entries after the head are sampled by skipping against the survival
         * products (same recurrence, same operation order as the oracle's table) */
        if (mc.mem_head > 16u) mc.mem_head = 16u;
        mc.mem_tail0 = 16u * mc.mem_head;
        std::vector<double> tail;
        {
            double survive = 1.0;
            for (size_t k = mc.mem_tail0; k < mem.size(); ++k) {
                survive = survive * (1.0 - mem[k]);
                tail.push_back(survive);
            }
        }
        mc.mem_tail_len = (uint32_t)tail.size();
        TRY(dev_alloc(h, &h->mem_tail_d, tail.size(), h->allocs, false));
        TRY(upload(h, h->mem_tail_d, tail.data(), tail.size()));
        TRY(dev_alloc(h, &h->mem_hi8_d, hi8.size(), h->allocs, false));
        TRY(upload(h, h->mem_hi8_d, hi8.data(), hi8.size()));
        TRY(dev_alloc(h, &h->mem_lo32_d, lo32.size(), h->allocs, false));
        TRY(upload(h, h->mem_lo32_d, lo32.data(), lo32.size()));
        TRYCUDA(cudaStreamSynchronize(h->stream));
        mc.mem_hi8 = h->mem_hi8_d;
        mc.mem_lo32 = h->mem_lo32_d;
        mc.mem_tail = h->mem_tail_d;
        if (dsim_configure_kernels(&h->geom) != 0) {
            g_create_error = "cannot configure the shared-memory size of k_move";
            dsim_destroy(h);
            return DSIM_ERR_CUDA;
        }
        const uint32_t hcap_exact = std::max(mc.n_mem, mc.adult_reset + 1u);
        if (o.history_capacity != 0 && o.history_capacity < hcap_exact && !(o.flags & DSIM_OPT_SHORT_HISTORY)) {
            g_create_error = "dsim_options.history_capacity is shorter than what the memory loop of lib.rs:854-866 can sample and a child "
                             "inherits (lib.rs:1841-1844): decisions would silently deviate from the reference once a history outgrows the "
                             "ring; pass 0 (exact) or set DSIM_OPT_SHORT_HISTORY";
            dsim_destroy(h);
            return DSIM_ERR_INVALID;
        }
        uint32_t hcap = o.history_capacity ? o.history_capacity : hcap_exact;
        h->hcap = (hcap + 3u) & ~3u;
        h->acap = compute_acap(mc.min_adapt);
    }
    TRY(alloc_families(h, o.family_capacity ? o.family_capacity : 1024));
    h->ctl_h->cap = (uint32_t)h->cap;
    h->ctl_h->hs_cap = (uint32_t)h->hs_cap;
    TRY(push_ctl(h));
    dsim_launch_init_views(h->ps, n, h->mc.recovery, h->geom, h->stream);
    h->total_launches += 1;
    TRYCUDA(cudaStreamSynchronize(h->stream));
#undef TRY
#undef TRYCUDA
    *out = h;
    return DSIM_OK;
}

int dsim_sync(dsim_handle *h) {
    if (!h) return DSIM_ERR_INVALID;
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

int dsim_set_patches(dsim_handle *h, const double *res, const double *max_res) {
    if (h) {
        int rc_ = ensure_reach(h);
        if (rc_) return rc_;
    }
    if (!h || !res || !max_res) return DSIM_ERR_INVALID;
    int rc;
    if ((rc = upload(h, h->ps.res, res, h->n_eco))) return rc;
    if ((rc = upload(h, const_cast<double *>(h->ps.max_res), max_res, h->n_eco))) return rc;
    dsim_launch_init_views(h->ps, h->n_nodes, h->mc.recovery, h->geom, h->stream);
    h->total_launches += 1;
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

int dsim_get_patches(dsim_handle *h, double *res, double *max_res) {
    if (!h) return DSIM_ERR_INVALID;
    int rc;
    if (res && (rc = download(h, res, h->ps.res, h->n_eco))) return rc;
    if (max_res && (rc = download(h, max_res, h->ps.max_res, h->n_eco))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

static int set_families_impl(dsim_handle *h, const dsim_families *f, uint64_t next_id_override, bool have_override) {
    if (!h || !f) return DSIM_ERR_INVALID;
    const uint64_t n = f->n;
    if (n && (!f->id || !f->culture || !f->location || !f->size || !f->stored)) return fail(h, DSIM_ERR_INVALID, "missing family array");
    if (n > 0xFFFFFF00ull) return fail(h, DSIM_ERR_CAPACITY, "too many families");
    {
        int rc_ = ensure_reach(h);
        if (rc_) return rc_;
    }
    for (uint64_t i = 0; i < n; ++i) {
        if (f->location[i] >= h->n_nodes) return fail(h, DSIM_ERR_INVALID, "family location out of range");
        if (i > 0 && f->id[i] <= f->id[i - 1]) return fail(h, DSIM_ERR_INVALID, "family ids must ascend");
    }
    CK(cudaStreamSynchronize(h->stream));
    int rc;
    uint64_t want = h->opt.family_capacity ? std::max<uint64_t>(h->opt.family_capacity, n + n / 4 + 256) : 2 * n + 1024;
    if (want > h->cap && h->mg_enabled)
        return fail(h, DSIM_ERR_CAPACITY, "multi-GPU handles have a fixed capacity: set dsim_options.family_capacity");
    if (want > h->cap) {
        free_list(h->fam_allocs); /* nothing to preserve */
        h->cap = h->hs_cap = 0;
        if ((rc = alloc_families(h, want))) return rc;
    }
    const uint32_t acap = h->acap;
    /* the caller's columns are packed into the four sub-records of the core record (dsim_state.h), in page-locked
     * staging memory: the uploads below run at the rate of the link and overlap the unpacking of the histories */
    const size_t host_bytes = stage_size<FamKey>(n) + stage_size<FamPlace>(n) + stage_size<FamStock>(n) + stage_size<FamClock>(n) +
                              stage_size<uint64_t>(n) + stage_size<uint16_t>(n) + stage_size<uint32_t>(1);
    if ((rc = stage_reserve(h, false, host_bytes))) return rc;
    size_t hoff_ = 0;
    FamKey *key = stage_take<FamKey>(h->stage_host, &hoff_, n);
    FamPlace *place = stage_take<FamPlace>(h->stage_host, &hoff_, n);
    FamStock *stock = stage_take<FamStock>(h->stage_host, &hoff_, n);
    FamClock *clk = stage_take<FamClock>(h->stage_host, &hoff_, n);
    uint64_t *parent = stage_take<uint64_t>(h->stage_host, &hoff_, n);
    uint16_t *birth = stage_take<uint16_t>(h->stage_host, &hoff_, n);
    uint32_t *bad_h = stage_take<uint32_t>(h->stage_host, &hoff_, 1);
    CK(cudaMemsetAsync(h->hv.hist_cnt, 0, h->hs_cap * sizeof(uint32_t), h->stream));
    CK(cudaMemsetAsync(h->hv.a_eco, 0xFF, h->hs_cap * acap * sizeof(uint16_t), h->stream));
    {
        /* history and adaptation travel as the caller's CSR arrays (through the device staging buffer) and are unpacked
         * on the device */
        const uint64_t nh = f->hist_off ? f->hist_off[n] : 0, na = f->adapt_off ? f->adapt_off[n] : 0;
        const size_t dev_bytes = stage_size<uint32_t>(1) + (nh ? stage_size<uint64_t>(n + 1) + stage_size<uint32_t>(nh) + stage_size<double>(nh) : 0) +
                                 (na ? stage_size<uint64_t>(n + 1) + stage_size<uint32_t>(na) + stage_size<double>(na) : 0);
        if ((rc = stage_reserve(h, true, dev_bytes))) return rc;
        size_t doff_ = 0;
        uint32_t *d_bad = stage_take<uint32_t>(h->stage_dev, &doff_, 1);
        CK(cudaMemsetAsync(d_bad, 0, sizeof(uint32_t), h->stream));
        if (n && nh) {
            uint64_t *d_off = stage_take<uint64_t>(h->stage_dev, &doff_, n + 1);
            uint32_t *d_patch = stage_take<uint32_t>(h->stage_dev, &doff_, nh);
            double *d_harv = stage_take<double>(h->stage_dev, &doff_, nh);
            if ((rc = upload(h, d_off, (const uint64_t *)f->hist_off, n + 1)) || (rc = upload(h, d_patch, (const uint32_t *)f->hist_patch, nh)) ||
                (rc = upload(h, d_harv, (const double *)f->hist_harvest, nh)))
                return rc;
            DSIM_LAUNCH(k_unpack_history, h->geom.n_sm * 8, 256, 0, h->stream, h->hv, (const uint64_t *)d_off, (const uint32_t *)d_patch,
                        (const double *)d_harv, (uint32_t)n, h->n_nodes, d_bad);
            h->total_launches += 1;
        }
        if (n && na) {
            uint64_t *d_off = stage_take<uint64_t>(h->stage_dev, &doff_, n + 1);
            uint32_t *d_eco = stage_take<uint32_t>(h->stage_dev, &doff_, na);
            double *d_val = stage_take<double>(h->stage_dev, &doff_, na);
            if ((rc = upload(h, d_off, (const uint64_t *)f->adapt_off, n + 1)) || (rc = upload(h, d_eco, (const uint32_t *)f->adapt_eco, na)) ||
                (rc = upload(h, d_val, (const double *)f->adapt_val, na)))
                return rc;
            DSIM_LAUNCH(k_unpack_adaptation, h->geom.n_sm * 4, 256, 0, h->stream, h->hv, (const uint64_t *)d_off, (const uint32_t *)d_eco,
                        (const double *)d_val, (uint32_t)n, h->p.minimum_adaptation, d_bad);
            h->total_launches += 1;
        }
        /* the packing of the core records runs on the host while the history arrays (95 % of the bytes) are on the link */
        for (uint64_t i = 0; i < n; ++i) {
            const bool clock = f->has_mutation_clock ? f->has_mutation_clock[i] != 0 : false;
            key[i].id = f->id[i];
            key[i].culture = f->culture[i];
            place[i].loc = f->location[i];
            place[i].size = f->size[i];
            place[i].hs = (uint32_t)i;
            place[i].misc = (uint32_t)(f->n_offspring ? f->n_offspring[i] : 0) | (clock ? DSIM_MISC_CLOCK : 0u);
            stock[i].stored = f->stored[i];
            stock[i].last_harvest = f->last_harvest ? f->last_harvest[i] : 0.0;
            clk[i].till_adult = f->till_adult ? f->till_adult[i] : 0;
            clk[i].till_mut = (clock && f->till_mutation) ? f->till_mutation[i] : 0;
            parent[i] = f->parent_id ? f->parent_id[i] : DSIM_NO_PARENT;
            birth[i] = f->birth_index ? f->birth_index[i] : 0;
        }
        h->parity = 0;
        h->mid_step = false;
        h->snap_valid = false;
        const FamCore &c = h->fam[0];
        if ((rc = upload(h, c.key, (const FamKey *)key, n)) || (rc = upload(h, c.place, (const FamPlace *)place, n)) ||
            (rc = upload(h, c.stock, (const FamStock *)stock, n)) || (rc = upload(h, c.clock, (const FamClock *)clk, n)) ||
            (rc = upload(h, h->hv.parent, (const uint64_t *)parent, n)) || (rc = upload(h, h->hv.birth_index, (const uint16_t *)birth, n)))
            return rc;
        rc = download(h, bad_h, (const uint32_t *)d_bad, 1);
        if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed in set_families");
        if (rc) return rc;
        const uint32_t bad = *bad_h;
        if (bad & 1u) return fail(h, DSIM_ERR_INVALID, "history node out of range");
        if (bad & 2u) return fail(h, DSIM_ERR_INVALID, "adaptation ecoregion out of range");
        if (bad & 4u) return fail(h, DSIM_ERR_CAPACITY, "more adapted ecoregions than adaptation slots");
    }
    CK(cudaMemsetAsync(h->hv.decay, 0, h->hs_cap * sizeof(uint32_t), h->stream));
    if ((rc = sync_ctl(h))) return rc;
    DevCtl &ctl = *h->ctl_h;
    const uint32_t t_keep = ctl.t, serial_keep = ctl.serial;
    std::memset(&ctl, 0, sizeof ctl);
    ctl.t = t_keep;
    ctl.serial = serial_keep; /* never reset: stale ChildTask tags of earlier steps must not match a later step */
    ctl.n_cur[0] = (uint32_t)n;
    ctl.hs_hwm = (uint32_t)n;
    ctl.next_id = have_override ? next_id_override : (n ? f->id[n - 1] + 1 : 0);
    ctl.cap = (uint32_t)h->cap;
    ctl.hs_cap = (uint32_t)h->hs_cap;
    if ((rc = push_ctl(h))) return rc;
    dsim_launch_iota(h->sc.members, (uint32_t)n, h->geom, h->stream); /* visiting order of the first k_move */
    h->total_launches += 1;
    h->n_known = n;
    h->steps_since_check = 0;
    if ((rc = recensus(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

int dsim_set_families(dsim_handle *h, const dsim_families *f) {
    if (!h || !f) return DSIM_ERR_INVALID;
    if (!h->mg_enabled) return set_families_impl(h, f, 0, false);
    /* multi-GPU: keep the families of the own band; ids continue after the largest id of the FULL vector */
    const uint64_t n = f->n;
    if (n && (!f->id || !f->culture || !f->location || !f->size || !f->stored)) return fail(h, DSIM_ERR_INVALID, "missing family array");
    std::vector<uint64_t> keep;
    for (uint64_t i = 0; i < n; ++i) {
        if (f->location[i] >= h->n_nodes) return fail(h, DSIM_ERR_INVALID, "family location out of range");
        if (i > 0 && f->id[i] <= f->id[i - 1]) return fail(h, DSIM_ERR_INVALID, "family ids must ascend");
        if (f->location[i] >= h->mg_lo && f->location[i] < h->mg_hi) keep.push_back(i);
    }
    const uint64_t m = keep.size();
    std::vector<uint64_t> id(m), parent(m), culture(m), hoff(m + 1, 0), aoff(m + 1, 0);
    std::vector<uint16_t> birth(m), noff(m);
    std::vector<uint32_t> loc(m), size(m), ta(m), tm(m), hp, ae;
    std::vector<uint8_t> clk(m);
    std::vector<double> stored(m), lh(m), hh, av;
    for (uint64_t k = 0; k < m; ++k) {
        const uint64_t i = keep[k];
        id[k] = f->id[i];
        parent[k] = f->parent_id ? f->parent_id[i] : DSIM_NO_PARENT;
        birth[k] = f->birth_index ? f->birth_index[i] : 0;
        culture[k] = f->culture[i];
        loc[k] = f->location[i];
        size[k] = f->size[i];
        stored[k] = f->stored[i];
        lh[k] = f->last_harvest ? f->last_harvest[i] : 0.0;
        ta[k] = f->till_adult ? f->till_adult[i] : 0;
        tm[k] = f->till_mutation ? f->till_mutation[i] : 0;
        clk[k] = f->has_mutation_clock ? f->has_mutation_clock[i] : 0;
        noff[k] = f->n_offspring ? f->n_offspring[i] : 0;
        if (f->hist_off) {
            hp.insert(hp.end(), f->hist_patch + f->hist_off[i], f->hist_patch + f->hist_off[i + 1]);
            hh.insert(hh.end(), f->hist_harvest + f->hist_off[i], f->hist_harvest + f->hist_off[i + 1]);
        }
        hoff[k + 1] = hp.size();
        if (f->adapt_off) {
            ae.insert(ae.end(), f->adapt_eco + f->adapt_off[i], f->adapt_eco + f->adapt_off[i + 1]);
            av.insert(av.end(), f->adapt_val + f->adapt_off[i], f->adapt_val + f->adapt_off[i + 1]);
        }
        aoff[k + 1] = ae.size();
    }
    hp.push_back(0); hh.push_back(0.0); ae.push_back(0); av.push_back(0.0);
    dsim_families g;
    g.n = m;
    g.id = id.data(); g.parent_id = parent.data(); g.birth_index = birth.data(); g.culture = culture.data();
    g.location = loc.data(); g.size = size.data(); g.stored = stored.data(); g.last_harvest = lh.data();
    g.till_adult = ta.data(); g.till_mutation = tm.data(); g.has_mutation_clock = clk.data(); g.n_offspring = noff.data();
    g.hist_off = hoff.data(); g.hist_patch = hp.data(); g.hist_harvest = hh.data();
    g.adapt_off = aoff.data(); g.adapt_eco = ae.data(); g.adapt_val = av.data();
    return set_families_impl(h, &g, n ? f->id[n - 1] + 1 : 0, true);
}

int dsim_family_counts(dsim_handle *h, uint64_t *n, uint64_t *n_hist, uint64_t *n_adapt) {
    if (!h) return DSIM_ERR_INVALID;
    int rc;
    if (!n_hist && !n_adapt) { /* the number of families alone needs no snapshot */
        if ((rc = sync_ctl(h))) return rc;
        if (n) *n = h->mid_step ? h->ctl_h->n_next : h->ctl_h->n_cur[h->parity];
        return DSIM_OK;
    }
    if ((rc = take_snapshot(h, n_adapt != nullptr))) return rc;
    uint64_t nh = 0, na = 0;
    for (size_t i = 0; i < h->s_id.size(); ++i) {
        nh += h->s_hist_len[i];
        if (n_adapt) na += h->s_a_n[i];
    }
    if (n) *n = h->s_id.size();
    if (n_hist) *n_hist = nh;
    if (n_adapt) *n_adapt = na;
    return DSIM_OK;
}

int dsim_get_families(dsim_handle *h, dsim_families *f) {
    if (!h || !f) return DSIM_ERR_INVALID;
    int rc = take_snapshot(h, f->adapt_off != nullptr);
    if (rc) return rc;
    const size_t n = h->s_id.size();
    f->n = n;
    const uint64_t *parent = nullptr;
    const uint16_t *birth = nullptr;
    if (f->parent_id || f->birth_index) { /* lineage columns of the heavy slots, through the page-locked staging buffer */
        if ((rc = stage_reserve(h, false, stage_size<uint64_t>(h->hs_cap) + stage_size<uint16_t>(h->hs_cap)))) return rc;
        size_t so_ = 0;
        uint64_t *pp = stage_take<uint64_t>(h->stage_host, &so_, h->hs_cap);
        uint16_t *bb = stage_take<uint16_t>(h->stage_host, &so_, h->hs_cap);
        if ((rc = download(h, pp, (const uint64_t *)h->hv.parent, h->hs_cap)) || (rc = download(h, bb, (const uint16_t *)h->hv.birth_index, h->hs_cap)))
            return rc;
        CK(cudaStreamSynchronize(h->stream));
        parent = pp;
        birth = bb;
    }
    std::vector<uint64_t> &hoff = h->s_hoff;
    hoff.assign(n + 1, 0);
    for (size_t i = 0; i < n; ++i) {
        if (f->id) f->id[i] = h->s_id[i];
        if (f->parent_id) f->parent_id[i] = parent[h->s_hs[i]];
        if (f->birth_index) f->birth_index[i] = birth[h->s_hs[i]];
        if (f->culture) f->culture[i] = h->s_culture[i];
        if (f->location) f->location[i] = h->s_loc[i];
        if (f->size) f->size[i] = h->s_size[i];
        if (f->stored) f->stored[i] = h->s_stored[i];
        if (f->last_harvest) f->last_harvest[i] = h->s_lh[i];
        if (f->till_adult) f->till_adult[i] = h->s_till_adult[i];
        const bool clock = (h->s_misc[i] & DSIM_MISC_CLOCK) != 0;
        if (f->till_mutation) f->till_mutation[i] = clock ? h->s_till_mut[i] : 0;
        if (f->has_mutation_clock) f->has_mutation_clock[i] = clock ? 1 : 0;
        if (f->n_offspring) f->n_offspring[i] = (uint16_t)(h->s_misc[i] & 0xFFFFu);
        hoff[i + 1] = hoff[i] + h->s_hist_len[i];
    }
    if (f->hist_off) {
        std::copy(hoff.begin(), hoff.end(), f->hist_off);
        if (hoff[n] && (f->hist_patch || f->hist_harvest)) {
            std::vector<void *> tmp;
            uint64_t *d_off;
            uint32_t *d_hs, *d_patch;
            double *d_harv;
            if ((rc = dev_alloc(h, &d_off, n + 1, tmp, false)) || (rc = dev_alloc(h, &d_hs, n, tmp, false)) ||
                (rc = dev_alloc(h, &d_patch, hoff[n], tmp, false)) || (rc = dev_alloc(h, &d_harv, hoff[n], tmp, false)) ||
                (rc = upload(h, d_off, hoff.data(), n + 1)) || (rc = upload(h, d_hs, h->s_hs.data(), n))) {
                free_list(tmp);
                return rc;
            }
            DSIM_LAUNCH(k_pack_history, h->geom.n_sm * 4, 256, 0, h->stream, h->hv, (const uint32_t *)d_hs, (const uint64_t *)d_off,
                        (uint32_t)n, d_patch, d_harv);
            h->total_launches += 1;
            if (f->hist_patch) rc = download(h, f->hist_patch, d_patch, hoff[n]);
            if (!rc && f->hist_harvest) rc = download(h, f->hist_harvest, d_harv, hoff[n]);
            if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed in get_families");
            free_list(tmp);
            if (rc) return rc;
        }
    }
    if (f->adapt_off) {
        uint64_t na = 0;
        std::vector<std::pair<uint32_t, double>> row;
        for (size_t i = 0; i < n; ++i) {
            f->adapt_off[i] = na;
            row.clear();
            for (uint32_t k = 0; k < h->s_a_n[i]; ++k) row.push_back({h->s_a_eco[i * h->acap + k], h->s_a_val[i * h->acap + k]});
            std::sort(row.begin(), row.end());
            for (const auto &ev : row) {
                if (f->adapt_eco) f->adapt_eco[na] = ev.first;
                if (f->adapt_val) f->adapt_val[na] = ev.second;
                ++na;
            }
        }
        f->adapt_off[n] = na;
    }
    return DSIM_OK;
}

int dsim_set_time(dsim_handle *h, uint64_t t) {
    if (!h) return DSIM_ERR_INVALID;
    int rc = sync_ctl(h);
    if (rc) return rc;
    h->ctl_h->t = (uint32_t)t;
    return push_ctl(h);
}

int dsim_get_time(dsim_handle *h, uint64_t *t) {
    if (!h || !t) return DSIM_ERR_INVALID;
    int rc = sync_ctl(h);
    if (rc) return rc;
    *t = h->ctl_h->t;
    return DSIM_OK;
}

static int collect_storage(dsim_handle *h, std::vector<uint32_t> &node, std::vector<uint64_t> &cult, std::vector<uint64_t> &size) {
    int rc = sync_ctl(h);
    if (rc) return rc;
    /* the nodes visited by part 2 of the last step are the "occupied last step" list of the next one */
    const uint32_t prev = occ_prev_index(h);
    const uint32_t n = h->ctl_h->n_occ[prev];
    node.clear(); cult.clear(); size.clear();
    if (n == 0) return DSIM_OK;
    std::vector<void *> tmp;
    uint32_t *d_nb, *d_node;
    uint64_t *d_off, *d_cult, *d_size;
    if ((rc = dev_alloc(h, &d_nb, n, tmp, false))) { free_list(tmp); return rc; }
    DSIM_LAUNCH(k_pack_storage, h->geom.n_sm * 4, 256, 0, h->stream, h->ps, (const uint32_t *)h->sc.occ[prev], (const uint64_t *)nullptr, n,
                (uint32_t *)nullptr, (uint64_t *)nullptr, (uint64_t *)nullptr, d_nb);
    std::vector<uint32_t> nb(n);
    rc = download(h, nb.data(), d_nb, n);
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed in storage");
    if (rc) { free_list(tmp); return rc; }
    std::vector<uint64_t> off(n + 1, 0);
    for (uint32_t k = 0; k < n; ++k) off[k + 1] = off[k] + nb[k];
    const uint64_t total = off[n];
    if (total) {
        if ((rc = dev_alloc(h, &d_off, n + 1, tmp, false)) || (rc = dev_alloc(h, &d_node, total, tmp, false)) ||
            (rc = dev_alloc(h, &d_cult, total, tmp, false)) || (rc = dev_alloc(h, &d_size, total, tmp, false)) ||
            (rc = upload(h, d_off, off.data(), n + 1))) { free_list(tmp); return rc; }
        DSIM_LAUNCH(k_pack_storage, h->geom.n_sm * 4, 256, 0, h->stream, h->ps, (const uint32_t *)h->sc.occ[prev], (const uint64_t *)d_off, n,
                    d_node, d_cult, d_size, (uint32_t *)nullptr);
        node.resize(total); cult.resize(total); size.resize(total);
        rc = download(h, node.data(), d_node, total);
        if (!rc) rc = download(h, cult.data(), d_cult, total);
        if (!rc) rc = download(h, size.data(), d_size, total);
        if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed in storage");
    }
    h->total_launches += 2;
    free_list(tmp);
    if (rc) return rc;
    /* map semantics of lib.rs:628-632: a later band with the same culture overwrites; sort by (node, culture) */
    std::vector<uint64_t> order(total);
    for (uint64_t k = 0; k < total; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
        return node[a] != node[b] ? node[a] < node[b] : cult[a] < cult[b];
    });
    std::vector<uint32_t> n2;
    std::vector<uint64_t> c2, s2;
    for (uint64_t r = 0; r < total; ++r) {
        const uint64_t k = order[r];
        if (!n2.empty() && n2.back() == node[k] && c2.back() == cult[k]) {
            s2.back() = size[k];
            continue;
        }
        n2.push_back(node[k]); c2.push_back(cult[k]); s2.push_back(size[k]);
    }
    node.swap(n2); cult.swap(c2); size.swap(s2);
    return DSIM_OK;
}

int dsim_observe_population_counts(dsim_handle *h, uint64_t *n_nodes, uint64_t *n_entries) {
    if (!h || !n_nodes || !n_entries) return DSIM_ERR_INVALID;
    int rc = collect_storage(h, h->obs_node, h->obs_cult, h->obs_size);
    if (rc) return rc;
    h->obs_valid = true;
    uint64_t nodes = 0;
    for (size_t k = 0; k < h->obs_node.size(); ++k)
        if (k == 0 || h->obs_node[k] != h->obs_node[k - 1]) ++nodes;
    *n_nodes = nodes;
    *n_entries = h->obs_node.size();
    return DSIM_OK;
}

int dsim_observe_population(dsim_handle *h, uint32_t *node, double *lon, double *lat, uint64_t *off, uint64_t *culture,
                            uint64_t *population) {
    if (!h || !off) return DSIM_ERR_INVALID;
    if (!h->obs_valid) {
        int rc = collect_storage(h, h->obs_node, h->obs_cult, h->obs_size);
        if (rc) return rc;
    }
    h->obs_valid = false;
    uint64_t k_node = 0;
    const size_t total = h->obs_node.size();
    if ((total && (!node || !culture || !population))) return DSIM_ERR_INVALID;
    for (size_t k = 0; k < total; ++k) {
        const uint32_t v = h->obs_node[k];
        if (k == 0 || v != h->obs_node[k - 1]) {
            node[k_node] = v;
            if (lon) lon[k_node] = h->lon_h.empty() ? 0.0 : h->lon_h[v];
            if (lat) lat[k_node] = h->lat_h.empty() ? 0.0 : h->lat_h[v];
            off[k_node] = k;
            ++k_node;
        }
        culture[k] = h->obs_cult[k];
        population[k] = h->obs_size[k];
    }
    off[k_node] = total;
    return DSIM_OK;
}

int dsim_storage_count(dsim_handle *h, uint64_t *n) {
    if (!h || !n) return DSIM_ERR_INVALID;
    std::vector<uint32_t> node;
    std::vector<uint64_t> cult, size;
    int rc = collect_storage(h, node, cult, size);
    if (rc) return rc;
    *n = node.size();
    return DSIM_OK;
}

int dsim_get_storage(dsim_handle *h, uint32_t *node, uint64_t *culture, uint64_t *size) {
    if (!h) return DSIM_ERR_INVALID;
    std::vector<uint32_t> nd;
    std::vector<uint64_t> cu, sz;
    int rc = collect_storage(h, nd, cu, sz);
    if (rc) return rc;
    if (node) std::copy(nd.begin(), nd.end(), node);
    if (culture) std::copy(cu.begin(), cu.end(), culture);
    if (size) std::copy(sz.begin(), sz.end(), size);
    return DSIM_OK;
}

int dsim_set_storage(dsim_handle *h, uint64_t n, const uint32_t *node, const uint64_t *culture, const uint64_t *size) {
    if (!h || (n && (!node || !culture))) return DSIM_ERR_INVALID;
    if (h->mid_step) return fail(h, DSIM_ERR_INVALID, "set_storage between the two parts of a step");
    if (n > h->cap) return fail(h, DSIM_ERR_CAPACITY, "more storage entries than family capacity");
    for (uint64_t i = 0; i < n; ++i)
        if (node[i] >= h->n_nodes) return fail(h, DSIM_ERR_INVALID, "storage node out of range");
    /* group by node, keeping insertion order within a node; a repeated (node, culture) overwrites */
    std::vector<uint64_t> order(n);
    for (uint64_t i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return node[a] < node[b]; });
    std::vector<uint32_t> nodes, first, count, sizes;
    std::vector<uint64_t> cults;
    for (uint64_t r = 0; r < n; ++r) {
        const uint64_t i = order[r];
        if (nodes.empty() || nodes.back() != node[i]) {
            nodes.push_back(node[i]);
            first.push_back((uint32_t)cults.size());
            count.push_back(0);
        }
        bool dup = false;
        for (uint32_t k = first.back(); k < cults.size(); ++k)
            if (cults[k] == culture[i]) {
                sizes[k] = (uint32_t)(size ? size[i] : 0);
                dup = true;
            }
        if (dup) continue;
        cults.push_back(culture[i]);
        sizes.push_back((uint32_t)(size ? size[i] : 0));
        count.back() += 1;
    }
    int rc;
    /* previous-step node list = census nodes (rebuilt) + storage-only nodes */
    if ((rc = recensus(h))) return rc;
    DSIM_LAUNCH(k_clear_storage, h->geom.n_sm * 4, 256, 0, h->stream, h->ps, h->n_nodes);
    h->total_launches += 1;
    const uint32_t base = (uint32_t)h->cap;
    if (!nodes.empty()) {
        std::vector<void *> tmp;
        uint32_t *d_nodes, *d_first, *d_count;
        if ((rc = dev_alloc(h, &d_nodes, nodes.size(), tmp, false)) || (rc = dev_alloc(h, &d_first, nodes.size(), tmp, false)) ||
            (rc = dev_alloc(h, &d_count, nodes.size(), tmp, false)) || (rc = upload(h, d_nodes, nodes.data(), nodes.size())) ||
            (rc = upload(h, d_first, first.data(), nodes.size())) || (rc = upload(h, d_count, count.data(), nodes.size())) ||
            (rc = upload(h, h->ps.band_cult + base, cults.data(), cults.size())) ||
            (rc = upload(h, h->ps.band_size + base, sizes.data(), sizes.size()))) {
            free_list(tmp);
            return rc;
        }
        const uint32_t prev = occ_prev_index(h);
        DSIM_LAUNCH(k_apply_storage, h->geom.n_sm * 4, 256, 0, h->stream, h->ps, (const uint32_t *)d_nodes, (const uint32_t *)d_first,
                    (const uint32_t *)d_count, (uint32_t)nodes.size(), base, h->sc.occ[prev], h->ctl, prev);
        h->total_launches += 1;
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed in set_storage");
        free_list(tmp);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

static int host_check(dsim_handle *h) { /* read the live count, grow if the buffers are getting full */
    int rc = sync_ctl(h);
    if (rc) return rc;
    const uint64_t n_prev = h->n_known;
    const uint32_t steps = h->steps_since_check;
    h->steps_since_check = 0;
    h->n_known = h->ctl_h->n_cur[h->parity];
    if (h->ctl_h->errors & DSIM_MODEL_CAPACITY) return fail(h, DSIM_ERR_CAPACITY, "family capacity exceeded during a batch of steps");
    if (!h->huge_enabled && h->ctl_h->n_over > 0) { /* a node has outgrown k_patch's shared memory: add k_patch_huge to the step */
        h->huge_enabled = true;
        destroy_graphs(h);
        if (std::getenv("DSIM_DEBUG")) std::fprintf(stderr, "[dsim] a node holds more than 512 families: k_patch_huge joins the step (t = %u)\n", h->ctl_h->t);
    }
    if ((rc = ensure_capacity(h, h->n_known))) return rc;
    /* How many steps may run before the next look at the live count: the growth per step seen since the last check,
     * squared for safety (1.25 while nothing has been seen: a family splits at most once per step, the usual rate is a
     * per cent), must not reach 90 % of the capacity.  A population that doubles per step is checked — and its buffers
     * doubled — every step; a steady one every 8 steps as before. */
    double g = 1.25;
    if (steps > 0 && n_prev > 0) {
        const double r = std::pow((double)std::max<uint64_t>(h->n_known, 1) / (double)n_prev, 1.0 / (double)steps);
        g = std::max(1.02, r * r);
    }
    if (!h->mg_enabled && h->opt.family_capacity == 0 && (double)h->n_known * g * g > 0.9 * (double)h->cap) {
        const uint64_t want = std::max<uint64_t>(2 * h->n_known + 1024, (uint64_t)((double)h->n_known * g * g * 1.5));
        if (want > h->cap) {
            if ((rc = alloc_families(h, want))) return rc;
            h->ctl_h->cap = (uint32_t)h->cap;
            h->ctl_h->hs_cap = (uint32_t)h->hs_cap;
            if ((rc = push_ctl(h))) return rc;
        }
    }
    uint32_t k = 1;
    for (double x = (double)h->n_known * g * g; k < 8u && x * g <= 0.9 * (double)h->cap; x *= g) ++k;
    h->check_interval = k;
    return DSIM_OK;
}

int dsim_step(dsim_handle *h, uint32_t n_steps, dsim_step_stats *stats) {
    if (!h) return DSIM_ERR_INVALID;
    if (h->mid_step) return fail(h, DSIM_ERR_INVALID, "dsim_step between the two parts of a step");
    int rc;
    if ((rc = ensure_reach(h))) return rc;
    h->snap_valid = false;
    h->obs_valid = false;
    uint64_t births0 = 0, deaths0 = 0, updates0 = 0;
    if (stats) {
        if ((rc = sync_ctl(h))) return rc;
        births0 = h->ctl_h->births; deaths0 = h->ctl_h->deaths; updates0 = h->ctl_h->updates;
    }
    for (uint32_t s = 0; s < n_steps; ++s) {
        if (h->steps_since_check >= h->check_interval || h->n_known * 10 > h->cap * 7)
            if ((rc = host_check(h))) return rc;
        if (h->mg_enabled) {
            if ((rc = dsim_mg_step_enqueue(h))) return rc;
        } else if ((rc = launch_step(h))) {
            return rc;
        }
        h->steps_since_check += 1;
    }
    if (!stats) return DSIM_OK;
    if ((rc = host_check(h))) return rc;
    const DevCtl &c = *h->ctl_h;
    std::memset(stats, 0, sizeof *stats);
    stats->t = c.t;
    stats->live_families = c.n_cur[h->parity];
    stats->births = c.births - births0;
    stats->deaths = c.deaths - deaths0;
    stats->family_updates = c.updates - updates0;
    stats->occupied_patches = c.n_occ[occ_prev_index(h)];
    stats->model_errors = c.errors;
    if (c.errors & ~(uint32_t)DSIM_MODEL_CAPACITY) {
        h->err = "a model assertion of the reference failed (see dsim_step_stats.model_errors)";
        return DSIM_ERR_MODEL;
    }
    return DSIM_OK;
}

int dsim_step_timed(dsim_handle *h, uint32_t n_steps, dsim_step_stats *stats, double *device_ms) {
    if (!h || !stats || !device_ms) return DSIM_ERR_INVALID;
    /* the counters at the entry of THIS call: births, deaths and family_updates cover exactly the n_steps timed steps,
     * whatever ran before (warm-up steps through dsim_step included) */
    int rc = sync_ctl(h);
    if (rc) return rc;
    const uint64_t b0 = h->ctl_h->births, d0 = h->ctl_h->deaths, u0 = h->ctl_h->updates;
    CK(cudaEventRecord(h->tev[0], h->stream));
    /* enqueue without the final read-back, then close the timed region on the device */
    rc = dsim_step(h, n_steps, nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(h->tev[1], h->stream));
    CK(cudaEventSynchronize(h->tev[1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->tev[0], h->tev[1]));
    *device_ms = ms;
    rc = dsim_step(h, 0, stats); /* reads the counters back; its own deltas are zero */
    stats->births = h->ctl_h->births - b0;
    stats->deaths = h->ctl_h->deaths - d0;
    stats->family_updates = h->ctl_h->updates - u0;
    return rc;
}

int dsim_step_part(dsim_handle *h, int part) {
    if (!h) return DSIM_ERR_INVALID;
    if (h->mg_enabled) return fail(h, DSIM_ERR_INVALID, "use dsim_mg_phase on a multi-GPU handle");
    int rc;
    if ((rc = ensure_reach(h))) return rc;
    h->snap_valid = false;
    if (part == 1) {
        if (h->mid_step) return fail(h, DSIM_ERR_INVALID, "part 1 twice");
        if ((rc = host_check(h))) return rc;
        if ((rc = launch_range(h, DSIM_K_LIFECYCLE, DSIM_K_SEGMENTS))) return rc;
        h->mid_step = true;
    } else if (part == 2) {
        if (!h->mid_step) return fail(h, DSIM_ERR_INVALID, "part 2 without part 1");
        if ((rc = launch_range(h, DSIM_K_SEGMENTS, DSIM_K_COUNT))) return rc;
        h->mid_step = false;
        h->parity ^= 1u;
        h->steps_since_check += 1;
    } else {
        return DSIM_ERR_INVALID;
    }
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

int dsim_reach_counts(dsim_handle *h, uint64_t *n_nearby, uint64_t *n_reach) {
    if (!h) return DSIM_ERR_INVALID;
    {
        int rc_ = ensure_reach(h);
        if (rc_) return rc_;
    }
    if (n_nearby) *n_nearby = h->n_near;
    if (n_reach) *n_reach = h->n_reach;
    return DSIM_OK;
}

int dsim_get_reach(dsim_handle *h, uint64_t *near_off, uint32_t *near_dst, uint64_t *reach_off, uint32_t *reach_dst,
                   double *reach_cost) {
    if (!h) return DSIM_ERR_INVALID;
    {
        int rc_ = ensure_reach(h);
        if (rc_) return rc_;
    }
    const uint32_t n = h->n_nodes;
    std::vector<HostRowHdr> hdr(n);
    std::vector<uint32_t> noff(n + 1), dst(h->n_reach_padded);
    std::vector<double> cost(h->n_reach_padded);
    int rc;
    if ((rc = download(h, hdr.data(), (const HostRowHdr *)h->rt.hdr, n)) || (rc = download(h, noff.data(), h->rt.near_off, n + 1)) ||
        (rc = download(h, dst.data(), h->rt.dst, h->n_reach_padded)) || (rc = download(h, cost.data(), h->rt.cost, h->n_reach_padded)))
        return rc;
    if (near_dst && (rc = download(h, near_dst, h->rt.near_dst, h->n_near))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    uint64_t w = 0;
    std::vector<std::pair<uint32_t, double>> row;
    for (uint32_t v = 0; v < n; ++v) {
        if (near_off) near_off[v] = noff[v];
        if (reach_off) reach_off[v] = w;
        row.clear(); /* merge the sensed-first layout back into one ascending list */
        for (uint32_t e = hdr[v].start; e < hdr[v].start + hdr[v].len; ++e)
            if (dst[e] != 0xFFFFFFFFu) row.push_back({dst[e], cost[e]});
        std::sort(row.begin(), row.end());
        for (const auto &qc : row) {
            if (reach_dst) reach_dst[w] = qc.first;
            if (reach_cost) reach_cost[w] = qc.second;
            ++w;
        }
    }
    if (near_off) near_off[n] = noff[n];
    if (reach_off) reach_off[n] = w;
    return DSIM_OK;
}

int dsim_reach_digest(dsim_handle *h, uint64_t out[5]) {
    if (!h || !out) return DSIM_ERR_INVALID;
    {
        int rc_ = ensure_reach(h);
        if (rc_) return rc_;
    }
    std::vector<void *> tmp;
    unsigned long long *acc = nullptr;
    int rc = dev_alloc(h, &acc, 5, tmp);
    if (rc) return rc;
    const struct { const void *p; uint64_t words; } part[5] = {
        {h->rt.hdr, (uint64_t)h->n_nodes * 4u},       {h->rt.dst, h->n_reach_padded},
        {h->rt.cost, h->n_reach_padded * 2u},         {h->rt.hash + (size_t)h->reach_lo * 256u, (uint64_t)(h->reach_hi - h->reach_lo) * 128u},
        {h->rt.near_dst, h->n_near}};
    for (int k = 0; k < 5; ++k)
        DSIM_LAUNCH(k_digest32, h->geom.n_sm * 8, 256, 0, h->stream, (const uint32_t *)part[k].p, part[k].words, acc + k);
    DSIM_LAUNCH(k_digest32, h->geom.n_sm * 8, 256, 0, h->stream, (const uint32_t *)h->rt.near_off, (uint64_t)h->n_nodes + 1u, acc + 4);
    h->total_launches += 6;
    unsigned long long host[5];
    rc = download(h, host, (const unsigned long long *)acc, 5);
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed in dsim_reach_digest");
    free_list(tmp);
    for (int k = 0; k < 5; ++k) out[k] = host[k];
    return rc;
}

int dsim_profile_enable(dsim_handle *h, int on) {
    if (!h) return DSIM_ERR_INVALID;
    CK(cudaStreamSynchronize(h->stream));
    h->profiling = on != 0;
    return DSIM_OK;
}

int dsim_profile_read(dsim_handle *h, dsim_kernel_times *out, int reset) {
    if (!h || !out) return DSIM_ERR_INVALID;
    out->n = DSIM_K_COUNT;
    for (int k = 0; k < DSIM_K_COUNT; ++k) {
        out->name[k] = dsim_kernel_names[k];
        out->ms[k] = h->k_ms[k];
        out->launches[k] = h->k_launches[k];
        if (reset) {
            h->k_ms[k] = 0.0;
            h->k_launches[k] = 0;
        }
    }
    return DSIM_OK;
}

int dsim_launch_count(dsim_handle *h, uint64_t *n) {
    if (!h || !n) return DSIM_ERR_INVALID;
    *n = h->total_launches;
    return DSIM_OK;
}

} /* extern "C" */

/* ================================================================================================ *
 * Multi-GPU: band partition (dsim_mg_kernels.cu).  One handle per GPU / process.
 * ================================================================================================ */
#if !defined(DSIM_EMU)
#include <nccl.h>
#endif

struct dsim_mg_state {
    bool enabled = false;
    uint32_t rank = 0, world = 1, own_lo = 0, own_hi = 0;
    MgBuffers mb;
    std::vector<void *> allocs;
    uint64_t split_bytes = 0, emig_bytes = 0, halo_bytes = 0;
    uint32_t occ_par = 0;
#if !defined(DSIM_EMU)
    ncclComm_t comm = nullptr;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
#endif
    bool nccl = false, graph_failed = false, p2p = false;
    bool graph_huge = false; /* the captured steps include k_patch_huge (dsim_handle::huge_enabled at capture time) */
    std::vector<void *> ipc_opened;
    cudaEvent_t seg_ev[8] = {};
    cudaEvent_t life_ev = nullptr, split_ev = nullptr; /* edges of the rotated step (mg_enqueue_step) */
    bool halo_pending = false; /* the last step's halo messages have arrived (or will) but are not unpacked yet */
    double seg_ms[7] = {};
    uint64_t seg_steps = 0;
    uint64_t launches_per_step = 0;
    /* per-kernel device times while profiling: event pairs around every launch of mg_launch_phase, read at step end */
    std::vector<cudaEvent_t> kev;
    std::vector<int> kev_id; /* >= 0: DSIM_K_*, < 0: -1 - DSIM_MGK_* */
    size_t kev_used = 0;
    double mgk_ms[DSIM_MGK_BUMP + 1] = {};
};

namespace {

dsim_mg_state *mg_of(dsim_handle *h);

StepArgs mg_args(dsim_handle *h, bool part2) {
    dsim_mg_state *m = mg_of(h);
    StepArgs a = make_args(h, h->parity);
    a.occ_par = m->occ_par;
    a.mg.enabled = 1;
    a.mg.rank = m->rank;
    a.mg.world = m->world;
    a.mg.own_lo = m->own_lo;
    a.mg.own_hi = m->own_hi;
    a.mg.split_send = m->mb.split_send;
    a.mg.split_recv = m->mb.split_recv;
    a.mg.scap = m->mb.scap;
    a.mg.p2p = m->mb.p2p;
    a.mg.seq = m->mb.seq;
    a.mg.flags = m->mb.flags;
    a.mg.counters = m->mb.counters;
    a.mg.emig_list = m->mb.emig_list;
    a.mg.ecap = m->mb.ecap;
    a.mg.halo_cult[0] = m->mb.halo_cult[0];
    a.mg.halo_cult[1] = m->mb.halo_cult[1];
    if (part2) { /* part 2 runs in place on the re-packed buffer */
        a.next = a.cur;
        a.sc.fslot = m->mb.fslot2;
    }
    return a;
}

} // namespace

/* the multi-GPU state hangs off the handle through a side table keyed by the handle pointer, so that
 * the single-GPU struct stays untouched */
#include <map>
namespace {
std::map<dsim_handle *, dsim_mg_state *> g_mg;
dsim_mg_state *mg_of(dsim_handle *h) {
    auto it = g_mg.find(h);
    return it == g_mg.end() ? nullptr : it->second;
}
} // namespace

extern "C" {

int dsim_mg_halo_width(dsim_handle *h, uint32_t *nodes) {
    if (!h || !nodes) return DSIM_ERR_INVALID;
    *nodes = 2u * h->reach_span;
    return DSIM_OK;
}

int dsim_mg_configure(dsim_handle *h, uint32_t rank, uint32_t world, const uint32_t *bounds) {
    if (!h || !bounds || world == 0 || rank >= world) return DSIM_ERR_INVALID;
    if (bounds[0] != 0 || bounds[world] != h->n_nodes) return fail(h, DSIM_ERR_INVALID, "bounds must cover [0, n_nodes]");
    const uint32_t hw = 2u * h->reach_span;
    for (uint32_t r = 0; r < world; ++r)
        if (bounds[r + 1] <= bounds[r] || (world > 1 && bounds[r + 1] - bounds[r] < hw))
            return fail(h, DSIM_ERR_INVALID, "every band must hold at least dsim_mg_halo_width() nodes");
    dsim_mg_state *m = mg_of(h);
    if (!m) {
        m = new dsim_mg_state;
        g_mg[h] = m;
    }
    free_list(m->allocs);
    std::memset(&m->mb, 0, sizeof m->mb);
    m->enabled = true;
    m->rank = rank;
    m->world = world;
    m->own_lo = bounds[rank];
    m->own_hi = bounds[rank + 1];
    m->occ_par = h->parity;
    h->mg_occ_prev = m->occ_par ^ 1u;
    MgBuffers &mb = m->mb;
    mb.has_peer[0] = rank > 0 ? 1u : 0u;
    mb.has_peer[1] = rank + 1 < world ? 1u : 0u;
    mb.scap = (uint32_t)std::max<uint64_t>(1024, h->cap / 8);
    mb.ecap = (uint32_t)std::max<uint64_t>(256, std::min<uint64_t>(h->cap / 64, 1024));
    if (const char *e = std::getenv("DSIM_MG_EMIGRANT_SLOTS")) mb.ecap = (uint32_t)std::max(16, std::atoi(e));
    mb.rec_bytes = mg_rec_bytes(h->acap, h->hcap);
    mb.hw = hw;
    mb.ccap = std::max<uint32_t>(1024u, hw / 2u);
    m->split_bytes = (1 + (uint64_t)mb.scap) * 8;
    m->emig_bytes = 16 + (uint64_t)mb.ecap * mb.rec_bytes;
    m->halo_bytes = 16 + (uint64_t)mb.hw * (sizeof(PatchView) + sizeof(PatchBands)) + (uint64_t)mb.ccap * 8;
    int rc;
#define MA(ptr, count)                                         \
    if ((rc = dev_alloc(h, &(ptr), (count), m->allocs)) != 0) { \
        free_list(m->allocs);                                  \
        return rc;                                             \
    }
    MA(mb.split_send, 1 + (size_t)mb.scap) MA(mb.split_recv, 2 * (size_t)world * (1 + (size_t)mb.scap))
    MA(mb.counters, MG_N_COUNTERS) MA(mb.emig_list, 2 * (size_t)mb.ecap)
    MA(mb.emig_hs, 2 * (size_t)mb.ecap) MA(mb.fslot2, h->cap)
    for (int s = 0; s < 2; ++s) {
        MA(mb.emig_send[s], m->emig_bytes) MA(mb.emig_recv[s], m->emig_bytes)
        MA(mb.halo_send[s], m->halo_bytes) MA(mb.halo_recv[s], m->halo_bytes)
        MA(mb.halo_cult[s], mb.ccap)
    }
    MA(mb.seq, 1) MA(mb.done, 4) MA(mb.flags, MG_FLAG_SPLIT0 + 2 * DSIM_MG_MAX_WORLD)
#undef MA
    CK(cudaStreamSynchronize(h->stream));
#if !defined(DSIM_EMU)
    for (auto &e : m->seg_ev)
        if (!e) CK(cudaEventCreate(&e));
    if (!m->life_ev) CK(cudaEventCreateWithFlags(&m->life_ev, cudaEventDisableTiming));
    if (!m->split_ev) CK(cudaEventCreateWithFlags(&m->split_ev, cudaEventDisableTiming));
#endif
    m->halo_pending = false;
    h->mg_enabled = true;
    h->mg_lo = m->own_lo;
    h->mg_hi = m->own_hi;
    destroy_graphs(h);
    return DSIM_OK;
}

int dsim_mg_buffer(dsim_handle *h, int channel, int peer, int recv, void **dev_ptr, uint64_t *bytes) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !dev_ptr || !bytes) return DSIM_ERR_INVALID;
    const MgBuffers &mb = m->mb;
    switch (channel) {
    case DSIM_MG_CH_SPLIT:
        if (recv && (peer < 0 || (uint32_t)peer >= m->world)) return DSIM_ERR_INVALID;
        *dev_ptr = recv ? (void *)(mb.split_recv + (size_t)peer * (1 + (size_t)mb.scap)) : (void *)mb.split_send;
        *bytes = m->split_bytes;
        return DSIM_OK;
    case DSIM_MG_CH_EMIGRANTS:
        if (peer < 0 || peer > 1) return DSIM_ERR_INVALID;
        *dev_ptr = recv ? mb.emig_recv[peer] : mb.emig_send[peer];
        *bytes = m->emig_bytes;
        return DSIM_OK;
    case DSIM_MG_CH_HALO:
        if (peer < 0 || peer > 1) return DSIM_ERR_INVALID;
        *dev_ptr = recv ? mb.halo_recv[peer] : mb.halo_send[peer];
        *bytes = m->halo_bytes;
        return DSIM_OK;
    default: return DSIM_ERR_INVALID;
    }
}

int dsim_mg_memcpy(dsim_handle *h, void *dst, const void *src, uint64_t bytes) {
    if (!h || !dst || !src) return DSIM_ERR_INVALID;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

enum { MG_PHASE_MOVE_A = 100, MG_PHASE_MOVE_B, MG_PHASE_MOVE_FORKED, MG_PHASE_ROTATED_HEAD, MG_PHASE_FINISH_HALO, MG_PHASE_FINISH_BOOK }; /* pieces of MOVE and FINISH for the overlapped schedule */

static int mg_launch_phase(dsim_handle *h, dsim_mg_state *m, int phase, bool fuse_finish = false) {
    const StepArgs a1 = mg_args(h, false), a2 = mg_args(h, true);
    MgBuffers mb = m->mb;
    mb.fuse_finish = fuse_finish ? 1u : 0u;
    auto mark = [&](int id) { /* profiling: an event before and after the launch */
        if (!h->profiling) return;
        if (m->kev_used == m->kev.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            m->kev.push_back(e);
            m->kev_id.push_back(0);
        }
        m->kev_id[m->kev_used] = id;
        cudaEventRecord(m->kev[m->kev_used++], h->stream);
    };
    auto K = [&](int k, const StepArgs &a) {
        mark(k);
        dsim_launch_kernel(k, a, h->geom, h->stream);
        mark(k);
        h->k_launches[k] += 1;
        h->total_launches += 1;
    };
    auto M = [&](int k, const StepArgs &a) {
        mark(-1 - k);
        dsim_mg_launch(k, a, mb, h->geom, h->stream);
        mark(-1 - k);
        h->total_launches += 1;
    };
    switch (phase) {
    case DSIM_MG_PHASE_LIFECYCLE:
        K(DSIM_K_LIFECYCLE, a1); M(DSIM_MGK_SPLIT_IDS, a1);
        break;
    case DSIM_MG_PHASE_MOVE: /* the children's ids (from the SPLIT exchange) are first needed by k_children */
        K(DSIM_K_MOVE, a1); K(DSIM_K_CHILDREN, a1); M(DSIM_MGK_SORT_OUT, a1); M(DSIM_MGK_PACK, a1);
        break;
    case MG_PHASE_MOVE_A: K(DSIM_K_MOVE, a1); break;
    case MG_PHASE_MOVE_B: K(DSIM_K_CHILDREN, a1); M(DSIM_MGK_SORT_OUT, a1); M(DSIM_MGK_PACK, a1); break;
    case MG_PHASE_ROTATED_HEAD:
        /* second stream: the halo of the PREVIOUS step is unpacked (after waiting for the neighbours' flags) next to
         * k_lifecycle, then k_mg_split_ids sends the SPLIT lists next to k_move */
        CK(cudaEventRecord(h->fork_ev, h->stream));
        CK(cudaStreamWaitEvent(h->stream2, h->fork_ev, 0));
        dsim_mg_launch(DSIM_MGK_HALO_UNPACK, a2, mb, h->geom, h->stream2);
        CK(cudaEventRecord(h->join_ev, h->stream2));
        K(DSIM_K_LIFECYCLE, a1);
        CK(cudaEventRecord(m->life_ev, h->stream));
        CK(cudaStreamWaitEvent(h->stream2, m->life_ev, 0));
        dsim_mg_launch(DSIM_MGK_SPLIT_IDS, a1, mb, h->geom, h->stream2);
        /* k_children follows on the second stream, next to k_move: it waits for the SPLIT flags of all ranks and then,
         * task by task, for the tag a parent's tile stores when it has decided (ChildTask.ready) */
        dsim_launch_kernel(DSIM_K_CHILDREN, a1, h->geom, h->stream2);
        h->k_launches[DSIM_K_CHILDREN] += 1;
        CK(cudaEventRecord(m->split_ev, h->stream2));
        h->total_launches += 3;
        CK(cudaStreamWaitEvent(h->stream, h->join_ev, 0)); /* k_move reads the halo views */
        K(DSIM_K_MOVE, a1);
        CK(cudaStreamWaitEvent(h->stream, m->split_ev, 0)); /* the children are written: k_mg_sort_out / k_mg_pack see every family */
        CK(cudaEventRecord(h->fork_ev, h->stream));
        CK(cudaStreamWaitEvent(h->stream2, h->fork_ev, 0));
        dsim_mg_launch(DSIM_MGK_SORT_OUT, a1, mb, h->geom, h->stream2);
        h->total_launches += 1;
        CK(cudaEventRecord(h->join_ev, h->stream2));
        M(DSIM_MGK_PACK, a1);
        CK(cudaStreamWaitEvent(h->stream, h->join_ev, 0));
        break;
    case MG_PHASE_MOVE_FORKED: /* k_mg_sort_out (re-packs the stayers) on the second stream next to k_mg_pack (sends the emigrants) */
        K(DSIM_K_MOVE, a1); K(DSIM_K_CHILDREN, a1);
        CK(cudaEventRecord(h->fork_ev, h->stream));
        CK(cudaStreamWaitEvent(h->stream2, h->fork_ev, 0));
        dsim_mg_launch(DSIM_MGK_SORT_OUT, a1, mb, h->geom, h->stream2);
        h->total_launches += 1;
        CK(cudaEventRecord(h->join_ev, h->stream2));
        M(DSIM_MGK_PACK, a1);
        CK(cudaStreamWaitEvent(h->stream, h->join_ev, 0));
        break;
    case MG_PHASE_FINISH_HALO: /* on the second stream, next to MG_PHASE_FINISH_BOOK */
        dsim_mg_launch(DSIM_MGK_HALO_UNPACK, a2, mb, h->geom, h->stream2);
        h->total_launches += 1;
        break;
    case MG_PHASE_FINISH_BOOK: M(DSIM_MGK_FINISH, a2); break;
    case DSIM_MG_PHASE_PATCH:
        M(DSIM_MGK_UNPACK, a2); K(DSIM_K_SEGMENTS, a2); K(DSIM_K_SCATTER, a2);
#if defined(DSIM_EMU)
        for (int k = DSIM_K_PATCH_ONE; k <= DSIM_K_PATCH_HUGE; ++k) K(k, a2);
#else
        {
            int rc = launch_patch_kernels(h, a2);
            if (rc) return rc;
        }
#endif
        M(DSIM_MGK_HALO_PACK, a2);
        break;
    case DSIM_MG_PHASE_FINISH:
        M(DSIM_MGK_HALO_UNPACK, a2); M(DSIM_MGK_FINISH, a2);
        break;
    case DSIM_MG_PHASE_HALO_PACK:
        CK(cudaMemsetAsync(mb.counters, 0, MG_N_COUNTERS * sizeof(uint32_t), h->stream));
        M(DSIM_MGK_HALO_PACK, a2);
        break;
    case DSIM_MG_PHASE_HALO_UNPACK:
        M(DSIM_MGK_HALO_UNPACK, a2);
        break;
    default: return DSIM_ERR_INVALID;
    }
    CK(cudaGetLastError());
    return DSIM_OK;
}

int dsim_mg_phase(dsim_handle *h, int phase) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m) return DSIM_ERR_INVALID;
    h->snap_valid = false;
    int rc = ensure_reach(h);
    if (rc) return rc;
    if (m->halo_pending) { /* a step of the library's own (rotated) schedule left its halo to the next one */
        if ((rc = mg_launch_phase(h, m, DSIM_MG_PHASE_HALO_UNPACK))) return rc;
        m->halo_pending = false;
    }
    rc = mg_launch_phase(h, m, phase);
    if (rc) return rc;
    if (phase == DSIM_MG_PHASE_FINISH) { /* the occupied-node lists swap roles */
        m->occ_par ^= 1u;
        h->mg_occ_prev = m->occ_par ^ 1u;
    }
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

int dsim_mg_unique_id(void *out) {
#if defined(DSIM_EMU)
    (void)out;
    return DSIM_ERR_COMM;
#else
    if (!out) return DSIM_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) <= DSIM_MG_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return DSIM_ERR_COMM;
    std::memset(out, 0, DSIM_MG_UNIQUE_ID_BYTES);
    std::memcpy(out, &id, sizeof id);
    return DSIM_OK;
#endif
}

int dsim_mg_init_nccl(dsim_handle *h, const void *unique_id) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !unique_id) return DSIM_ERR_INVALID;
#if defined(DSIM_EMU)
    return fail(h, DSIM_ERR_COMM, "no NCCL in the CPU debugging build");
#else
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof id);
    if (ncclCommInitRank(&m->comm, (int)m->world, id, (int)m->rank) != ncclSuccess) return fail(h, DSIM_ERR_COMM, "ncclCommInitRank failed");
    m->nccl = true;
    return DSIM_OK;
#endif
}

int dsim_mg_ipc_export(dsim_handle *h, void *out) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !out) return DSIM_ERR_INVALID;
#if defined(DSIM_EMU)
    return fail(h, DSIM_ERR_COMM, "no peer memory in the CPU debugging build");
#else
    static_assert(6 * sizeof(cudaIpcMemHandle_t) == DSIM_MG_IPC_BYTES, "handle block size");
    cudaIpcMemHandle_t *hd = reinterpret_cast<cudaIpcMemHandle_t *>(out);
    void *bufs[6] = {m->mb.emig_recv[0], m->mb.emig_recv[1], m->mb.halo_recv[0], m->mb.halo_recv[1], m->mb.split_recv, m->mb.flags};
    for (int k = 0; k < 6; ++k) CK(cudaIpcGetMemHandle(&hd[k], bufs[k]));
    return DSIM_OK;
#endif
}

int dsim_mg_ipc_connect(dsim_handle *h, const void *all) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !all) return DSIM_ERR_INVALID;
#if defined(DSIM_EMU)
    return fail(h, DSIM_ERR_COMM, "no peer memory in the CPU debugging build");
#else
    if (m->world > DSIM_MG_MAX_WORLD) return fail(h, DSIM_ERR_INVALID, "peer-memory transport: at most 8 ranks");
    MgBuffers &mb = m->mb;
    const cudaIpcMemHandle_t *hd = reinterpret_cast<const cudaIpcMemHandle_t *>(all);
    auto open = [&](uint32_t rank, int k, void **out) -> int {
        CK(cudaIpcOpenMemHandle(out, hd[rank * 6 + k], cudaIpcMemLazyEnablePeerAccess));
        m->ipc_opened.push_back(*out);
        return DSIM_OK;
    };
    int rc;
    for (uint32_t r = 0; r < m->world; ++r) {
        if (r == m->rank) {
            mb.peer_split_recv[r] = mb.split_recv;
            mb.peer_flags[r] = mb.flags;
            continue;
        }
        void *p = nullptr;
        if ((rc = open(r, 4, &p))) return rc;
        mb.peer_split_recv[r] = (uint64_t *)p;
        if ((rc = open(r, 5, &p))) return rc;
        mb.peer_flags[r] = (uint32_t *)p;
    }
    for (int s = 0; s < 2; ++s) {
        if (!mb.has_peer[s]) continue;
        const uint32_t r = s == 0 ? m->rank - 1u : m->rank + 1u;
        void *p = nullptr; /* my lower neighbour receives from me in its UPPER-side buffers (index 1), and vice versa */
        if ((rc = open(r, 0 + (1 - s), &p))) return rc;
        mb.peer_emig_recv[s] = (unsigned char *)p;
        if ((rc = open(r, 2 + (1 - s), &p))) return rc;
        mb.peer_halo_recv[s] = (unsigned char *)p;
    }
    mb.p2p = 1u;
    m->p2p = true;
    for (int k = 0; k < 2; ++k) /* the captured steps used NCCL */
        if (m->gexec[k]) {
            cudaGraphExecDestroy(m->gexec[k]);
            m->gexec[k] = nullptr;
        }
    return DSIM_OK;
#endif
}

#if !defined(DSIM_EMU)
#define NCK(call)                                                                         \
    do {                                                                                  \
        ncclResult_t r_ = (call);                                                         \
        if (r_ != ncclSuccess) {                                                          \
            char buf_[256];                                                               \
            std::snprintf(buf_, sizeof buf_, "%s failed: %s", #call, ncclGetErrorString(r_)); \
            return fail(h, DSIM_ERR_COMM, buf_);                                          \
        }                                                                                 \
    } while (0)

static int mg_exchange(dsim_handle *h, dsim_mg_state *m, int channel, cudaStream_t stream = nullptr) {
    const MgBuffers &mb = m->mb;
    if (stream == nullptr) stream = h->stream;
    /* peer-memory transport: nothing to do here — the packing kernels (k_mg_split_ids, k_mg_pack, k_mg_halo_pack) store
     * straight into the receivers' buffers and publish a flag, the consuming kernels (k_children, k_mg_unpack,
     * k_mg_halo_unpack) start by waiting for it */
    if (m->p2p) return DSIM_OK;
    if (channel == DSIM_MG_CH_SPLIT) {
        NCK(ncclAllGather(mb.split_send, mb.split_recv, 1 + (size_t)mb.scap, ncclUint64, m->comm, stream));
        return DSIM_OK;
    }
    const uint64_t bytes = channel == DSIM_MG_CH_EMIGRANTS ? m->emig_bytes : m->halo_bytes;
    NCK(ncclGroupStart());
    for (int s = 0; s < 2; ++s) {
        if (!mb.has_peer[s]) continue;
        const int peer = s == 0 ? (int)m->rank - 1 : (int)m->rank + 1;
        unsigned char *snd = channel == DSIM_MG_CH_EMIGRANTS ? mb.emig_send[s] : mb.halo_send[s];
        unsigned char *rcv = channel == DSIM_MG_CH_EMIGRANTS ? mb.emig_recv[s] : mb.halo_recv[s];
        NCK(ncclSend(snd, bytes, ncclChar, peer, m->comm, stream));
        NCK(ncclRecv(rcv, bytes, ncclChar, peer, m->comm, stream));
    }
    NCK(ncclGroupEnd());
    return DSIM_OK;
}
#endif

#if !defined(DSIM_EMU)
static int mg_enqueue_step(dsim_handle *h, dsim_mg_state *m) {
    /* Two of the three exchanges ride on the second stream next to work that does not need them: the SPLIT
     * all-gather (children's ids, first used by k_children) next to k_move, and the HALO exchange (views for the next
     * step's k_move) next to the allocator bookkeeping.  The communicator sees its operations in one order: every
     * exchange is joined back into the main stream before the next one is enqueued. */
    int rc;
    if (m->p2p) {
        /* Peer memory: the messages travel inside the packing kernels, and the step is ROTATED — it begins by unpacking
         * the halo the previous step sent (k_move is its first reader) and ends right after k_mg_halo_pack has sent its
         * own, so no rank waits for a neighbour at the end of a step:
         *   main stream   k_lifecycle | k_move | k_mg_pack | k_mg_unpack ... patch kernels | k_mg_halo_pack (+ end-of-step bookkeeping)
         *   second stream k_mg_halo_unpack, k_mg_split_ids, k_children (next to k_move) | k_mg_sort_out (next to k_mg_pack)
         * 17 launches, 10 of them in sequence. */
        if ((rc = mg_launch_phase(h, m, MG_PHASE_ROTATED_HEAD))) return rc;
        if ((rc = mg_launch_phase(h, m, DSIM_MG_PHASE_PATCH, true))) return rc; /* k_mg_halo_pack's last block ends the step */
        return DSIM_OK;
    }
    if ((rc = mg_launch_phase(h, m, DSIM_MG_PHASE_LIFECYCLE))) return rc;
    CK(cudaEventRecord(h->fork_ev, h->stream));
    CK(cudaStreamWaitEvent(h->stream2, h->fork_ev, 0));
    if ((rc = mg_exchange(h, m, DSIM_MG_CH_SPLIT, h->stream2))) return rc;
    CK(cudaEventRecord(h->join_ev, h->stream2));
    if ((rc = mg_launch_phase(h, m, MG_PHASE_MOVE_A))) return rc;
    CK(cudaStreamWaitEvent(h->stream, h->join_ev, 0));
    if ((rc = mg_launch_phase(h, m, MG_PHASE_MOVE_B))) return rc;
    if ((rc = mg_exchange(h, m, DSIM_MG_CH_EMIGRANTS))) return rc;
    if ((rc = mg_launch_phase(h, m, DSIM_MG_PHASE_PATCH))) return rc;
    CK(cudaEventRecord(h->fork_ev, h->stream));
    CK(cudaStreamWaitEvent(h->stream2, h->fork_ev, 0));
    if ((rc = mg_exchange(h, m, DSIM_MG_CH_HALO, h->stream2))) return rc;
    if ((rc = mg_launch_phase(h, m, MG_PHASE_FINISH_HALO))) return rc;
    CK(cudaEventRecord(h->join_ev, h->stream2));
    if ((rc = mg_launch_phase(h, m, MG_PHASE_FINISH_BOOK))) return rc;
    CK(cudaStreamWaitEvent(h->stream, h->join_ev, 0));
    return DSIM_OK;
}
#endif

/* one full multi-GPU step with the built-in transport; everything is enqueued, nothing waits on the host.
 * The 17 kernels and 3 NCCL operations of a step are captured once per occupied-list parity and replayed. */
int dsim_mg_step_enqueue(dsim_handle *h) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m) return DSIM_ERR_INVALID;
#if defined(DSIM_EMU)
    return fail(h, DSIM_ERR_COMM, "no transport: drive dsim_mg_phase yourself");
#else
    if (!m->nccl && !m->p2p)
        return fail(h, DSIM_ERR_COMM, "no transport: call dsim_mg_init_nccl or dsim_mg_ipc_connect, or drive dsim_mg_phase yourself");
    int rc = DSIM_OK;
    if (h->opt.use_graphs && !m->graph_failed && !h->profiling) {
        if (m->graph_huge != h->huge_enabled) { /* k_patch_huge has joined the step: capture again */
            for (auto &g : m->gexec)
                if (g) {
                    cudaGraphExecDestroy(g);
                    g = nullptr;
                }
            m->graph_huge = h->huge_enabled;
        }
        if (!m->gexec[m->occ_par]) {
            if (std::getenv("DSIM_DEBUG")) std::fprintf(stderr, "[dsim] rank %u: capturing the multi-GPU step (parity %u)\n", m->rank, m->occ_par);
            cudaGraph_t g = nullptr;
            const uint64_t launches0 = h->total_launches;
            bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                rc = mg_enqueue_step(h, m);
                ok = cudaStreamEndCapture(h->stream, &g) == cudaSuccess && rc == DSIM_OK && g != nullptr;
            }
            if (ok) apply_l2_window(h, g);
            if (ok) ok = cudaGraphInstantiate(&m->gexec[m->occ_par], g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
            m->launches_per_step = h->total_launches - launches0;
            h->total_launches = launches0;
            if (!ok) { /* fall back to plain launches for the rest of the run */
                (void)cudaGetLastError();
                m->gexec[m->occ_par] = nullptr;
                m->graph_failed = true;
            }
        }
        if (m->gexec[m->occ_par]) {
            CK(cudaGraphLaunch(m->gexec[m->occ_par], h->stream));
            if (m->p2p) m->halo_pending = true;
            h->total_launches += m->launches_per_step;
            m->occ_par ^= 1u;
            h->mg_occ_prev = m->occ_par ^ 1u;
            return DSIM_OK;
        }
    }
    if (h->profiling) { /* per-segment device times (dsim_profile_enable): plain launches, one sync per step */
        if (m->halo_pending) { /* a rotated step left its halo to the next one: this path unpacks at the END of a step */
            if ((rc = mg_launch_phase(h, m, DSIM_MG_PHASE_HALO_UNPACK))) return rc;
            m->halo_pending = false;
        }
        static const int seq[7] = {DSIM_MG_PHASE_LIFECYCLE, -1 - DSIM_MG_CH_SPLIT, DSIM_MG_PHASE_MOVE, -1 - DSIM_MG_CH_EMIGRANTS,
                                   DSIM_MG_PHASE_PATCH, -1 - DSIM_MG_CH_HALO, DSIM_MG_PHASE_FINISH};
        for (int k = 0; k < 7; ++k) {
            CK(cudaEventRecord(m->seg_ev[k], h->stream));
            rc = seq[k] >= 0 ? mg_launch_phase(h, m, seq[k]) : mg_exchange(h, m, -1 - seq[k]);
            if (rc) return rc;
        }
        CK(cudaEventRecord(m->seg_ev[7], h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int k = 0; k < 7; ++k) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, m->seg_ev[k], m->seg_ev[k + 1]));
            m->seg_ms[k] += ms;
        }
        for (size_t k = 0; k + 1 < m->kev_used; k += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, m->kev[k], m->kev[k + 1]) != cudaSuccess) continue;
            const int id = m->kev_id[k];
            if (id >= 0)
                h->k_ms[id] += ms;
            else
                m->mgk_ms[-1 - id] += ms;
        }
        m->kev_used = 0;
        m->seg_steps += 1;
    } else {
        if ((rc = mg_enqueue_step(h, m))) return rc;
        if (m->p2p) m->halo_pending = true;
    }
    m->occ_par ^= 1u;
    h->mg_occ_prev = m->occ_par ^ 1u;
    return DSIM_OK;
#endif
}

/* diagnostics: mean device milliseconds of the seven segments of a multi-GPU step measured while profiling was on
 * (lifecycle, SPLIT exchange, move, EMIGRANTS exchange, patch, HALO exchange, finish) */
int dsim_mg_segment_times(dsim_handle *h, double *out7) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !out7) return DSIM_ERR_INVALID;
    for (int k = 0; k < 7; ++k) out7[k] = m->seg_steps ? m->seg_ms[k] / (double)m->seg_steps : 0.0;
    return DSIM_OK;
}

/* diagnostics: mean device milliseconds per launch of the multi-GPU kernels measured while profiling was on, in the order
 * k_mg_split_ids, k_mg_sort_out, k_mg_pack, k_mg_unpack, k_mg_halo_pack, k_mg_halo_unpack, k_mg_finish; then the step
 * kernels k_lifecycle, k_move, k_children, k_segments, k_scatter (the patch kernels run concurrently and are not timed
 * one by one) */
int dsim_mg_kernel_times(dsim_handle *h, double *out12) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !out12) return DSIM_ERR_INVALID;
    const double inv = m->seg_steps ? 1.0 / (double)m->seg_steps : 0.0;
    for (int k = 0; k < 7; ++k) out12[k] = m->mgk_ms[k] * inv;
    for (int k = 0; k < 5; ++k) out12[7 + k] = h->k_ms[k] * inv;
    return DSIM_OK;
}

int dsim_mg_halo_sync(dsim_handle *h) { /* initial halo, after the state was set on every rank */
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m) return DSIM_ERR_INVALID;
#if defined(DSIM_EMU)
    return fail(h, DSIM_ERR_COMM, "no transport");
#else
    if (!m->nccl && !m->p2p) return fail(h, DSIM_ERR_COMM, "no transport");
    int rc;
    if (m->p2p) { /* peer memory: an exchange round of its own; every rank calls dsim_mg_halo_sync equally often */
        dsim_mg_launch(DSIM_MGK_BUMP, mg_args(h, false), m->mb, h->geom, h->stream);
        h->total_launches += 1;
    }
    if ((rc = mg_launch_phase(h, m, DSIM_MG_PHASE_HALO_PACK))) return rc;
    rc = mg_exchange(h, m, DSIM_MG_CH_HALO);
    if (rc) return rc;
    if ((rc = mg_launch_phase(h, m, DSIM_MG_PHASE_HALO_UNPACK))) return rc;
    m->halo_pending = false;
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
#endif
}

int dsim_mg_reduce_stats(dsim_handle *h, dsim_step_stats *stats) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !stats) return DSIM_ERR_INVALID;
#if defined(DSIM_EMU)
    return fail(h, DSIM_ERR_COMM, "no transport");
#else
    if (!m->nccl) return fail(h, DSIM_ERR_COMM, "no transport");
    uint64_t v[6] = {stats->live_families, stats->births, stats->deaths, stats->family_updates, stats->occupied_patches,
                     stats->model_errors};
    uint64_t *d = nullptr;
    std::vector<void *> tmp;
    int rc = dev_alloc(h, &d, 6, tmp, false);
    if (!rc) rc = upload(h, d, v, 6);
    if (!rc && ncclAllReduce(d, d, 5, ncclUint64, ncclSum, m->comm, h->stream) != ncclSuccess) rc = fail(h, DSIM_ERR_COMM, "ncclAllReduce failed");
    if (!rc && ncclAllReduce(d + 5, d + 5, 1, ncclUint64, ncclMax, m->comm, h->stream) != ncclSuccess) rc = fail(h, DSIM_ERR_COMM, "ncclAllReduce failed");
    if (!rc) rc = download(h, v, d, 6);
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, DSIM_ERR_CUDA, "synchronize failed");
    free_list(tmp);
    if (rc) return rc;
    stats->live_families = v[0]; stats->births = v[1]; stats->deaths = v[2]; stats->family_updates = v[3];
    stats->occupied_patches = v[4];
    return DSIM_OK;
#endif
}

/* diagnostics: the MG_* counters of the last step (MG_MAX_EMIG = most emigrants per side and step so far) */
int dsim_mg_debug_counters(dsim_handle *h, uint32_t *out8) {
    dsim_mg_state *m = h ? mg_of(h) : nullptr;
    if (!m || !out8) return DSIM_ERR_INVALID;
    int rc = download(h, out8, m->mb.counters, MG_N_COUNTERS);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return DSIM_OK;
}

void dsim_mg_release(dsim_handle *h) {
    dsim_mg_state *m = mg_of(h);
    if (!m) return;
#if !defined(DSIM_EMU)
    for (auto &g : m->gexec)
        if (g) cudaGraphExecDestroy(g);
    if (m->comm) ncclCommDestroy(m->comm);
    for (void *p : m->ipc_opened) cudaIpcCloseMemHandle(p);
    for (auto &e : m->seg_ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : m->kev) cudaEventDestroy(e);
    if (m->life_ev) cudaEventDestroy(m->life_ev);
    if (m->split_ev) cudaEventDestroy(m->split_ev);
#endif
    free_list(m->allocs);
    g_mg.erase(h);
    delete m;
}

} /* extern "C" */
