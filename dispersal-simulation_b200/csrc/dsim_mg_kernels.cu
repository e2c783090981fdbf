/*
 * dsim_mg_kernels.cu — kernels of the band-partitioned multi-GPU step.
 *
 * The reference is a single process (no communication layer at all); this is new code.
 * Every rank holds the static tables with GLOBAL node ids and owns the nodes [own_lo, own_hi) and the
 * families located there.  Per step three fixed-size, device-resident messages travel between ranks:
 *   SPLIT      all ranks <-> all ranks: ids of the families that split this step, so that children get
 *              the same consecutive ids as in a single-GPU run (ids key every random draw);
 *   EMIGRANTS  rank <-> rank +-1: families whose new node belongs to the neighbour (56-byte record,
 *              adaptation slots, history ring);
 *   HALO       rank <-> rank +-1: the 32-byte views (census, quality, band cultures) of the 2S nodes on
 *              each side of a boundary, S = largest node-id distance in the reach table.
 * Nothing here depends on the order in which families are stored, so the partitioned run is
 * bit-identical to the single-GPU run (tests/test_multi_rank.py).
 */
#include "dsim_kernels.h"
#include "dsim_math.h"

/* ---- SPLIT ------------------------------------------------------------------------------------ */
/* ids of the local families that split, at their local child rank; split_send[0] = count */
__global__ void __launch_bounds__(256) k_mg_split_ids(StepArgs a, MgBuffers mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t n = ctl->n_cur[a.parity];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = i >> 5, lt = (1u << (i & 31u)) - 1u;
        const uint32_t split_mask = a.sc.split_bits[c];
        if (!((split_mask >> (i & 31u)) & 1u)) continue;
        uint32_t vs = 0;
        vs += a.sc.word_pre[c] >> 16;
        const uint32_t crank = a.sc.blk_split[c >> 3] + vs + (uint32_t)__popc(split_mask & lt);
        if (crank < mb.scap)
            mb.split_send[1 + crank] = a.cur.key[i].id;
        else
            atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        mb.split_send[0] = ctl->n_children;
        if (mb.p2p) *mb.seq += 1u; /* the exchange round of this step (first kernel of the step that knows the transport) */
    }
}

/* global rank of every local child = number of splitting parents, on any rank, with a smaller id.
 * One warp per child: the lanes stride over the gathered lists. */
__global__ void __launch_bounds__(256) k_mg_child_rank(StepArgs a, MgBuffers mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t n_local = ctl->n_children;
    const size_t stride = 1 + (size_t)mb.scap;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    /* peer-memory transport: the lists of even and odd rounds arrive in two areas (see k_mg_push) */
    const uint64_t *recv = mb.split_recv + (mb.p2p ? (size_t)(*mb.seq & 1u) * a.mg.world * stride : 0u);
    for (uint32_t c = gw; c < n_local; c += nw) {
        const uint64_t id = mb.split_send[1 + c];
        uint32_t smaller = 0;
        for (uint32_t r = 0; r < a.mg.world; ++r) {
            const uint64_t *list = recv + r * stride;
            const uint32_t m = (uint32_t)list[0];
            for (uint32_t k = lane; k < m; k += 32u) smaller += list[1 + k] < id ? 1u : 0u;
        }
        smaller += __shfl_xor_sync(DSIM_FULL, smaller, 16);
        smaller += __shfl_xor_sync(DSIM_FULL, smaller, 8);
        smaller += __shfl_xor_sync(DSIM_FULL, smaller, 4);
        smaller += __shfl_xor_sync(DSIM_FULL, smaller, 2);
        smaller += __shfl_xor_sync(DSIM_FULL, smaller, 1);
        if (lane == 0) mb.child_grank[c] = smaller;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint64_t total = 0;
        for (uint32_t r = 0; r < a.mg.world; ++r) total += recv[r * stride];
        ctl->next_id = ctl->child_id_base + total; /* k_lifecycle's tail had added the local count only */
        mb.counters[MG_N_STAY] = 0;
        mb.counters[MG_N_EMIG0] = 0;
        mb.counters[MG_N_EMIG1] = 0;
        mb.counters[MG_HALO_CULT0] = 0;
        mb.counters[MG_HALO_CULT1] = 0;
    }
}

/* ---- EMIGRANTS -------------------------------------------------------------------------------- */
/* After part 1 the step's families sit in `next` (survivors + children).  Those whose node is owned stay:
 * they are copied to `cur` (which part 1 no longer needs) at positions handed out by an atomic counter
 * — storage order never influences results.  The others are listed per side for k_mg_pack. */
__global__ void __launch_bounds__(256) k_mg_sort_out(StepArgs a, MgBuffers mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t n = ctl->n_next;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_round = (n + 31u) & ~31u; /* whole warps take part in the ballots */
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x) {
        const bool in = j < n;
        const uint32_t loc = in ? a.next.place[j].loc : 0u;
        const bool stay = in && loc >= a.mg.own_lo && loc < a.mg.own_hi;
        /* one atomic per warp for the families that stay */
        const uint32_t m = __ballot_sync(DSIM_FULL, stay);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(&mb.counters[MG_N_STAY], (uint32_t)__popc(m));
        base = __shfl_sync(DSIM_FULL, base, 0);
        if (stay) {
            const uint32_t p = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
            a.cur.key[p] = a.next.key[j];
            a.cur.place[p] = a.next.place[j];
            a.cur.stock[p] = a.next.stock[j];
            a.cur.clock[p] = a.next.clock[j];
            mb.fslot2[p] = a.sc.fslot[j];
        } else if (in) {
            const uint32_t side = loc < a.mg.own_lo ? 0u : 1u;
            const uint32_t e = atomicAdd(&mb.counters[MG_N_EMIG0 + side], 1u);
            if (e < mb.ecap)
                mb.emig_list[side * mb.ecap + e] = j;
            else
                atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
        }
    }
}

/* one warp per emigrant: record + adaptation slots + history (chronological) into the send buffer */
__global__ void __launch_bounds__(128) k_mg_pack(StepArgs a, MgBuffers mb) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t hcap = a.hv.hcap, acap = a.hv.acap;
    for (uint32_t side = 0; side < 2u; ++side) {
        uint32_t m = mb.counters[MG_N_EMIG0 + side];
        if (m > mb.ecap) m = mb.ecap;
        unsigned char *buf = mb.emig_send[side];
        if (gw == 0 && lane == 0) {
            *reinterpret_cast<uint64_t *>(buf) = m;
            atomicMax(&mb.counters[MG_MAX_EMIG], mb.counters[MG_N_EMIG0 + side]);
        }
        for (uint32_t e = gw; e < m; e += nw) {
            const uint32_t j = mb.emig_list[side * mb.ecap + e];
            const FamPlace pl = a.next.place[j];
            const uint32_t hs = pl.hs;
            unsigned char *rec = buf + 16 + (size_t)e * mb.rec_bytes;
            const uint32_t cnt = a.hv.hist_cnt[hs];
            const uint32_t valid = cnt < hcap ? cnt : hcap;
            if (lane == 0) {
                uint64_t *q = reinterpret_cast<uint64_t *>(rec);
                const FamKey ky = a.next.key[j];
                const FamStock sk = a.next.stock[j];
                const FamClock ck = a.next.clock[j];
                q[0] = ky.id;
                q[1] = ky.culture;
                q[2] = dsim_f64_to_bits(sk.stored);
                q[3] = dsim_f64_to_bits(sk.last_harvest);
                q[4] = a.hv.parent[hs];
                uint32_t *w = reinterpret_cast<uint32_t *>(rec + 40);
                w[0] = pl.loc;
                w[1] = pl.size;
                w[2] = ck.till_adult;
                w[3] = ck.till_mut;
                w[4] = pl.misc;
                w[5] = valid;
                w[6] = a.hv.decay[hs];
                w[7] = a.hv.birth_index[hs];
                /* the emigrant's heavy slot is free again: listed behind the slots of the dead */
                mb.emig_hs[atomicAdd(&mb.counters[MG_N_FREED], 1u)] = hs;
            }
            uint16_t *r_eco = reinterpret_cast<uint16_t *>(rec + MG_REC_HDR);
            uint32_t *r_cnt = reinterpret_cast<uint32_t *>(rec + MG_REC_HDR + mg_pad8(acap * 2u));
            double *r_val = reinterpret_cast<double *>(rec + MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u));
            for (uint32_t k = lane; k < acap; k += 32u) {
                r_eco[k] = a.hv.a_eco[(size_t)hs * acap + k];
                r_cnt[k] = a.hv.a_cnt[(size_t)hs * acap + k];
                r_val[k] = a.hv.a_val[(size_t)hs * acap + k];
            }
            unsigned char *hist = rec + MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u) + acap * 8u;
            uint32_t *r_hp = reinterpret_cast<uint32_t *>(hist);
            double *r_hh = reinterpret_cast<double *>(hist + mg_pad8(hcap * 4u));
            for (uint32_t k = lane; k < valid; k += 32u) { /* chronological k = newest-first valid-1-k */
                const uint32_t pos = (cnt - 1u - (valid - 1u - k)) % hcap;
                r_hp[k] = a.hv.hist_patch[(size_t)hs * hcap + pos];
                r_hh[k] = a.hv.hist_harv[(size_t)hs * hcap + pos];
            }
        }
    }
}

/* one warp per immigrant: append to `cur`, take a heavy slot, count the arrival */
__global__ void __launch_bounds__(128) k_mg_unpack(StepArgs a, MgBuffers mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t hcap = a.hv.hcap, acap = a.hv.acap;
    const uint32_t n_stay = mb.counters[MG_N_STAY];
    uint32_t base = 0;
    for (uint32_t side = 0; side < 2u; ++side) {
        const unsigned char *buf = mb.emig_recv[side];
        const uint32_t m = mb.has_peer[side] ? (uint32_t)*reinterpret_cast<const uint64_t *>(buf) : 0u;
        for (uint32_t e = gw; e < m; e += nw) {
            const unsigned char *rec = buf + 16 + (size_t)e * mb.rec_bytes;
            const uint32_t p = n_stay + base + e;
            /* heavy slots: the children popped the first n_children, the immigrants continue */
            const uint32_t pop = ctl->n_children + base + e, ft = ctl->free_top;
            const uint32_t hs = (pop < ft) ? a.sc.free_hs[ft - 1u - pop] : ctl->hs_hwm + (pop - ft);
            const uint32_t *w = reinterpret_cast<const uint32_t *>(rec + 40);
            const uint32_t loc = w[0], valid = w[5];
            if (p >= ctl->cap || hs >= ctl->hs_cap || loc < a.mg.own_lo || loc >= a.mg.own_hi) {
                if (lane == 0) atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
                continue;
            }
            const uint16_t *r_eco = reinterpret_cast<const uint16_t *>(rec + MG_REC_HDR);
            const uint32_t *r_cnt = reinterpret_cast<const uint32_t *>(rec + MG_REC_HDR + mg_pad8(acap * 2u));
            const double *r_val = reinterpret_cast<const double *>(rec + MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u));
            for (uint32_t k = lane; k < acap; k += 32u) {
                a.hv.a_eco[(size_t)hs * acap + k] = r_eco[k];
                a.hv.a_cnt[(size_t)hs * acap + k] = r_cnt[k];
                a.hv.a_val[(size_t)hs * acap + k] = r_val[k];
            }
            const unsigned char *hist = rec + MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u) + acap * 8u;
            const uint32_t *r_hp = reinterpret_cast<const uint32_t *>(hist);
            const double *r_hh = reinterpret_cast<const double *>(hist + mg_pad8(hcap * 4u));
            for (uint32_t k = lane; k < valid; k += 32u) {
                a.hv.hist_patch[(size_t)hs * hcap + k] = r_hp[k];
                a.hv.hist_harv[(size_t)hs * hcap + k] = r_hh[k];
            }
            if (lane == 0) {
                const uint64_t *q = reinterpret_cast<const uint64_t *>(rec);
                FamKey ky;
                ky.id = q[0];
                ky.culture = q[1];
                FamStock sk;
                sk.stored = dsim_bits_to_f64(q[2]);
                sk.last_harvest = dsim_bits_to_f64(q[3]);
                FamPlace pl;
                pl.loc = loc;
                pl.size = w[1];
                pl.hs = hs;
                pl.misc = w[4];
                FamClock ck;
                ck.till_adult = w[2];
                ck.till_mut = w[3];
                a.cur.key[p] = ky;
                a.cur.stock[p] = sk;
                a.cur.place[p] = pl;
                a.cur.clock[p] = ck;
                a.hv.parent[hs] = q[4];
                a.hv.hist_cnt[hs] = valid;
                a.hv.decay[hs] = w[6];
                a.hv.birth_index[hs] = (uint16_t)w[7];
                mb.fslot2[p] = atomicAdd(&a.ps.cnt[loc], 1u); /* rank 0 = first arrival: k_segments lists the node */
            }
        }
        base += m;
    }
    if (gw == 0 && lane == 0) { /* the families of part 2 */
        ctl->n_next = n_stay + base;
        mb.counters[MG_N_IMMIG] = base;
        if (n_stay + base > ctl->cap) ctl->errors |= DSIM_MODEL_CAPACITY;
    }
}

/* ---- HALO ------------------------------------------------------------------------------------- */
/* the views of my boundary strips, with the band cultures they point to packed behind them:
 * message = [header 16][views hw x 16][bands hw x 16][cultures ccap x 8] */
__global__ void __launch_bounds__(256) k_mg_halo_pack(StepArgs a, MgBuffers mb) {
    for (uint32_t side = 0; side < 2u; ++side) {
        if (!mb.has_peer[side]) continue;
        const uint32_t first = side == 0u ? a.mg.own_lo : a.mg.own_hi - mb.hw;
        unsigned char *buf = mb.halo_send[side];
        PatchView *views = reinterpret_cast<PatchView *>(buf + 16);
        PatchBands *bands = reinterpret_cast<PatchBands *>(buf + 16 + (size_t)mb.hw * sizeof(PatchView));
        uint64_t *cults = reinterpret_cast<uint64_t *>(buf + 16 + (size_t)mb.hw * (sizeof(PatchView) + sizeof(PatchBands)));
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < mb.hw; k += gridDim.x * blockDim.x) {
            PatchView v = a.ps.view[first + k];
            PatchBands b = a.ps.bands[first + k];
            if (v.nb >= 2u) { /* the first culture travels inline */
                const uint32_t at = atomicAdd(&mb.counters[MG_HALO_CULT0 + side], v.nb);
                if (at + v.nb <= mb.ccap) {
                    for (uint32_t c = 0; c < v.nb; ++c) cults[at + c] = a.ps.band_cult[b.stor_off + c];
                    b.stor_off = DSIM_STOR_HALO | (side << 30) | at; /* relative; the receiver knows the side */
                } else {
                    atomicOr(&a.ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
                    v.nb = 1u;
                }
            }
            views[k] = v;
            bands[k] = b;
        }
    }
}

__global__ void __launch_bounds__(256) k_mg_halo_unpack(StepArgs a, MgBuffers mb) {
    for (uint32_t side = 0; side < 2u; ++side) {
        if (!mb.has_peer[side]) continue;
        /* what arrives from the lower neighbour is its UPPER strip: the nodes just below mine */
        const uint32_t first = side == 0u ? a.mg.own_lo - mb.hw : a.mg.own_hi;
        const unsigned char *buf = mb.halo_recv[side];
        const PatchView *views = reinterpret_cast<const PatchView *>(buf + 16);
        const PatchBands *bands = reinterpret_cast<const PatchBands *>(buf + 16 + (size_t)mb.hw * sizeof(PatchView));
        const uint64_t *cults = reinterpret_cast<const uint64_t *>(buf + 16 + (size_t)mb.hw * (sizeof(PatchView) + sizeof(PatchBands)));
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < mb.hw; k += gridDim.x * blockDim.x) {
            PatchBands b = bands[k];
            if (b.stor_off & DSIM_STOR_HALO) b.stor_off = DSIM_STOR_HALO | (side << 30) | (b.stor_off & 0x3FFFFFFFu);
            a.ps.view[first + k] = views[k];
            a.ps.bands[first + k] = b;
        }
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < mb.ccap; k += gridDim.x * blockDim.x)
            mb.halo_cult[side][k] = cults[k];
    }
}

/* ---- peer-memory transport ------------------------------------------------------------------------------------
 * k_mg_push copies the USED part of a send buffer into the neighbour's receive buffer with 16-byte remote stores
 * (NVLink peer memory, mapped through CUDA IPC), then the last block to finish publishes the round number in the
 * neighbour's flag word; k_mg_wait spins on this rank's flag words until the round has arrived.
 * EMIGRANTS and HALO (neighbour channels) need no double buffering: a rank cannot start the push of round r+1 on one
 * of them before that neighbour has consumed round r (it first has to receive the neighbour's later messages of
 * round r).  SPLIT is all-to-all, and nothing orders a rank's push of round r+1 after the consumption of round r
 * (k_mg_child_rank, which runs after that rank's own k_move) on a rank three or more bands away: its receive area and
 * flag words are therefore doubled by round parity.  Round r+2 cannot be pushed before every rank has pushed round
 * r+1 (the pusher's HALO r+1 needs EMIGRANTS r+1 of its neighbour, which needs SPLIT r+1 of every rank), i.e. before
 * every rank has consumed round r. */
#if defined(DSIM_EMU)
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { *p = v; }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) { return *p; }
__device__ __forceinline__ void fence_sys() {}
__device__ __forceinline__ void backoff() {}
#else
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_sys() { __threadfence_system(); }
__device__ __forceinline__ void backoff() { __nanosleep(100); }
#endif

__device__ __forceinline__ void copy16(unsigned char *dst, const unsigned char *src, uint64_t bytes) { /* bytes: a multiple of 16 */
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < bytes / 16u; k += (uint64_t)gridDim.x * blockDim.x) d[k] = s[k];
}

__global__ void __launch_bounds__(256) k_mg_push(StepArgs a, MgBuffers mb, int channel) {
    const uint32_t round = *mb.seq;
    if (channel == DSIM_MG_CH_SPLIT) {
        const uint64_t words = 1u + mb.split_send[0]; /* count, ids: 8-byte words (a rank's slot is only 8-byte aligned) */
        for (uint32_t r = 0; r < a.mg.world; ++r) {
            uint64_t *dst = mb.peer_split_recv[r] + ((size_t)(round & 1u) * a.mg.world + a.mg.rank) * (1 + (size_t)mb.scap);
            for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < words; k += (uint64_t)gridDim.x * blockDim.x) dst[k] = mb.split_send[k];
        }
    } else {
        for (uint32_t side = 0; side < 2u; ++side) {
            if (!mb.has_peer[side]) continue;
            if (channel == DSIM_MG_CH_EMIGRANTS) {
                const uint64_t m = *reinterpret_cast<const uint64_t *>(mb.emig_send[side]);
                copy16(mb.peer_emig_recv[side], mb.emig_send[side], 16u + m * mb.rec_bytes); /* rec_bytes is a multiple of 16 */
            } else {
                uint32_t nc = mb.counters[MG_HALO_CULT0 + side];
                if (nc > mb.ccap) nc = mb.ccap;
                const uint64_t bytes = (16u + (uint64_t)mb.hw * (sizeof(PatchView) + sizeof(PatchBands)) + (uint64_t)nc * 8u + 15u) & ~(uint64_t)15u;
                copy16(mb.peer_halo_recv[side], mb.halo_send[side], bytes);
            }
        }
    }
    /* the last block to get here publishes the round in the receivers' flag words */
    fence_sys();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(&mb.done[channel], 1u);
        if (prev == gridDim.x - 1u) {
            mb.done[channel] = 0u;
            fence_sys();
            if (channel == DSIM_MG_CH_SPLIT) {
                for (uint32_t r = 0; r < a.mg.world; ++r)
                    st_release_sys(mb.peer_flags[r] + MG_FLAG_SPLIT0 + (round & 1u) * DSIM_MG_MAX_WORLD + a.mg.rank, round);
            } else {
                const uint32_t slot0 = channel == DSIM_MG_CH_EMIGRANTS ? MG_FLAG_EMIG0 : MG_FLAG_HALO0;
                /* what I send to my lower neighbour (side 0) arrives from ITS upper side (slot 1), and vice versa */
                if (mb.has_peer[0]) st_release_sys(mb.peer_flags[a.mg.rank - 1u] + slot0 + 1u, round);
                if (mb.has_peer[1]) st_release_sys(mb.peer_flags[a.mg.rank + 1u] + slot0 + 0u, round);
            }
        }
    }
}

__global__ void __launch_bounds__(32) k_mg_wait(StepArgs a, MgBuffers mb, int channel) {
    const uint32_t round = *mb.seq;
    uint32_t slot = 0xFFFFFFFFu;
    if (channel == DSIM_MG_CH_SPLIT) {
        if (threadIdx.x < a.mg.world) slot = MG_FLAG_SPLIT0 + (round & 1u) * DSIM_MG_MAX_WORLD + threadIdx.x;
    } else if (threadIdx.x < 2u && mb.has_peer[threadIdx.x]) {
        slot = (channel == DSIM_MG_CH_EMIGRANTS ? MG_FLAG_EMIG0 : MG_FLAG_HALO0) + threadIdx.x;
    }
    if (slot != 0xFFFFFFFFu) {
        uint32_t spins = 0;
        /* rounds only grow; a neighbour that is a step ahead may already have published the next one on a neighbour
         * channel (a SPLIT slot of this parity next sees round + 2, which nobody can publish before this round is consumed) */
        while ((int32_t)(ld_acquire_sys(mb.flags + slot) - round) < 0) {
            backoff();
            if (++spins > (1u << 25)) { /* several seconds: the neighbour is gone; flag it instead of hanging the device */
                atomicOr(&a.ctl->errors, (uint32_t)DSIM_MODEL_COMM_TIMEOUT);
                break;
            }
        }
    }
    fence_sys();
}

/* ---- end of step ------------------------------------------------------------------------------ */
__global__ void k_mg_finish(StepArgs a, MgBuffers mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    /* heavy-slot allocator: children and immigrants popped, the dead and the emigrants are pushed */
    const uint32_t ft = ctl->free_top, pops = ctl->n_children + mb.counters[MG_N_IMMIG];
    const uint32_t dead = ctl->n_dead, freed = mb.counters[MG_N_FREED];
    const uint32_t popped = pops < ft ? pops : ft;
    const uint32_t base = ft - popped;
    for (uint32_t k = tid; k < dead; k += nth) a.sc.free_hs[base + k] = a.sc.dead_hs[k];
    for (uint32_t k = tid; k < freed; k += nth) a.sc.free_hs[base + dead + k] = mb.emig_hs[k];
}

__global__ void k_mg_finish2(StepArgs a, MgBuffers mb) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    DevCtl *ctl = a.ctl;
    const uint32_t ft = ctl->free_top, pops = ctl->n_children + mb.counters[MG_N_IMMIG];
    const uint32_t popped = pops < ft ? pops : ft;
    ctl->hs_hwm += pops - popped;
    ctl->free_top = ft - popped + ctl->n_dead + mb.counters[MG_N_FREED];
    if (ctl->hs_hwm > ctl->hs_cap) ctl->errors |= DSIM_MODEL_CAPACITY;
    mb.counters[MG_N_FREED] = 0;
    ctl->n_cur[a.parity] = ctl->n_next; /* part 2 ran in place on `cur`: the buffers do not swap */
    ctl->t += 1u;
}

/* an exchange round outside the steps (the initial halo over peer memory): every rank calls it equally often */
__global__ void k_mg_bump(StepArgs a, MgBuffers mb) {
    (void)a;
    if (blockIdx.x == 0 && threadIdx.x == 0) *mb.seq += 1u;
}

/* ---- launchers -------------------------------------------------------------------------------- */
void dsim_mg_launch(int which, const StepArgs &a, const MgBuffers &mb, const LaunchGeom &g, cudaStream_t s) {
    const uint32_t sm = g.n_sm;
    switch (which) {
    case DSIM_MGK_SPLIT_IDS: DSIM_LAUNCH(k_mg_split_ids, sm * 4, 256, 0, s, a, mb); break;
    case DSIM_MGK_CHILD_RANK: DSIM_LAUNCH(k_mg_child_rank, sm * 2, 256, 0, s, a, mb); break;
    case DSIM_MGK_SORT_OUT: DSIM_LAUNCH(k_mg_sort_out, sm * 8, 256, 0, s, a, mb); break;
    case DSIM_MGK_PACK: DSIM_LAUNCH(k_mg_pack, sm * 4, 128, 0, s, a, mb); break;
    case DSIM_MGK_UNPACK: DSIM_LAUNCH(k_mg_unpack, sm * 4, 128, 0, s, a, mb); break;
    case DSIM_MGK_HALO_PACK: DSIM_LAUNCH(k_mg_halo_pack, sm * 2, 256, 0, s, a, mb); break;
    case DSIM_MGK_HALO_UNPACK: DSIM_LAUNCH(k_mg_halo_unpack, sm * 2, 256, 0, s, a, mb); break;
    case DSIM_MGK_FINISH: DSIM_LAUNCH(k_mg_finish, sm, 256, 0, s, a, mb); break;
    case DSIM_MGK_FINISH2: DSIM_LAUNCH(k_mg_finish2, 1, 32, 0, s, a, mb); break;
    case DSIM_MGK_PUSH_SPLIT: DSIM_LAUNCH(k_mg_push, 8, 256, 0, s, a, mb, (int)DSIM_MG_CH_SPLIT); break;
    case DSIM_MGK_PUSH_EMIG: DSIM_LAUNCH(k_mg_push, sm, 256, 0, s, a, mb, (int)DSIM_MG_CH_EMIGRANTS); break;
    case DSIM_MGK_PUSH_HALO: DSIM_LAUNCH(k_mg_push, 32, 256, 0, s, a, mb, (int)DSIM_MG_CH_HALO); break;
    case DSIM_MGK_WAIT_SPLIT: DSIM_LAUNCH(k_mg_wait, 1, 32, 0, s, a, mb, (int)DSIM_MG_CH_SPLIT); break;
    case DSIM_MGK_WAIT_EMIG: DSIM_LAUNCH(k_mg_wait, 1, 32, 0, s, a, mb, (int)DSIM_MG_CH_EMIGRANTS); break;
    case DSIM_MGK_WAIT_HALO: DSIM_LAUNCH(k_mg_wait, 1, 32, 0, s, a, mb, (int)DSIM_MG_CH_HALO); break;
    case DSIM_MGK_BUMP: DSIM_LAUNCH(k_mg_bump, 1, 32, 0, s, a, mb); break;
    default: break;
    }
}
