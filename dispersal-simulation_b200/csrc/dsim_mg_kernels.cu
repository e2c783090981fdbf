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
 * bit-identical to the single-GPU run (tests/test_multi_rank.py, tests/test_multi_rank_ipc.py, bench.py at N > 1).
 *
 * A message is written by the kernel that produces it and read by the kernel that needs it.  With the peer-memory
 * transport (mb.p2p) the producer stores straight into the receiver's buffer over NVLink and its last block publishes a
 * flag (mg_last_block); the consumer starts by waiting for the flag (mg_wait_flags).  With NCCL or a caller-driven
 * transport the same kernels write and read the local send / receive buffers and the copy happens between them.
 */
#include "dsim_kernels.h"
#include "dsim_math.h"
#include "dsim_mg_sync.h"

/* The last block of a packing kernel to finish publishes the message.  Every thread that stored part of it fences at
 * device scope before its block takes a ticket (the pattern of a last-block reduction); the thread that takes the last
 * ticket has thereby observed all of them, and its ONE system-scope fence, followed by relaxed system-scope stores of the
 * flag words, releases the whole message to the receivers (fences are cumulative: causality order is transitive over
 * the ticket's device-scope synchronisation and the flag's system-scope one).  A system-scope fence with peer stores in
 * flight costs about 5 us on B200 — with one in every writing thread as well, a message cost 10 us instead of 5. */
template <typename F> __device__ __forceinline__ void mg_last_block(uint32_t *ticket, bool wrote, F publish) {
    if (wrote) __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(ticket, 1u);
        if (prev == gridDim.x - 1u) {
            *ticket = 0u;
            mg_fence_sys();
            publish();
        }
    }
}

/* ---- SPLIT ------------------------------------------------------------------------------------ */
/* ids of the local families that split, at their local child rank; split_send[0] = count.
 * Peer-memory transport: the list also goes straight into every rank's receive area of this round's parity (remote
 * 8-byte stores) and the last block to finish starts the exchange round of the step and publishes it in the
 * receivers' flag words — the SPLIT message costs no kernel of its own.
 * SPLIT is all-to-all, and nothing orders a rank's message of round r+1 after the consumption of round r (k_children,
 * which runs after that rank's own k_move) on a rank three or more bands away: receive areas and flag words are
 * doubled by round parity.  Round r+2 cannot be sent before every rank has sent round r+1 (the sender's HALO r+1 needs
 * EMIGRANTS r+1 of its neighbour, which needs SPLIT r+1 of every rank), i.e. before every rank has consumed round r. */
__global__ void __launch_bounds__(256) k_mg_split_ids(StepArgs a, MgBuffers mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t n = ctl->n_cur[a.parity];
    const uint32_t round = mb.p2p ? *mb.seq + 1u : 0u; /* read before any block can have taken the last ticket */
    const size_t stride = 1 + (size_t)mb.scap;
    const size_t slot = ((size_t)(round & 1u) * a.mg.world + a.mg.rank) * stride;
    bool wrote = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = i >> 5, lt = (1u << (i & 31u)) - 1u;
        const uint32_t split_mask = a.sc.split_bits[c];
        if (!((split_mask >> (i & 31u)) & 1u)) continue;
        uint32_t vs = 0;
        vs += a.sc.word_pre[c] >> 16;
        const uint32_t crank = a.sc.blk_split[c >> 3] + vs + (uint32_t)__popc(split_mask & lt);
        if (crank < mb.scap) {
            const uint64_t id = a.cur.key[i].id;
            mb.split_send[1 + crank] = id;
            if (mb.p2p) {
                for (uint32_t r = 0; r < a.mg.world; ++r) mb.peer_split_recv[r][slot + 1 + crank] = id;
                wrote = true;
            }
        } else {
            atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        mb.split_send[0] = ctl->n_children;
        if (mb.p2p) {
            for (uint32_t r = 0; r < a.mg.world; ++r) mb.peer_split_recv[r][slot] = ctl->n_children;
            wrote = true;
        }
    }
    if (!mb.p2p) return;
    mg_last_block(&mb.done[DSIM_MG_CH_SPLIT], wrote, [&] {
        *mb.seq = round; /* the exchange round of this step */
        for (uint32_t r = 0; r < a.mg.world; ++r)
            mg_st_flag(mb.peer_flags[r] + MG_FLAG_SPLIT0 + (round & 1u) * DSIM_MG_MAX_WORLD + a.mg.rank, round);
    });
}

/* ---- EMIGRANTS -------------------------------------------------------------------------------- */
/* After part 1 the step's families sit in `next` (survivors + children).  Those whose node is owned stay:
 * they are copied to `cur` (which part 1 no longer needs) at positions handed out by an atomic counter
 * — storage order never influences results.  The others were listed per side by k_move / k_children when they wrote
 * them (mg_list_emigrant), so k_mg_pack does not depend on this kernel and runs next to it. */
__global__ void __launch_bounds__(256) k_mg_sort_out(StepArgs a, MgBuffers mb) {
    const uint32_t n = a.ctl->n_next;
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
        }
    }
}

/* one warp per emigrant: record + adaptation slots + history (chronological) into the send buffer — with the
 * peer-memory transport straight into the neighbour's receive buffer (remote stores over NVLink: the transfer overlaps
 * the packing record by record), published by the last block to finish */
__global__ void __launch_bounds__(128) k_mg_pack(StepArgs a, MgBuffers mb) {
    /* a BLOCK per emigrant: its 7.8 KB (628-entry history ring) are read with every load of a pass in flight before the
     * first store — a warp per emigrant walked the ring in 20 dependent load -> remote store steps (41 us per launch) */
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t hcap = a.hv.hcap, acap = a.hv.acap;
    uint32_t m0 = mb.counters[MG_N_EMIG0], m1 = mb.counters[MG_N_EMIG1];
    if (blockIdx.x == 0 && tid == 0) atomicMax(&mb.counters[MG_MAX_EMIG], m0 > m1 ? m0 : m1);
    if (m0 > mb.ecap) m0 = mb.ecap;
    if (m1 > mb.ecap) m1 = mb.ecap;
    unsigned char *const bufs[2] = {(mb.p2p && mb.has_peer[0]) ? mb.peer_emig_recv[0] : mb.emig_send[0],
                                    (mb.p2p && mb.has_peer[1]) ? mb.peer_emig_recv[1] : mb.emig_send[1]};
    bool wrote = false;
    if (blockIdx.x == 0 && tid < 2u) {
        *reinterpret_cast<uint64_t *>(bufs[tid]) = tid == 0u ? m0 : m1;
        wrote = true;
    }
    for (uint32_t eg = blockIdx.x; eg < m0 + m1; eg += gridDim.x) {
        const uint32_t side = eg < m0 ? 0u : 1u, e = eg - (side ? m0 : 0u);
        unsigned char *buf = bufs[side];
        wrote = true;
        const uint32_t j = mb.emig_list[side * mb.ecap + e];
        const FamPlace pl = a.next.place[j];
        const uint32_t hs = pl.hs;
        unsigned char *rec = buf + 16 + (size_t)e * mb.rec_bytes;
        const uint32_t cnt = a.hv.hist_cnt[hs];
        const uint32_t valid = cnt < hcap ? cnt : hcap;
        if (tid == 0) {
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
        for (uint32_t k = tid; k < acap; k += nt) {
            r_eco[k] = a.hv.a_eco[(size_t)hs * acap + k];
            r_cnt[k] = a.hv.a_cnt[(size_t)hs * acap + k];
            r_val[k] = a.hv.a_val[(size_t)hs * acap + k];
        }
        unsigned char *hist = rec + MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u) + acap * 8u;
        uint32_t *r_hp = reinterpret_cast<uint32_t *>(hist);
        double *r_hh = reinterpret_cast<double *>(hist + mg_pad8(hcap * 4u));
        const size_t hb = (size_t)hs * hcap;
        const uint32_t oldest = cnt - valid; /* chronological entry k sits at ring position (oldest + k) % hcap */
        for (uint32_t k0 = tid; k0 < valid; k0 += 4u * nt) {
            uint32_t hp[4];
            double hh[4];
#pragma unroll
            for (uint32_t u = 0; u < 4u; ++u) {
                const uint32_t k = k0 + u * nt;
                hp[u] = 0u;
                hh[u] = 0.0;
                if (k < valid) {
                    const uint32_t pos = (oldest + k) % hcap;
                    hp[u] = a.hv.hist_patch[hb + pos];
                    hh[u] = a.hv.hist_harv[hb + pos];
                }
            }
#pragma unroll
            for (uint32_t u = 0; u < 4u; ++u) {
                const uint32_t k = k0 + u * nt;
                if (k < valid) {
                    r_hp[k] = hp[u];
                    r_hh[k] = hh[u];
                }
            }
        }
    }
    if (!mb.p2p) return;
    mg_last_block(&mb.done[DSIM_MG_CH_EMIGRANTS], wrote, [&] {
        const uint32_t round = *mb.seq;
        /* what I send to my lower neighbour (side 0) arrives from ITS upper side (slot 1), and vice versa */
        if (mb.has_peer[0]) mg_st_flag(mb.peer_flags[a.mg.rank - 1u] + MG_FLAG_EMIG0 + 1u, round);
        if (mb.has_peer[1]) mg_st_flag(mb.peer_flags[a.mg.rank + 1u] + MG_FLAG_EMIG0 + 0u, round);
    });
}

/* one warp per immigrant: append to `cur`, take a heavy slot, count the arrival */
__global__ void __launch_bounds__(128) k_mg_unpack(StepArgs a, MgBuffers mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t gw = blockIdx.x, lane = tid; /* a block per immigrant; (gw, lane) == (0, 0) does the bookkeeping below */
    const uint32_t hcap = a.hv.hcap, acap = a.hv.acap;
    if (mb.p2p) mg_wait_flags(mb.flags, MG_FLAG_EMIG0, 2u, mb.has_peer[0] | (mb.has_peer[1] << 1), *mb.seq, &ctl->errors);
    const uint32_t n_stay = mb.counters[MG_N_STAY];
    const uint32_t m0 = mb.has_peer[0] ? (uint32_t)*reinterpret_cast<const uint64_t *>(mb.emig_recv[0]) : 0u;
    const uint32_t m1 = mb.has_peer[1] ? (uint32_t)*reinterpret_cast<const uint64_t *>(mb.emig_recv[1]) : 0u;
    const uint32_t base = m0 + m1;
    for (uint32_t eg = blockIdx.x; eg < base; eg += gridDim.x) {
        const uint32_t side = eg < m0 ? 0u : 1u, e = eg - (side ? m0 : 0u);
        const unsigned char *rec = mb.emig_recv[side] + 16 + (size_t)e * mb.rec_bytes;
        const uint32_t p = n_stay + eg;
        /* heavy slots: the children popped the first n_children, the immigrants continue */
        const uint32_t pop = ctl->n_children + eg, ft = ctl->free_top;
        const uint32_t hs = (pop < ft) ? a.sc.free_hs[ft - 1u - pop] : ctl->hs_hwm + (pop - ft);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(rec + 40);
        const uint32_t loc = w[0], valid = w[5];
        if (p >= ctl->cap || hs >= ctl->hs_cap || loc < a.mg.own_lo || loc >= a.mg.own_hi) {
            if (tid == 0) atomicOr(&ctl->errors, (uint32_t)DSIM_MODEL_CAPACITY);
            continue;
        }
        const uint16_t *r_eco = reinterpret_cast<const uint16_t *>(rec + MG_REC_HDR);
        const uint32_t *r_cnt = reinterpret_cast<const uint32_t *>(rec + MG_REC_HDR + mg_pad8(acap * 2u));
        const double *r_val = reinterpret_cast<const double *>(rec + MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u));
        for (uint32_t k = tid; k < acap; k += nt) {
            a.hv.a_eco[(size_t)hs * acap + k] = r_eco[k];
            a.hv.a_cnt[(size_t)hs * acap + k] = r_cnt[k];
            a.hv.a_val[(size_t)hs * acap + k] = r_val[k];
        }
        const unsigned char *hist = rec + MG_REC_HDR + mg_pad8(acap * 2u) + mg_pad8(acap * 4u) + acap * 8u;
        const uint32_t *r_hp = reinterpret_cast<const uint32_t *>(hist);
        const double *r_hh = reinterpret_cast<const double *>(hist + mg_pad8(hcap * 4u));
        const size_t hb = (size_t)hs * hcap;
        for (uint32_t k0 = tid; k0 < valid; k0 += 4u * nt) { /* every load of a pass in flight before the first store */
            uint32_t hp[4];
            double hh[4];
#pragma unroll
            for (uint32_t u = 0; u < 4u; ++u) {
                const uint32_t k = k0 + u * nt;
                hp[u] = 0u;
                hh[u] = 0.0;
                if (k < valid) {
                    hp[u] = r_hp[k];
                    hh[u] = r_hh[k];
                }
            }
#pragma unroll
            for (uint32_t u = 0; u < 4u; ++u) {
                const uint32_t k = k0 + u * nt;
                if (k < valid) {
                    a.hv.hist_patch[hb + k] = hp[u];
                    a.hv.hist_harv[hb + k] = hh[u];
                }
            }
        }
        if (tid == 0) {
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
    if (gw == 0 && lane == 0) { /* the families of part 2 */
        ctl->n_next = n_stay + base;
        mb.counters[MG_N_IMMIG] = base;
        if (n_stay + base > ctl->cap) ctl->errors |= DSIM_MODEL_CAPACITY;
    }
}

/* ---- end of step ------------------------------------------------------------------------------ */
/* one block: heavy-slot allocator (children and immigrants popped, the dead and the emigrants pushed), t += 1 */
__device__ __forceinline__ void mg_finish_body(const StepArgs &a, const MgBuffers &mb) {
    DevCtl *ctl = a.ctl;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t ft = ctl->free_top, pops = ctl->n_children + mb.counters[MG_N_IMMIG];
    const uint32_t dead = ctl->n_dead, freed = mb.counters[MG_N_FREED];
    const uint32_t popped = pops < ft ? pops : ft;
    const uint32_t base = ft - popped;
    for (uint32_t k = tid; k < dead; k += nth) a.sc.free_hs[base + k] = a.sc.dead_hs[k];
    for (uint32_t k = tid; k < freed; k += nth) a.sc.free_hs[base + dead + k] = mb.emig_hs[k];
    __syncthreads();
    if (tid != 0) return;
    ctl->hs_hwm += pops - popped;
    ctl->free_top = ft - popped + dead + freed;
    if (ctl->hs_hwm > ctl->hs_cap) ctl->errors |= DSIM_MODEL_CAPACITY;
    mb.counters[MG_N_FREED] = 0;
    /* the next step's counters: emigrants are listed by k_move / k_children (next to which k_mg_split_ids runs),
     * stayers are counted by k_mg_sort_out, halo cultures by k_mg_halo_pack */
    mb.counters[MG_N_STAY] = 0;
    mb.counters[MG_N_EMIG0] = 0;
    mb.counters[MG_N_EMIG1] = 0;
    mb.counters[MG_HALO_CULT0] = 0;
    mb.counters[MG_HALO_CULT1] = 0;
    ctl->n_cur[a.parity] = ctl->n_next; /* part 2 ran in place on `cur`: the buffers do not swap */
    ctl->t += 1u;
}
__global__ void __launch_bounds__(256) k_mg_finish(StepArgs a, MgBuffers mb) { mg_finish_body(a, mb); }

/* ---- HALO ------------------------------------------------------------------------------------- */
/* the views of my boundary strips, with the band cultures they point to packed behind them:
 * message = [header 16][views hw x 16][bands hw x 16][cultures ccap x 8] */
__global__ void __launch_bounds__(256) k_mg_halo_pack(StepArgs a, MgBuffers mb) {
    for (uint32_t side = 0; side < 2u; ++side) {
        if (!mb.has_peer[side]) continue;
        const uint32_t first = side == 0u ? a.mg.own_lo : a.mg.own_hi - mb.hw;
        unsigned char *buf = mb.p2p ? mb.peer_halo_recv[side] : mb.halo_send[side];
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
    if (!mb.p2p) return;
    /* the last block to finish publishes the message (mg_last_block) and, in the library's own step (mb.fuse_finish),
     * goes on to do the end-of-step bookkeeping: nothing else of the step is running any more */
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(&mb.done[DSIM_MG_CH_HALO], 1u);
        s_last = prev == gridDim.x - 1u ? 1u : 0u;
        if (s_last) {
            mb.done[DSIM_MG_CH_HALO] = 0u;
            mg_fence_sys();
            const uint32_t round = *mb.seq;
            if (mb.has_peer[0]) mg_st_flag(mb.peer_flags[a.mg.rank - 1u] + MG_FLAG_HALO0 + 1u, round);
            if (mb.has_peer[1]) mg_st_flag(mb.peer_flags[a.mg.rank + 1u] + MG_FLAG_HALO0 + 0u, round);
        }
    }
    __syncthreads();
    if (s_last && mb.fuse_finish) mg_finish_body(a, mb);
}

__global__ void __launch_bounds__(256) k_mg_halo_unpack(StepArgs a, MgBuffers mb) {
    if (mb.p2p) mg_wait_flags(mb.flags, MG_FLAG_HALO0, 2u, mb.has_peer[0] | (mb.has_peer[1] << 1), *mb.seq, &a.ctl->errors);
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

/* an exchange round outside the steps (the initial halo over peer memory): every rank calls it equally often */
__global__ void k_mg_bump(StepArgs a, MgBuffers mb) {
    (void)a;
    if (blockIdx.x == 0 && threadIdx.x == 0) *mb.seq += 1u;
}

/* ---- launchers -------------------------------------------------------------------------------- */
void dsim_mg_launch(int which, const StepArgs &a, const MgBuffers &mb, const LaunchGeom &g, cudaStream_t s) {
    const uint32_t sm = g.n_sm;
    switch (which) {
    case DSIM_MGK_SPLIT_IDS: DSIM_LAUNCH(k_mg_split_ids, sm, 256, 0, s, a, mb); break;
    case DSIM_MGK_SORT_OUT: DSIM_LAUNCH(k_mg_sort_out, sm * 8, 256, 0, s, a, mb); break;
    case DSIM_MGK_PACK: DSIM_LAUNCH(k_mg_pack, sm * 2, 128, 0, s, a, mb); break; /* a block per emigrant; every block ends with a ticket */
    case DSIM_MGK_UNPACK: DSIM_LAUNCH(k_mg_unpack, sm * 2, 128, 0, s, a, mb); break;
    case DSIM_MGK_HALO_PACK: DSIM_LAUNCH(k_mg_halo_pack, sm, 128, 0, s, a, mb); break;
    case DSIM_MGK_HALO_UNPACK: DSIM_LAUNCH(k_mg_halo_unpack, sm, 128, 0, s, a, mb); break;
    case DSIM_MGK_FINISH: DSIM_LAUNCH(k_mg_finish, 1, 256, 0, s, a, mb); break;
    case DSIM_MGK_BUMP: DSIM_LAUNCH(k_mg_bump, 1, 32, 0, s, a, mb); break;
    default: break;
    }
}
