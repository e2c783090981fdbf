"""Bit-identity check of the band-partitioned multi-process run against a single-GPU run of the same small case.

Used by bench.py at N > 1 (before the timed region: the JSON line carries `mg_bit_identical`), by
tools/mg_nccl_check.py and by tests/mg_ipc_worker.py.  `dist` is an initialised torch.distributed module (any
backend: only *_object collectives are used); every rank calls `run_check` with its own device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import (_FAM_FIELDS, OPT_SHORT_HISTORY, Options, Sim, StepStats, mg_band_bounds, mg_unique_id, synth_families, synth_grid)

CASE = dict(W=64, H=400, families=6000, steps=30, hist_fill=60, hist_cap=96)


def run_check(lib, synth, dist, rank: int, world: int, device: int, transport: str = "peer", case: dict | None = None,
              use_nccl: bool = True, log=None, lag_rank=None):
    """Returns (ok, detail) on every rank.  transport: "peer" (CUDA IPC remote stores; with use_nccl=False no
    communicator is created at all) or "nccl"."""
    c = dict(CASE, **(case or {}))
    g, res, mx = synth_grid(c["W"], c["H"], 1, lib=synth)
    fam = synth_families(c["W"], c["H"], c["families"], 1, hist_fill=c["hist_fill"], lib=synth)

    def make(mg):
        opt = Options()
        lib.dsim_default_options(C.byref(opt))
        opt.device = device
        opt.history_capacity = c["hist_cap"]
        opt.flags |= OPT_SHORT_HISTORY
        opt.family_capacity = 4 * c["families"]
        s = Sim(lib, g, seed=7, options=opt)
        if mg:
            s.mg_configure(rank, world, mg_band_bounds(fam.location, g.n_nodes, world, s.mg_halo_width()))
        s.set_patches(res, mx)
        s.set_families(fam)
        return s

    sim = make(True)
    if use_nccl:
        uid = [mg_unique_id(lib) if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.mg_init_nccl(uid[0])            # communicator + initial halo over NCCL
    if transport == "peer":
        handles = [None] * world
        dist.all_gather_object(handles, sim.mg_ipc_export())
        sim.mg_ipc_connect(handles)
        if not use_nccl:
            dist.barrier()                   # every rank is connected before the first remote store
            sim.mg_halo_sync()               # initial halo over peer memory
    if lag_rank is None:
        st = sim.step(c["steps"], allow_model_errors=True)
    else:
        # steps issued one at a time, one rank's host sleeping before each: its device falls behind and the other ranks
        # run ahead as far as the messages let them (a rank three bands away gets a whole SPLIT round ahead: the case
        # the per-parity receive areas of the SPLIT channel exist for)
        import time
        st = None
        for _ in range(c["steps"]):
            if rank == lag_rank:
                time.sleep(0.02)
            s1 = sim.step(1, allow_model_errors=True)
            if st is None:
                st = s1
            else:
                st.births += s1.births
                st.deaths += s1.deaths
                st.model_errors |= s1.model_errors
                st.live_families = s1.live_families
    f = sim.get_families()
    mine = {name: getattr(f, name) for name, _ in _FAM_FIELDS}
    mine["hist"] = [tuple(map(np.ndarray.tolist, f.history_of(i))) for i in range(f.n)]
    r_, _ = sim.get_patches()
    lo, hi = int(g.eco_off[sim.mg_bounds[rank]]), int(g.eco_off[sim.mg_bounds[rank + 1]])
    mine["res"], mine["lo"], mine["hi"] = r_[lo:hi], lo, hi
    mine["stats"] = (st.live_families, st.births, st.deaths, st.model_errors)
    parts = [None] * world
    dist.gather_object(mine, parts if rank == 0 else None, dst=0)
    ok, detail = True, ""
    if rank == 0:
        ref = make(False)
        sr = ref.step(c["steps"], allow_model_errors=True)
        fr = ref.get_families()
        live, births, deaths = (sum(p["stats"][k] for p in parts) for k in range(3))
        detail = (f"world {world}, transport {transport}{'' if use_nccl else ' (no NCCL communicator)'}: live {live} (ref {sr.live_families}), "
                  f"births {births} (ref {sr.births}), per rank {[len(p['id']) for p in parts]}")
        ok &= live == sr.live_families and births == sr.births and deaths == sr.deaths
        ok &= all(p["stats"][3] == 0 for p in parts) and sr.model_errors == 0
        ids = np.concatenate([p["id"] for p in parts])
        order = np.argsort(ids)
        for name, _ in _FAM_FIELDS:
            a = np.concatenate([p[name] for p in parts])[order]
            b = getattr(fr, name)
            same = len(a) == len(b) and ((a.view(np.uint64) == b.view(np.uint64)).all() if a.dtype == np.float64 else (a == b).all())
            if not same:
                detail += f"; MISMATCH {name}"
                ok = False
        hist = [h for p in parts for h in p["hist"]]
        if ok:
            for k, i in enumerate(order):
                pr, hr = fr.history_of(k)
                if hist[i][0] != pr.tolist() or hist[i][1] != hr.tolist():
                    detail += f"; MISMATCH history of family {k}"
                    ok = False
                    break
        rr, _ = ref.get_patches()
        for p in parts:
            if (p["res"].view(np.uint64) != rr[p["lo"]:p["hi"]].view(np.uint64)).any():
                detail += "; MISMATCH resources"
                ok = False
        ref.close()
        if log:
            log("[mg check] " + detail + (" -> OK: bit-identical to the single-GPU run" if ok else " -> FAILED"))
    flag = [bool(ok), detail]
    dist.broadcast_object_list(flag, src=0)
    dist.barrier()                           # nobody frees its receive buffers while a neighbour may still store into them
    sim.close()
    return flag[0], flag[1]
