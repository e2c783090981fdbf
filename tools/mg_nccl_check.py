#!/usr/bin/env python
"""Run under torchrun with N ranks: the NCCL band-partitioned run must equal a single-GPU run bit for bit.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mg_nccl_check.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "dispersal-simulation_b200", "python"))
import dsim_b200 as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib, synth = D.load(), D.load_synth()
W, H, NF, STEPS, HIST = 64, 400, 6000, int(os.environ.get("STEPS", "30")), 60
g, res, mx = D.synth_grid(W, H, 1, lib=synth)
fam = D.synth_families(W, H, NF, 1, hist_fill=HIST, lib=synth)


def make(device, mg):
    opt = D.Options()
    lib.dsim_default_options(C.byref(opt))
    opt.device = device
    opt.history_capacity = 96
    opt.family_capacity = 4 * NF
    s = D.Sim(lib, g, seed=7, options=opt)
    if mg:
        bounds = D.mg_band_bounds(fam.location, g.n_nodes, world, s.mg_halo_width())
        s.mg_configure(rank, world, bounds)
    s.set_patches(res, mx)
    s.set_families(fam)
    return s


sim = make(local, True)
uid = [D.mg_unique_id(lib) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
sim.mg_init_nccl(uid[0])
if os.environ.get("DSIM_MG_TRANSPORT", "peer") == "peer":
    handles = [None] * world
    dist.all_gather_object(handles, sim.mg_ipc_export())
    sim.mg_ipc_connect(handles)
    if rank == 0:
        print("[mg check] transport: peer memory (CUDA IPC over NVLink)")
st = sim.step(STEPS, allow_model_errors=True)
tot = sim.mg_reduce_stats(D.StepStats.from_buffer_copy(st))
f = sim.get_families()
mine = {name: getattr(f, name) for name, _ in D._FAM_FIELDS}
mine["hist"] = [tuple(map(np.ndarray.tolist, f.history_of(i))) for i in range(f.n)]
r_, _ = sim.get_patches()
lo, hi = int(g.eco_off[sim.mg_bounds[rank]]), int(g.eco_off[sim.mg_bounds[rank + 1]])
mine["res"], mine["lo"], mine["hi"] = r_[lo:hi], lo, hi
parts = [None] * world
dist.gather_object(mine, parts if rank == 0 else None, dst=0)
ok = True
if rank == 0:
    ref = make(local, False)
    sr = ref.step(STEPS, allow_model_errors=True)
    fr = ref.get_families()
    ids = np.concatenate([p["id"] for p in parts])
    order = np.argsort(ids)
    print(f"[mg check] world {world}: live {tot.live_families} (ref {sr.live_families}), births {tot.births} (ref {sr.births}), "
          f"per rank {[len(p['id']) for p in parts]}")
    ok &= tot.live_families == sr.live_families and tot.births == sr.births and tot.deaths == sr.deaths
    for name, _ in D._FAM_FIELDS:
        a = np.concatenate([p[name] for p in parts])[order]
        b = getattr(fr, name)
        same = (a.view(np.uint64) == b.view(np.uint64)).all() if a.dtype == np.float64 else (a == b).all()
        if not same:
            print("MISMATCH", name)
            ok = False
    hist = [h for p in parts for h in p["hist"]]
    for k, i in enumerate(order):
        pr, hr = fr.history_of(k)
        if hist[i][0] != pr.tolist() or hist[i][1] != hr.tolist():
            print("MISMATCH history of family", k)
            ok = False
            break
    rr, _ = ref.get_patches()
    for p in parts:
        if (p["res"].view(np.uint64) != rr[p["lo"]:p["hi"]].view(np.uint64)).any():
            print("MISMATCH resources")
            ok = False
    print("[mg check]", "OK: bit-identical to the single-GPU run" if ok else "FAILED")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
