"""Worker of tests/test_multi_rank_gloo.py: one rank of a 2-process band-partitioned run on the CPU.
Kernels: the test-only emulator build; transport: torch.distributed (gloo) send/recv of the channel buffers."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
from helpers import D  # noqa: E402


def tensor_of(ptr, nbytes):
    return torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr)))


def exchange(sim, channel, rank, world):
    if channel == D.MG_CH_SPLIT:
        sp, nb = sim.mg_buffer(channel, 0, False)
        mine = tensor_of(sp, nb)
        outs = [tensor_of(sim.mg_buffer(channel, r, True)[0], nb) for r in range(world)]
        dist.all_gather(outs, mine)
        return
    reqs = []
    for side, peer in ((0, rank - 1), (1, rank + 1)):
        if peer < 0 or peer >= world:
            continue
        sp, nb = sim.mg_buffer(channel, side, False)
        rp, _ = sim.mg_buffer(channel, side, True)
        reqs.append(dist.isend(tensor_of(sp, nb), peer))
        reqs.append(dist.irecv(tensor_of(rp, nb), peer))
    for r in reqs:
        r.wait()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    oracle, synth, emu = helpers.load_oracle(), helpers.load_synth(), helpers.load_emu()
    W, H, NF, STEPS = 16, 60, 220, 8
    g, res, mx = D.synth_grid(W, H, 1, lib=synth)
    fam = D.synth_families(W, H, NF, 1, hist_fill=20, lib=synth)
    opt = D.Options()
    emu.dsim_default_options(C.byref(opt))
    opt.history_capacity = 96
    opt.flags |= D.OPT_SHORT_HISTORY
    opt.family_capacity = 4 * NF + 1024
    sim = D.Sim(emu, g, seed=7, options=opt)
    bounds = D.mg_band_bounds(fam.location, g.n_nodes, world, sim.mg_halo_width())
    sim.mg_configure(rank, world, bounds)
    sim.set_patches(res, mx)
    sim.set_families(fam)
    sim.mg_phase(D.MG_PHASE_HALO_PACK)
    exchange(sim, D.MG_CH_HALO, rank, world)
    sim.mg_phase(D.MG_PHASE_HALO_UNPACK)
    for _ in range(STEPS):
        sim.mg_phase(D.MG_PHASE_LIFECYCLE)
        exchange(sim, D.MG_CH_SPLIT, rank, world)
        sim.mg_phase(D.MG_PHASE_MOVE)
        exchange(sim, D.MG_CH_EMIGRANTS, rank, world)
        sim.mg_phase(D.MG_PHASE_PATCH)
        exchange(sim, D.MG_CH_HALO, rank, world)
        sim.mg_phase(D.MG_PHASE_FINISH)
    f = sim.get_families(history=False, adaptation=False)
    mine = {name: getattr(f, name) for name, _ in D._FAM_FIELDS}
    parts = [None] * world
    dist.gather_object(mine, parts if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        ref = helpers.oracle_sim(oracle, g, seed=7)
        ref.set_patches(res, mx)
        ref.set_families(fam)
        ref.step(STEPS, allow_model_errors=True)
        fr = ref.get_families(history=False, adaptation=False)
        ids = np.concatenate([p["id"] for p in parts])
        order = np.argsort(ids)
        for name, _ in D._FAM_FIELDS:
            a = np.concatenate([p[name] for p in parts])[order]
            b = getattr(fr, name)
            same = len(a) == len(b) and ((a.view(np.uint64) == b.view(np.uint64)).all() if a.dtype == np.float64 else (a == b).all())
            if not same:
                print("MISMATCH", name, flush=True)
                ok = False
        assert all(len(p["id"]) > 0 for p in parts)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
