"""Worker of tests/test_multi_rank_ipc.py: one rank of a band-partitioned run with one PROCESS per rank, all on
cuda:0.  The step's three exchanges are the product's own peer-memory transport (k_mg_push: remote stores through CUDA
IPC mappings + release flags; k_mg_wait: acquire spins), replayed from the captured multi-GPU graph; no NCCL
communicator exists.  torch.distributed (gloo) only carries the IPC handles and the results for the comparison."""
import os
import sys

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
from helpers import D  # noqa: E402
from dsim_b200 import mgcheck  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    lib, synth = D.load(), helpers.load_synth()
    case = {"W": 48, "H": int(os.environ.get("MG_H", "240")), "families": 3000, "steps": int(os.environ.get("MG_STEPS", "12")),
            "hist_fill": 40}
    lag = os.environ.get("MG_LAG_RANK")
    ok, detail = mgcheck.run_check(lib, synth, dist, rank, world, 0, transport="peer", case=case, use_nccl=False,
                                   log=lambda m: print(m, flush=True), lag_rank=int(lag) if lag else None)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
