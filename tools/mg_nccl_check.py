#!/usr/bin/env python
"""Run under torchrun with N ranks: the band-partitioned run must equal a single-GPU run bit for bit.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mg_nccl_check.py
   DSIM_MG_TRANSPORT=peer (default) | nccl;  DSIM_MG_NO_NCCL=1: peer memory without any communicator."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "dispersal-simulation_b200", "python"))
import dsim_b200 as D  # noqa: E402
from dsim_b200 import mgcheck  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok, detail = mgcheck.run_check(D.load(), D.load_synth(), dist, rank, world, local,
                               transport=os.environ.get("DSIM_MG_TRANSPORT", "peer"),
                               case={"steps": int(os.environ.get("STEPS", "30"))},
                               use_nccl=os.environ.get("DSIM_MG_NO_NCCL", "0") != "1",
                               log=lambda m: print(m, flush=True))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
