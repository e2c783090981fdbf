"""The real multi-process data plane on the GPU: k_mg_push / k_mg_wait over CUDA IPC peer mappings and the captured
multi-GPU step graph, one process per rank (all on cuda:0, time-sliced), against the single-GPU state bit for bit.
world 4 also exercises the all-to-all SPLIT channel with its per-parity receive areas."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world,lag", [(2, None), (4, None), (4, 3)])
def test_peer_transport_processes_on_one_gpu(world, lag):
    """lag = a rank whose host sleeps before every single step: the others run ahead as far as the messages allow (the
    SPLIT channel's per-parity receive areas must keep a lagging far rank from reading the next round's lists)."""
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2", MG_H="240" if world == 2 else "400")
        if lag is not None:
            env["MG_LAG_RANK"] = str(lag)
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "mg_ipc_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=400)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out[-3000:]}"
    assert "bit-identical" in outs[0]
