"""Golden trajectories (tests/golden/*.npz, generated from the oracle by tests/golden/make_golden.py).

CPU: the oracle still reproduces them (mode 0 = reference-shaped and mode 1 = memoised reach), and so do
the CUDA kernels under the test-only emulator.  GPU: the real library reproduces them bit for bit."""
import os

import numpy as np
import pytest

import helpers
from helpers import D

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402


def _check(out, name):
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    for key in gold.files:
        a, b = gold[key], out[key]
        assert a.shape == b.shape, (name, key, a.shape, b.shape)
        if a.dtype == np.float64:
            assert (a.view(np.uint64) == b.view(np.uint64)).all(), (name, key)
        else:
            assert (a == b).all(), (name, key)


@pytest.mark.parametrize("name", sorted(mg.CASES))
@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_reproduces_golden(oracle, synth, name, mode):
    out = mg.run_case(oracle, synth, name, sim_factory=lambda g, p: helpers.oracle_sim(oracle, g, params=p, seed=11, mode=mode))
    _check(out, name)


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_emulated_kernels_reproduce_golden(oracle, emu, synth, name):
    out = mg.run_case(oracle, synth, name, sim_factory=lambda g, p: D.Sim(emu, g, params=p, seed=11))
    _check(out, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_gpu_reproduces_golden(oracle, gpu_lib, synth, name):
    out = mg.run_case(oracle, synth, name, sim_factory=lambda g, p: D.Sim(gpu_lib, g, params=p, seed=11))
    _check(out, name)
