"""The C-ABI libraries load and export every symbol their headers declare (no compute, no GPU needed)."""
import ctypes as C
import os
import re

import pytest

import helpers
from helpers import D, REPO


def _declared(header, prefix):
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z_0-9]+)\s*\(", text)))


def test_product_library_exports_dsim_h():
    helpers._make(os.path.join(REPO, "dispersal-simulation_b200"))
    lib = C.CDLL(D.LIB_PATH)
    names = _declared("dsim.h", "dsim_")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dsim_abi_version() == 1


def test_product_library_exports_gavin_h():
    lib = C.CDLL(D.LIB_PATH)
    names = _declared("gavin.h", "gavin_")
    assert names == ["gavin_last_error", "gavin_run"]
    for n in names:
        assert hasattr(lib, n), n


def test_gavin_without_device_fails_loudly():
    """The batched gavin2017 model has no CPU fallback either: without a GPU gavin_run refuses."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import numpy as np
    from dsim_b200 import gavin as GB
    lib = C.CDLL(D.LIB_PATH)
    with pytest.raises(RuntimeError) as e:
        GB.run_sweep(lib, 4, 4, np.full((1, 32), 5.0), np.array([1.1]), np.array([1], np.uint64), np.array([100.0]))
    assert "-2" in str(e.value) and "no CPU fallback" in str(e.value)


def test_synth_library_exports_dsim_synth_h(synth):
    for n in _declared("dsim_synth.h", "dsim_synth_"):
        assert hasattr(synth, n), n


def test_oracle_mirrors_the_abi(oracle):
    """oracle/oracle.h exports the same entry points under the prefix oracle_ (test infrastructure)."""
    skip = {"dsim_profile_enable", "dsim_profile_read", "dsim_launch_count", "dsim_step_timed", "dsim_reach_digest",
            "dsim_observe_population", "dsim_observe_population_counts"}  # product-side observer: the oracle exposes the storage map it is derived from
    for n in _declared("dsim.h", "dsim_"):
        if n in skip or n.startswith("dsim_mg_"):   # multi-GPU entry points have no counterpart in the reference or the oracle
            continue
        assert hasattr(oracle, "oracle_" + n[len("dsim_"):]), n


def test_struct_layouts_match_header():
    assert C.sizeof(D.Params) == 88
    assert C.sizeof(D.Options) == 24
    assert C.sizeof(D.StepStats) == 56
    assert C.sizeof(D.FamiliesC) == 8 * 19
    assert C.sizeof(D.GraphC) == 8 * 10


def test_no_device_fails_loudly(synth):
    """Without a GPU dsim_create must refuse (DSIM_ERR_NO_DEVICE): there is no CPU fallback."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = C.CDLL(D.LIB_PATH)
    g, res, mx = D.synth_grid(6, 6, 1, lib=synth)
    with pytest.raises(D.DsimError) as e:
        D.Sim(lib, g, seed=1)
    assert e.value.code == -2, e.value
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_binding_has_no_fallback_path():
    src = open(os.path.join(REPO, "dispersal-simulation_b200", "python", "dsim_b200", "__init__.py")).read()
    assert "oracle" not in src.replace("the CPU oracle (prefix ``oracle_``)", "").replace("oracle_", "")
    assert "libdsim_emu" not in src
