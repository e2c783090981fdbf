"""The callers either side of the step loop: `run` (src/cli.rs:123-156), the
stdout log protocol that supplement/plotting/plot.py:290-381 parses, and state files for resuming
(lib.rs:324-346).  CPU: through the emulated kernels; GPU: through the product library."""
import ctypes as C
import io
import os

import numpy as np
import pytest

import helpers
import parity_cases as pc
from helpers import D
from dsim_b200 import driver


def _sim(lib, synth, prefix, n_fam=150, params_edit=None):
    g, res, mx = D.synth_grid(20, 24, 1, lib=synth)
    fam = D.synth_families(20, 24, n_fam, 1, hist_fill=20, lib=synth)
    p = D.Params()
    getattr(lib, prefix + "default_params")(C.byref(p))
    if params_edit:
        params_edit(p)
    opt = D.Options()
    getattr(lib, prefix + "default_options")(C.byref(opt))
    opt.history_capacity = 96
    opt.flags |= D.OPT_SHORT_HISTORY
    s = D.Sim(lib, g, params=p, seed=7, options=opt, prefix=prefix)
    s.set_patches(res, mx)
    s.set_families(fam)
    return s, g, opt


def _check_log_protocol(lib, synth, prefix, oracle):
    sim, g, _ = _sim(lib, synth, prefix)
    out = io.StringIO()
    end = driver.run(sim, 9, driver.Settings(log_every=3, store_every=0, statefile=os.devnull), out)
    text = out.getvalue()
    assert end == "Ended" and text.endswith("Ended\n")
    steps, path = driver.parse_log(text)
    assert [t for t, _ in steps] == [0, 3, 6, 9], "logged when t % log_every == 0 (lib.rs:649)"
    assert len(path) == 4
    # the logged populations are the storage map of the oracle after the same number of steps
    o = helpers.oracle_sim(oracle, g, params=sim.params, seed=7, mode=1)
    res, mx = sim.get_patches()
    g0, res0, mx0 = D.synth_grid(20, 24, 1, lib=synth)
    o.set_patches(res0, mx0)
    o.set_families(D.synth_families(20, 24, 150, 1, hist_fill=20, lib=synth))
    done = 0
    for t, content in steps:
        o.step(t + 1 - done, allow_model_errors=True)
        done = t + 1
        node, cul, size = o.get_storage()
        want = sorted((float(g.lon[v]), float(g.lat[v]), int(n), int(c)) for v, c, n in zip(node, cul, size))
        assert sorted(content) == want, f"POPULATION line of t={t}"
    fo = o.get_families(history=False, adaptation=False)
    assert path[-1][1] == int(fo.location[0])
    # ... and the reference's own parser expressions (taken out of plot.py's syntax tree, tests/golden/log_parser.json)
    # read the same out of the log: dispatch on the line prefixes as plot.py:290-370 does
    import json
    import re
    from ast import literal_eval
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "log_parser.json")))
    bitvec = re.compile(ref["bitvec_pattern"])
    season = sim.params.season_length_in_years
    ref_steps, ref_path, timestamp = [], [], None
    for line in io.StringIO(text):
        if line.startswith("t: "):
            timestamp = eval(ref["timestamp_expr"], {"line": line, "season_length_in_years": season, "int": int})
        elif line.startswith("POPULATION: ["):
            line = bitvec.sub(ref["bitvec_repl"], line)
            ref_steps.append((timestamp, eval(ref["content_expr"], {"line": line, "literal_eval": literal_eval, "len": len, "int": int})))
        elif line.startswith("F:"):
            match = re.search(ref["f_regex"], line)
            assert match, line
            assert match.group(3) and int(match.group(3), 2) < 2 ** 64, "plot.py's own expression reads the culture bits (src/debug.rs:17, lib.rs:260-272)"
            ref_path.append((match.group(1), int(match.group(2))))
    assert [c for _, c in ref_steps] == [c for _, c in steps]
    assert [ts for ts, _ in ref_steps] == [t * season for t, _ in steps]
    assert ref_path == path
    known = tuple(ref["prefixes"])
    other = [line for line in io.StringIO(text) if line.strip() and not line.startswith(known)]
    assert all(re.fullmatch(r"\d+ % \d+\n", line) for line in other), other  # only the `t % store_every` trace of src/cli.rs:142-149


def _check_resume(lib, synth, prefix, tmp_path):
    statefile = str(tmp_path / "state.npz")
    a, g, opt = _sim(lib, synth, prefix)
    driver.run(a, 4, driver.Settings(log_every=0, store_every=0, statefile=statefile), io.StringIO())  # stores at t == max_t
    assert os.path.exists(statefile)
    a.step(5, allow_model_errors=True)
    b = driver.load_state(lib, g, statefile, options=opt, prefix=prefix)
    assert b.get_time() == 5
    b.step(5, allow_model_errors=True)
    msgs = helpers.compare_sims(a, b, hist_cap=96, what="resumed run")
    assert not msgs, "\n".join(msgs)


def _check_died_out(lib, synth, prefix):
    def starve(p):
        p.maximum_resources_one_adult_can_harvest = 1e-6
    sim, _, _ = _sim(lib, synth, prefix, n_fam=30, params_edit=starve)
    out = io.StringIO()
    assert driver.run(sim, 200, driver.Settings(log_every=50, statefile=os.devnull), out) == "Died out"
    assert out.getvalue().endswith("Died out\n")


def test_log_protocol_emu(oracle, emu, synth):
    _check_log_protocol(emu, synth, "dsim_", oracle)


def test_resume_emu(emu, synth, tmp_path):
    _check_resume(emu, synth, "dsim_", tmp_path)


def test_died_out_emu(emu, synth):
    _check_died_out(emu, synth, "dsim_")


def test_rust_float_format():
    assert driver._rust_f64(45.0) == "45.0" and driver._rust_f64(-120.5) == "-120.5"
    assert driver._rust_f64(1e21) == "1e21" and driver._rust_f64(1e-7) == "1e-7"
    assert driver._oyr_debug(1.5) == "1.50 (1095000.00 kCal)"


@pytest.mark.gpu
def test_log_protocol_gpu(oracle, gpu_lib, synth):
    _check_log_protocol(gpu_lib, synth, "dsim_", oracle)


@pytest.mark.gpu
def test_resume_gpu(gpu_lib, synth, tmp_path):
    _check_resume(gpu_lib, synth, "dsim_", tmp_path)


@pytest.mark.gpu
def test_died_out_gpu(gpu_lib, synth):
    _check_died_out(gpu_lib, synth, "dsim_")
