"""CPU suite: the oracle's multi-step run / convergence loop against vectors recorded from the reference
(tests/golden/golden_converge.json, produced by tests/golden/make_golden_converge.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden_converge.json")) as f:
    CASES = json.load(f)["cases"]


def case_id(c):
    return f"{c['mode']}-side{c['side']}-seed{c['seed']}"


@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_oracle_run_matches_reference(c):
    world = oracle.initial_world(c["side"], c["seed"])
    stable = oracle.initial_stable(world, c["spawn"])
    if c["mode"] == "converge":
        n = oracle.run(world, stable, c["side"], c["spawn"], c["stable"], c["limit"] + 1, until_fixed=True)
    else:
        n = oracle.run(world, stable, c["side"], c["spawn"], c["stable"], c["limit"])
    assert n == c["steps"]
    assert int(oracle.reward(stable)) == c["reward"] and int(oracle.alive(world)) == c["alive"]
    assert hashlib.sha256(world.tobytes()).hexdigest() == c["world_sha"]
    assert hashlib.sha256(stable.tobytes()).hexdigest() == c["stable_sha"]
    assert oracle.breakdown(stable).tolist() == c["breakdown"]


def test_golden_has_converged_and_budget_limited_cases():
    conv = [c for c in CASES if c["mode"] == "converge"]
    assert any(c["steps"] < c["limit"] + 1 for c in conv) and any(c["steps"] == c["limit"] + 1 for c in conv)


# ---- the bit-sliced run algorithm of csrc/cgl_env_run.cu, restated on the host from the same header -------------
import ctypes  # noqa: E402
import subprocess  # noqa: E402

import numpy as np  # noqa: E402,F811

from conftest import ROOT  # noqa: E402

TWIN_DIR = os.path.join(ROOT, "tests", "twin")


@pytest.fixture(scope="module")
def twin():
    subprocess.run(["make", "-C", TWIN_DIR], check=True, stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(os.path.join(TWIN_DIR, "libcgl_twin.so"))
    vp, u32, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
    L.twin_pack.argtypes = [vp, vp, ctypes.c_uint64, u32, u32]
    L.twin_unpack.argtypes = [vp, vp, ctypes.c_uint64, u32, u32]
    L.twin_env_run_sliced.argtypes = [vp, vp, u32, u32, ci, ci, ci, ci, ci]
    L.twin_env_run_sliced.restype = ci
    return L


def _sliced_run(twin, cells, stable, side, max_steps, stop, spawn, smax, decay=0, emin=-128):
    words = np.zeros(side * side // 32, np.uint32)
    twin.twin_pack(cells.ctypes.data, words.ctypes.data, 1, side, side)
    st = stable.copy()
    n = twin.twin_env_run_sliced(words.ctypes.data, st.ctypes.data, side, max_steps, int(stop), spawn, smax, decay, emin)
    out = np.zeros(side * side, np.uint8)
    twin.twin_unpack(words.ctypes.data, out.ctypes.data, 1, side, side)
    return n, out, st


@pytest.mark.parametrize("c", [c for c in CASES if c["side"] % 32 == 0], ids=case_id)
def test_sliced_run_algorithm_matches_reference_vectors(twin, c):
    world = oracle.initial_world(c["side"], c["seed"])
    stable = oracle.initial_stable(world, c["spawn"])
    conv = c["mode"] == "converge"
    n, w, s = _sliced_run(twin, world, stable, c["side"], c["limit"] + conv, conv, c["spawn"], c["stable"])
    assert n == c["steps"]
    assert hashlib.sha256(w.tobytes()).hexdigest() == c["world_sha"]
    assert hashlib.sha256(s.tobytes()).hexdigest() == c["stable_sha"]


@pytest.mark.parametrize("decay", [0, 1])
@pytest.mark.parametrize("side,spawn,smax,emin", [(32, -2, 2, -6), (64, -128, 127, -128), (96, 5, 127, -3), (32, 0, 0, 0)])
def test_sliced_run_algorithm_random_states_vs_oracle(twin, decay, side, spawn, smax, emin):
    rs = np.random.RandomState(side + decay)
    for trial in range(4):
        dens = [0.05, 0.2, 0.5, 0.8][trial]
        world = (rs.random_sample(side * side) < dens).astype(np.uint8)
        stable = rs.randint(-128, 128, size=side * side).astype(np.int8)       # arbitrary bytes, dead cells too
        steps = int(rs.randint(1, 40))
        w_o, s_o = world.copy(), stable.copy()
        if decay:
            n_o = oracle.run_rule(w_o, s_o, side, spawn, smax, steps, oracle.DEAD_DECAY, 0, emin, until_fixed=True)
        else:
            n_o = oracle.run(w_o, s_o, side, spawn, smax, steps, until_fixed=True)
        n, w, s = _sliced_run(twin, world, stable, side, steps, True, spawn, smax, decay, emin)
        assert n == n_o and np.array_equal(w, w_o) and np.array_equal(s, s_o), (side, trial)
