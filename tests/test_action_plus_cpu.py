"""CPU suite for the CGL_action+ fork.

  * the oracle's fork functions replay every trace recorded from the fork's own CPU back end
    (tests/golden/golden_action_plus.npz): initial stability with `empty`, masked toggles (single and 2x2
    blocks from the fork's helper), the saturating dead-cell rule of its CPU step;
  * the product's 4-cells-per-word rule (cgl_bits.cuh built by g++ as tests/twin) equals its scalar rule
    exhaustively for all three dead-cell rules, and the scalar rule equals the oracle on the traces;
  * the decay rule of the fork's CUDA kernel (restated from the kernel text only) is checked for its
    defining properties against the oracle restatement.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import FORK_TRACES, ROOT
from oracle import oracle

TWIN_DIR = os.path.join(ROOT, "tests", "twin")


@pytest.fixture(scope="module")
def twin():
    subprocess.run(["make", "-C", TWIN_DIR], check=True, stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(os.path.join(TWIN_DIR, "libcgl_twin.so"))
    vp, u64, u32, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    L.twin_pack.argtypes = [vp, vp, u64, u32, u32]
    L.twin_unpack.argtypes = [vp, vp, u64, u32, u32]
    L.twin_life_generic.argtypes = [vp, vp, u64, u32, u32, ci]
    L.twin_stable_generic_rule.argtypes = [vp, vp, vp, u64, u32, ci, ci, ci, ci, ci]
    L.twin_rule_mismatches.argtypes = [ci, ci, ci, ci, ci]
    L.twin_rule_mismatches.restype = ctypes.c_uint64
    return L


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("name", sorted(FORK_TRACES))
def test_oracle_replays_fork_trace(name):
    tr = FORK_TRACES[name]
    world = tr.worlds[0].copy()
    stable = oracle.initial_stable_fork(world, tr.spawn, tr.empty)
    assert np.array_equal(stable, tr.stables[0])
    for t in range(tr.T):
        a = tr.action(t)
        if a is not None:
            oracle.toggle_masked(world, stable, a, tr.spawn)
        assert np.array_equal(stable, tr.toggled_stables[t]), (name, t)
        oracle.step_rule(world, stable, tr.side, tr.spawn, tr.stable_max, oracle.DEAD_SAT, tr.empty, tr.empty_min)
        assert np.array_equal(world, tr.worlds[t + 1]) and np.array_equal(stable, tr.stables[t + 1]), (name, t)
        assert int(oracle.reward(stable)) == tr.stability[t + 1] and int(oracle.alive(world)) == tr.alives[t + 1]


@pytest.mark.parametrize("name", sorted(FORK_TRACES))
def test_twin_scalar_rule_replays_fork_trace(twin, name):
    tr = FORK_TRACES[name]
    side, W = tr.side, (tr.side + 31) // 32
    cells = tr.worlds[0].copy()
    s = tr.stables[0].copy()
    for t in range(tr.T):
        a = tr.action(t)
        if a is not None:
            oracle.toggle_masked(cells, s, a, tr.spawn)
        w = np.zeros(side * W, np.uint32)
        twin.twin_pack(P(cells), P(w), 1, side, side)
        nxt = np.zeros_like(w)
        twin.twin_life_generic(P(w), P(nxt), 1, side, side, 1)
        twin.twin_stable_generic_rule(P(w), P(nxt), P(s), 1, side, tr.spawn, tr.stable_max, oracle.DEAD_SAT,
                                      tr.empty, tr.empty_min)
        cells = np.zeros(side * side, np.uint8)
        twin.twin_unpack(P(nxt), P(cells), 1, side, side)
        assert np.array_equal(cells, tr.worlds[t + 1]) and np.array_equal(s, tr.stables[t + 1]), (name, t)


@pytest.mark.parametrize("rule", [0, 1, 2])
@pytest.mark.parametrize("spawn,smax,empty,emin", [(-2, 2, 0, -128), (-2, 2, -1, -5), (-3, 4, -100, -90), (5, 127, 3, 7),
                                                   (-128, 127, 127, 127), (0, 0, -128, 0)])
def test_word_rule_equals_scalar_rule_exhaustively(twin, rule, spawn, smax, empty, emin):
    assert twin.twin_rule_mismatches(rule, spawn, smax, empty, emin) == 0


def test_decay_rule_properties():
    """CGL_action+/CGL.py:190-193: survivors and births as in the base env; every cell that is dead after the
    step moves one down per step until it sits at empty_min, and stays there."""
    side, spawn, smax, emin = 16, -2, 3, -6
    world = oracle.initial_world(side, 11)
    stable = oracle.initial_stable_fork(world, spawn, 0)
    base_w, base_s = world.copy(), oracle.initial_stable(world, spawn)
    for t in range(30):
        prev_w, prev_s = world.copy(), stable.copy()
        oracle.step_rule(world, stable, side, spawn, smax, oracle.DEAD_DECAY, 0, emin)
        oracle.step(base_w, base_s, side, spawn, smax)
        assert np.array_equal(world, base_w)                           # the world does not depend on the rule
        live = world != 0
        surv, born = live & (prev_w != 0), live & (prev_w == 0)
        assert np.array_equal(stable[born], np.full(born.sum(), spawn, np.int8))
        assert np.array_equal(stable[surv], np.where(prev_s[surv] == smax, prev_s[surv], prev_s[surv] + 1))
        dead = ~live
        want = np.where(prev_s[dead] == emin, prev_s[dead], (prev_s[dead].astype(np.int16) - 1).astype(np.int8))
        assert np.array_equal(stable[dead], want)
    assert stable[world == 0].min() == emin


def test_dead_zero_rule_is_the_base_env():
    side = 12
    world = oracle.initial_world(side, 2)
    a_w, a_s = world.copy(), oracle.initial_stable(world, -2)
    b_w, b_s = world.copy(), oracle.initial_stable_fork(world, -2, 0)
    for _ in range(12):
        oracle.step(a_w, a_s, side, -2, 2)
        oracle.step_rule(b_w, b_s, side, -2, 2, oracle.DEAD_ZERO)
        assert np.array_equal(a_w, b_w) and np.array_equal(a_s, b_s)


import hashlib  # noqa: E402
import json  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "golden_action_plus_converge.json")) as _f:
    FORK_RUNS = json.load(_f)["cases"]


@pytest.mark.parametrize("c", FORK_RUNS, ids=lambda c: f"{c['mode']}-side{c['side']}-seed{c['seed']}")
def test_oracle_run_rule_matches_the_forks_loops(c):
    """validate.py:133-139 / plain step loops recorded from the fork's CPU back end."""
    world = oracle.initial_world(c["side"], c["seed"])
    stable = oracle.initial_stable_fork(world, c["spawn"], c["empty"])
    n = oracle.run_rule(world, stable, c["side"], c["spawn"], c["stable"], c["limit"] + (c["mode"] == "converge"),
                        oracle.DEAD_SAT, c["empty"], c["empty_min"], until_fixed=c["mode"] == "converge")
    assert n == c["steps"] and int(oracle.reward(stable)) == c["stability"] and int(oracle.alive(world)) == c["alive"]
    assert hashlib.sha256(world.tobytes()).hexdigest() == c["world_sha"]
    assert hashlib.sha256(stable.tobytes()).hexdigest() == c["stable_sha"]
    assert oracle.breakdown(stable).tolist() == c["breakdown"]


@pytest.mark.parametrize("name", [n for n in sorted(FORK_TRACES) if FORK_TRACES[n].side % 32 == 0])
def test_twin_fused_rule_path_replays_fork_trace(twin, name):
    """The fused kernel's logic with the fork's variants (nibble tables -> 4-cells-per-word rule, masked toggle
    applied inside the step), restated on the host from the same header."""
    twin.twin_env_step_fused_rule.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p] + \
        [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
    twin.twin_env_step_fused_rule.restype = ctypes.c_int
    tr = FORK_TRACES[name]
    side, W = tr.side, tr.side // 32
    w = np.zeros(side * W, np.uint32)
    cells0 = np.ascontiguousarray(tr.worlds[0])          # keep the buffer alive across the ctypes call
    twin.twin_pack(P(cells0), P(w), 1, side, side)
    s = tr.stables[0].copy()
    for t in range(tr.T):
        a = tr.action(t)
        assert a is None or isinstance(a, int)
        acts = np.array([tr.size if a is None else a], np.int32)
        nxt, rew, alv = np.zeros_like(w), np.zeros(1, np.int32), np.zeros(1, np.uint32)
        rc = twin.twin_env_step_fused_rule(P(w), P(nxt), P(s), 1, side, P(acts), tr.spawn, tr.stable_max, oracle.DEAD_SAT,
                                           tr.empty, tr.empty_min, 1, P(rew), P(alv))
        assert rc == 0
        w = nxt
        cells = np.zeros(side * side, np.uint8)
        twin.twin_unpack(P(w), P(cells), 1, side, side)
        assert np.array_equal(cells, tr.worlds[t + 1]) and np.array_equal(s, tr.stables[t + 1]), (name, t)
        assert int(rew[0]) == tr.stability[t + 1] and int(alv[0]) == tr.alives[t + 1]


@pytest.mark.parametrize("rule", [0, 1, 2])
def test_twin_fused_rule_path_random_vs_oracle(twin, rule):
    twin.twin_env_step_fused_rule.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p] + \
        [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
    twin.twin_env_step_fused_rule.restype = ctypes.c_int
    rs = np.random.RandomState(rule)
    for side in (32, 64, 96):
        size, W = side * side, side // 32
        spawn, smax, empty, emin = -2, 3, -1, -7
        world = rs.randint(2, size=size).astype(np.uint8)
        stable = rs.randint(-128, 128, size=size).astype(np.int8)
        w = np.zeros(side * W, np.uint32)
        twin.twin_pack(P(world), P(w), 1, side, side)
        s = stable.copy()
        for t in range(6):
            a = int(rs.randint(size + 1))
            acts = np.array([a], np.int32)
            nxt, rew, alv = np.zeros_like(w), np.zeros(1, np.int32), np.zeros(1, np.uint32)
            assert twin.twin_env_step_fused_rule(P(w), P(nxt), P(s), 1, side, P(acts), spawn, smax, rule, empty, emin, 1,
                                                 P(rew), P(alv)) == 0
            w = nxt
            oracle.toggle_masked(world, stable, a, spawn)
            oracle.step_rule(world, stable, side, spawn, smax, rule, empty, emin)
            cells = np.zeros(size, np.uint8)
            twin.twin_unpack(P(w), P(cells), 1, side, side)
            assert np.array_equal(cells, world) and np.array_equal(s, stable), (rule, side, t)
            assert int(rew[0]) == int(oracle.reward(stable))
