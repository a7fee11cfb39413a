"""CPU suite: pin the oracle (C port + numpy restatement) to the reference's golden vectors."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, TRACES
from oracle import oracle


@pytest.mark.parametrize("name", sorted(TRACES))
def test_c_oracle_replays_reference_trace(name):
    tr = TRACES[name]
    w = tr.worlds[0].copy()
    s = tr.stables[0].copy()
    assert oracle.reward(s) == tr.rewards[0] and oracle.alive(w) == tr.alives[0]
    for t in range(tr.T):
        a = tr.action(t)
        if a is not None:
            oracle.toggle(w, s, a, tr.spawn)
        assert oracle.reward(s) == tr.rewards_after_toggle[t]
        oracle.step(w, s, tr.side, tr.spawn, tr.stable_max)
        assert np.array_equal(w, tr.worlds[t + 1]), (name, t)
        assert np.array_equal(s, tr.stables[t + 1]), (name, t)
        assert oracle.reward(s) == tr.rewards[t + 1]
        assert oracle.alive(w) == tr.alives[t + 1]


@pytest.mark.parametrize("name", sorted(TRACES))
def test_numpy_restatement_replays_reference_trace(name):
    tr = TRACES[name]
    w = tr.worlds[0].copy()
    s = tr.stables[0].copy()
    for t in range(tr.T):
        a = tr.action(t)
        if a is not None:
            oracle.toggle(w, s, a, tr.spawn)
        w, s = oracle.numpy_step(w, s, tr.side, tr.spawn, tr.stable_max)
        assert np.array_equal(w, tr.worlds[t + 1]) and np.array_equal(s, tr.stables[t + 1]), (name, t)


def test_survey_known_answers():
    """Numbers recorded in SURVEY.md section 8c from the reference's CPU step."""
    tr = TRACES["blinker5"]                      # CGL/bench.py --cpu --blink
    assert tr.rewards[256] == -129 and tr.alives[256] == 3
    assert tr.stables[255].reshape(5, 5)[1:4, 2].tolist() == [-128, 127, -128]
    assert TRACES["block4"].rewards[1:].tolist() == [-4, 0, 4, 8, 8, 8]
    tr = TRACES["rand64_plain"]
    assert (tr.alives[0], tr.rewards[0]) == (2083, -4166)
    assert list(zip(tr.alives[1:].tolist(), tr.rewards[1:].tolist())) == [
        (1079, -1507), (937, -1062), (918, -925), (909, -891), (876, -835),
        (835, -760), (778, -696), (811, -715), (765, -664), (772, -691)]


def test_initial_state_matches_reference_rng():
    tr = TRACES["rand64_plain"]
    assert np.array_equal(oracle.initial_world(64, 0), tr.worlds[0])
    assert np.array_equal(oracle.initial_stable(tr.worlds[0], -2), tr.stables[0])


def test_batch_matches_single_and_threads():
    side, B = 64, 6
    w = np.stack([TRACES[f"env64_{e}"].worlds[0] for e in range(B)]).copy()
    s = np.stack([TRACES[f"env64_{e}"].stables[0] for e in range(B)]).copy()
    for t in range(5):
        acts = np.array([TRACES[f"env64_{e}"].actions[t, 0] for e in range(B)], np.int32)
        rew, alv = oracle.step_batch(w, s, side, acts, -2, 2, threads=3)
        for e in range(B):
            tr = TRACES[f"env64_{e}"]
            assert np.array_equal(w[e], tr.worlds[t + 1]) and np.array_equal(s[e], tr.stables[t + 1])
            assert rew[e] == tr.rewards[t + 1] and alv[e] == tr.alives[t + 1]
    with pytest.raises(ValueError):
        oracle.step_batch(w, s, side, np.full(B, side * side + 1, np.int32), -2, 2)


def test_life_mode_matches_env_world_plane():
    tr = TRACES["rand64_plain"]
    out = oracle.life(tr.worlds[0].reshape(64, 64), 10, threads=2)
    assert np.array_equal(out.reshape(-1), tr.worlds[10])
    # open-window light cone: the centre of a window cut from the torus is exact
    g, h = 3, 20
    W0 = tr.worlds[0].reshape(64, 64)
    win = W0[10 - g:10 + h + g, 7 - g:7 + h + g]
    got = oracle.life_open(win, g)[g:g + h, g:g + h]
    assert np.array_equal(got, tr.worlds[g].reshape(64, 64)[10:10 + h, 7:7 + h])


def test_toggle_errors_follow_reference():
    api = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_api.json")))
    w = np.array(api["side6_seed3_world"], np.uint8)
    s = oracle.initial_stable(w, -1)
    w0 = w.copy()
    oracle.toggle(w, s, [3, 3], -1)
    assert (w != w0).sum() == 1 and api["toggle_dup_flips_once"]
    oracle.toggle(w, s, 36, -1)
    assert (w != w0).sum() == 1
    for bad in (37, -1, [1, 99]):
        with pytest.raises(ValueError):
            oracle.toggle(w, s, bad, -1)
    oracle.toggle(w, s, [36], -1)   # single-element list == size is the silent no-op
    assert api["toggle_list_noop_single"] is None
