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
