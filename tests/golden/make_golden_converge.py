#!/usr/bin/env python
"""Golden vectors for the multi-step run / convergence loop, produced by RUNNING THE REFERENCE.

Build container only (needs /root/reference):

    GPU_CAPABLE=false PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_converge.py

Drives the unmodified /root/reference/CGL/CGL.py `sim(gpu=False)` through
  * the plain step loop of CGL/bench.py:39-40, and
  * the convergence loop of CGL/CGL_action+/validate.py:133-139 (written out below with the same
    statements: old = get_state(); step(); while not match(old) and count_down: ...),
and records the number of steps executed, reward, alive, SHA-256 of world and stability, and the
np.unique value counts of the stability vector (what breakdown_stable returns, CGL_action+/CGL.py:294-297).
Output: golden_converge.json.
"""
import contextlib
import hashlib
import io
import json
import os
import sys
import warnings

os.environ["GPU_CAPABLE"] = "false"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/CGL")
import numpy as np  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import CGL  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore", category=RuntimeWarning)


def make_sim(**kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return CGL.sim(gpu=False, **kw)


def record(env, steps):
    unique, counts = np.unique(env.stable, return_counts=True)
    return {"steps": steps, "reward": int(env.reward()), "alive": int(env.alive()),
            "world_sha": hashlib.sha256(env.world.tobytes()).hexdigest(),
            "stable_sha": hashlib.sha256(env.stable.tobytes()).hexdigest(),
            "breakdown": [[int(u) for u in unique], [int(c) for c in counts]]}


def main():
    cases = []
    # (side, seed, spawn, stable, limit): convergence loop, CONVERGENCE_LIMIT = limit
    for side, seed, spawn, stable, limit in [(10, 0, -2, 2, 300), (10, 1, -2, 2, 300), (10, 2, -1, 1, 300),
                                             (10, 5, -2, 2, 300), (16, 3, -2, 2, 400), (16, 4, -3, 127, 400),
                                             (7, 0, -2, 2, 100), (32, 0, -2, 2, 500), (32, 6, -2, 2, 500),
                                             (33, 1, -2, 2, 150), (64, 0, -2, 2, 60), (64, 2, 100, 120, 40)]:
        env = make_sim(side=side, seed=seed, spawnStabilityFactor=spawn, stableStabilityFactor=stable)
        old = env.get_state()
        env.step()
        steps = 1
        count_down = limit
        while not env.match(old) and count_down:
            old = env.get_state()
            env.step()
            steps += 1
            count_down -= 1
        c = {"mode": "converge", "side": side, "seed": seed, "spawn": spawn, "stable": stable, "limit": limit}
        c.update(record(env, steps))
        cases.append(c)
        print(c["mode"], side, seed, "steps", steps, "alive", c["alive"], file=sys.stderr)
    # plain step loop: bench.py:39-40
    for side, seed, spawn, stable, iters in [(32, 1, -2, 2, 130), (64, 1, -2, 2, 50), (96, 0, -2, 2, 12),
                                             (128, 0, -2, 2, 10), (12, 3, -1, 1, 40), (40, 2, -2, 2, 30),
                                             (32, 9, -2, 2, 0)]:
        env = make_sim(side=side, seed=seed, spawnStabilityFactor=spawn, stableStabilityFactor=stable)
        for _ in range(iters):
            env.step()
        c = {"mode": "steps", "side": side, "seed": seed, "spawn": spawn, "stable": stable, "limit": iters}
        c.update(record(env, iters))
        cases.append(c)
        print(c["mode"], side, seed, "steps", iters, "alive", c["alive"], file=sys.stderr)
    with open(os.path.join(HERE, "golden_converge.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_converge.py", "cases": cases}, f, indent=1)


if __name__ == "__main__":
    main()
