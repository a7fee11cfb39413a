#!/usr/bin/env python
"""Golden vectors for the CGL_action+ fork, produced by RUNNING THE FORK ITSELF (its CPU back end).

Build container only (needs /root/reference):

    GPU_CAPABLE=false PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_action_plus.py

Imports the unmodified /root/reference/CGL/CGL_action+/CGL.py and helper.py, drives `sim(gpu=False)`
(__step_state_cpu, CGL_action+/CGL.py:231-257: dead cells get min(stable + empty, empty_min)), its masked
toggle_state (:378-386) with single indices and with the 2x2 blocks of helper.take_action (:108-132), and
records every world / stability / stability() / alive().  The fork's CUDA kernel (:159-196) applies a
DIFFERENT dead-cell rule and cannot run here (PyCUDA + a GPU); it is restated from its text only.
Output: golden_action_plus.npz.
"""
import contextlib
import io
import os
import sys
import warnings

os.environ["GPU_CAPABLE"] = "false"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/CGL/CGL_action+")
import numpy as np  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import CGL  # noqa: E402  (the fork)
    import helper  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore", category=RuntimeWarning)   # int8 wrap in `stable + empty` is part of the behaviour


def make_sim(**kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return CGL.sim(gpu=False, **kw)


def trace(name, out, sim, actions):
    """actions: list of None | int | list[int] applied (toggle_state) before each step."""
    T, size = len(actions), sim.size
    worlds = np.zeros((T + 1, size), np.uint8)
    stables = np.zeros((T + 1, size), np.int8)
    stab = np.zeros(T + 1, np.int32)
    alives = np.zeros(T + 1, np.int32)
    toggled_stables = np.zeros((T, size), np.int8)            # stability right after toggle_state
    K = max([1] + [len(a) for a in actions if isinstance(a, (list, tuple))])
    acts = np.full((T, K), -1, np.int64)                      # -1 = padding / no toggle_state call
    worlds[0], stables[0], stab[0], alives[0] = sim.world, sim.stable, sim.stability(), sim.alive()
    for t, a in enumerate(actions):
        if a is not None:
            if isinstance(a, (list, tuple)):
                acts[t, :len(a)] = a
                sim.toggle_state(list(a))
            else:
                acts[t, 0] = a
                sim.toggle_state(np.int32(a))
        toggled_stables[t] = sim.stable
        sim.step()
        worlds[t + 1], stables[t + 1], stab[t + 1], alives[t + 1] = sim.world, sim.stable, sim.stability(), sim.alive()
    for key, val in (("side", sim.side), ("spawn", sim.spawnStabilityFactor), ("stable_max", sim.stableStabilityFactor),
                     ("empty", sim.empty), ("empty_min", sim.empty_min)):
        out[f"{name}/{key}"] = np.int64(val)
    out[f"{name}/worlds"] = np.packbits(worlds, axis=1)
    out[f"{name}/stables"] = stables
    out[f"{name}/toggled_stables"] = toggled_stables
    out[f"{name}/stability"] = stab
    out[f"{name}/alives"] = alives
    out[f"{name}/actions"] = acts
    out[f"{name}/max_density"] = np.float64(sim.max_density)
    assert worlds.max() <= 1
    print(f"{name:24s} side={sim.side:4d} T={T:3d} stability[-1]={stab[-1]} alive[-1]={alives[-1]}", file=sys.stderr)


def main():
    out = {}
    # defaults of the fork (empty 0, empty_min -128): dead cells drop to -128 on their first step
    trace("default32", out, make_sim(side=32, seed=1, spawnStabilityFactor=-2, stableStabilityFactor=2), [None] * 8)
    # a gentler floor with a negative `empty` (stable + empty stays inside int8)
    rs = np.random.RandomState(5)
    trace("floor64", out, make_sim(side=64, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2, empty=-1, empty_min=-5),
          [int(rs.randint(64 * 64 + 1)) for _ in range(10)])
    # int8 wrap of stable + empty BEFORE the minimum (NumPy 2 scalar arithmetic)
    trace("wrap16", out, make_sim(side=16, seed=2, spawnStabilityFactor=-3, stableStabilityFactor=4, empty=-100, empty_min=-90),
          [None] * 6)
    trace("posempty10", out, make_sim(side=10, seed=4, spawnStabilityFactor=-2, stableStabilityFactor=2, empty=3, empty_min=7),
          [int(a) for a in np.random.RandomState(6).randint(101, size=12)])
    # blank start (runBlank) driven only by 2x2 block actions from the fork's own helper
    sim = make_sim(side=12, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2, runBlank=True, empty=-1, empty_min=-4)
    rs = np.random.RandomState(9)
    blocks = []
    for _ in range(14):
        action, _good = helper.take_action(int(rs.randint(12 * 12 + 1)), 12, sim.world)
        blocks.append([int(a) for a in action] if len(action) > 1 else int(action[0]))
    trace("blank12_blocks", out, sim, blocks)
    # ragged and tiny tori, single toggles incl. the no-op, spawn 0 (the `stable == 0 -> empty` quirk)
    for side, spawn in ((5, -2), (7, 0), (33, -2), (40, -1)):
        rs = np.random.RandomState(100 + side)
        trace(f"tiny{side}", out, make_sim(side=side, seed=side, spawnStabilityFactor=spawn, stableStabilityFactor=3,
                                           empty=-2, empty_min=-6),
              [int(rs.randint(side * side + 1)) for _ in range(8)])
    trace("fused128", out, make_sim(side=128, seed=3, spawnStabilityFactor=-2, stableStabilityFactor=2, empty=-1, empty_min=-3),
          [7, None, 128 * 128, 16383])
    np.savez_compressed(os.path.join(HERE, "golden_action_plus.npz"), **out)
    converge_cases()


def converge_cases():
    """The fork's validation loop (CGL_action+/validate.py:133-139) on the fork's env, statement for statement;
    also plain step loops.  Output: golden_action_plus_converge.json."""
    import hashlib
    import json
    cases = []
    for mode, side, seed, spawn, stable, empty, emin, limit in [
            ("converge", 10, 0, -2, 2, 0, -128, 300), ("converge", 10, 1, -2, 2, -1, -5, 300),
            ("converge", 10, 5, -2, 2, -1, -5, 300), ("converge", 16, 3, -2, 2, -2, -9, 400),
            ("converge", 7, 0, -2, 3, 3, 7, 100), ("converge", 32, 0, -2, 2, -1, -4, 120),
            ("converge", 33, 1, -2, 2, -1, -4, 60), ("converge", 64, 0, -2, 2, -1, -3, 30),
            ("steps", 32, 1, -2, 2, -1, -6, 40), ("steps", 128, 0, -2, 2, 0, -128, 6), ("steps", 12, 3, -1, 1, -3, -20, 50)]:
        env = make_sim(side=side, seed=seed, spawnStabilityFactor=spawn, stableStabilityFactor=stable, empty=empty,
                       empty_min=emin)
        steps = 0
        if mode == "converge":
            old = env.get_state()
            env.step()
            steps = 1
            count_down = limit
            while not env.match(old) and count_down:
                old = env.get_state()
                env.step()
                steps += 1
                count_down -= 1
        else:
            for _ in range(limit):
                env.step()
            steps = limit
        cases.append({"mode": mode, "side": side, "seed": seed, "spawn": spawn, "stable": stable, "empty": empty,
                      "empty_min": emin, "limit": limit, "steps": steps, "stability": int(env.stability()),
                      "alive": int(env.alive()), "world_sha": hashlib.sha256(env.world.tobytes()).hexdigest(),
                      "stable_sha": hashlib.sha256(env.stable.tobytes()).hexdigest(),
                      "breakdown": [[int(v) for v in row] for row in env.breakdown_stable()]})
        print(mode, side, seed, "steps", steps, "alive", cases[-1]["alive"], file=sys.stderr)
    with open(os.path.join(HERE, "golden_action_plus_converge.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_action_plus.py", "cases": cases}, f, indent=1)


if __name__ == "__main__":
    main()
