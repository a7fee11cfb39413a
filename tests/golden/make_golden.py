#!/usr/bin/env python
"""Generate the golden vectors in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    GPU_CAPABLE=false PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports /root/reference/CGL/CGL.py unmodified, drives ``CGL.sim(gpu=False)`` (the pure-Python
CPU step, CGL/CGL.py:211-243) and records every intermediate ``world`` / ``stable`` / ``reward`` /
``alive`` so that the oracle (oracle/) and the CUDA path can be replayed against them without
the reference tree.  Output: golden_traces.npz (step traces) + golden_api.json (API behaviour).
"""
import contextlib
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
warnings.filterwarnings("ignore", category=RuntimeWarning)  # int8 wrap in CGL.py:238 is intended


def make_sim(**kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return CGL.sim(gpu=False, **kw)


def trace(name, out, sim, actions=None, steps=None):
    """Record a run.  actions: list (len T) of None | int | list[int] applied before each step."""
    if actions is None:
        actions = [None] * steps
    T = len(actions)
    size = sim.size
    worlds = np.zeros((T + 1, size), np.uint8)
    stables = np.zeros((T + 1, size), np.int8)
    rewards = np.zeros(T + 1, np.int32)
    alives = np.zeros(T + 1, np.uint32)
    rewards_after_toggle = np.zeros(T, np.int32)
    K = max([1] + [len(a) for a in actions if isinstance(a, (list, tuple))])
    acts = np.full((T, K), -1, np.int64)            # -1 == "no toggle_state call / padding"
    worlds[0], stables[0] = sim.world, sim.stable
    rewards[0], alives[0] = sim.reward(), sim.alive()
    for t, a in enumerate(actions):
        if a is not None:
            if isinstance(a, (list, tuple)):
                acts[t, :len(a)] = a
                sim.toggle_state(list(a))
            else:
                acts[t, 0] = a
                sim.toggle_state(np.int32(a))
        rewards_after_toggle[t] = sim.reward()
        sim.step()
        worlds[t + 1], stables[t + 1] = sim.world, sim.stable
        rewards[t + 1], alives[t + 1] = sim.reward(), sim.alive()
    out[f"{name}/side"] = np.int64(sim.side)
    out[f"{name}/spawn"] = np.int64(sim.spawnStabilityFactor)
    out[f"{name}/stable_max"] = np.int64(sim.stableStabilityFactor)
    out[f"{name}/worlds"] = np.packbits(worlds, axis=1)     # worlds are {0,1}
    out[f"{name}/stables"] = stables
    out[f"{name}/rewards"] = rewards
    out[f"{name}/alives"] = alives
    out[f"{name}/rewards_after_toggle"] = rewards_after_toggle
    out[f"{name}/actions"] = acts
    assert worlds.max() <= 1
    print(f"{name:28s} side={sim.side:4d} T={T:4d} reward[-1]={rewards[-1]} alive[-1]={alives[-1]}")


def main():
    out = {}
    # (1) CGL/bench.py:26-30,44-52 -- the 5x5 blinker, spawn -128 / stable 127, 256 steps
    blinker = np.array([[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [0, 1, 1, 1, 0], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]])
    trace("blinker5", out, make_sim(state=blinker, spawnStabilityFactor=-128, stableStabilityFactor=127), steps=256)
    # (2) 4x4 torus with a centred 2x2 block, spawn -2 / stable 2 (SURVEY 8c golden 2)
    blk = np.zeros((4, 4), np.uint8); blk[1:3, 1:3] = 1
    trace("block4", out, make_sim(state=blk, spawnStabilityFactor=-2, stableStabilityFactor=2), steps=6)
    # (3) bench/main defaults: side 64 seed 0, ten plain steps (CGL/bench.py:37-40)
    trace("rand64_plain", out, make_sim(side=64, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2), steps=10)
    # (4) the DQN loop shape: toggle_state(random action incl. the no-op) then step (CGL/main.py:66-67)
    rs = np.random.RandomState(123)
    acts = [int(rs.randint(64 * 64 + 1)) for _ in range(20)]
    acts[5] = 64 * 64                                  # force one explicit "do nothing"
    trace("rand64_actions", out, make_sim(side=64, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2), actions=acts)
    # (5) tiny and ragged tori (self-neighbour multiplicity, N3) with actions
    for side in (1, 2, 3, 4, 5, 7, 10, 31, 32, 33, 40):
        size = side * side
        rs = np.random.RandomState(1000 + side)
        acts = [int(rs.randint(size + 1)) for _ in range(8)]
        trace(f"tiny{side}", out, make_sim(side=side, seed=side, spawnStabilityFactor=-3, stableStabilityFactor=4), actions=acts)
    # (6) multi-index toggles (CGL_action+/helper.py:108-132 style 2x2 blocks) with duplicates (N2)
    rs = np.random.RandomState(7)
    acts = []
    for _ in range(10):
        c = int(rs.randint(100))
        x, y = c % 10, c - c % 10
        r, d = (x + 1) % 10, (y + 10) % 100
        acts.append([x + y, r + y, x + d, r + d])
    acts[3] = [5, 5, 5, 17]                            # duplicates toggle once
    acts[6] = [99, 0, 99, 0]
    trace("multi10", out, make_sim(side=10, seed=3, spawnStabilityFactor=-2, stableStabilityFactor=2), actions=acts)
    # (7) int8 wrap: spawn above stable_max -> counter walks 5..127,-128..2 on a still life
    blk6 = np.zeros((6, 6), np.uint8); blk6[2:4, 2:4] = 1
    trace("wrap6", out, make_sim(state=blk6, spawnStabilityFactor=5, stableStabilityFactor=2), steps=300)
    # (8) pulsar 25x25 (CGL/run.py:62-86), period 3
    pulsar = np.zeros((25, 25), np.uint8)
    for r0 in (5, 10, 12, 17):
        for c0 in (9, 15):
            pulsar[r0, c0:c0 + 3] = 1
    for c0 in (7, 12, 14, 19):
        for r0 in (7, 13):
            pulsar[r0:r0 + 3, c0] = 1
    trace("pulsar25", out, make_sim(state=pulsar, spawnStabilityFactor=-2, stableStabilityFactor=2), steps=7)
    # (9) the batched-config shapes: side 128 and side 64 envs seeded s0+e, actions RandomState(seed+1e6)
    for side, n_envs, T in ((128, 2, 3), (64, 6, 5), (96, 1, 3), (200, 1, 2)):
        for e in range(n_envs):
            rs = np.random.RandomState(e + 10 ** 6)
            acts = [int(rs.randint(side * side + 1)) for _ in range(T)]
            trace(f"env{side}_{e}", out, make_sim(side=side, seed=e, spawnStabilityFactor=-2, stableStabilityFactor=2), actions=acts)
    # (10) a live->dead toggle leaves stable=spawn on a dead cell until the next step (N2)
    s = make_sim(side=6, seed=3, spawnStabilityFactor=-2, stableStabilityFactor=2)
    live = int(np.flatnonzero(s.world)[0]); dead = int(np.flatnonzero(s.world == 0)[0])
    trace("toggle6", out, s, actions=[live, dead, live, 36, [live, dead], None, dead])
    np.savez_compressed(os.path.join(HERE, "golden_traces.npz"), **out)

    # ---- API behaviour (exceptions, getters, save/load/reset/update_state/match) ----
    api = {}
    s = make_sim(side=6, seed=3)
    api["side6_seed3_world"] = s.world.tolist()
    w0 = s.world.copy()
    s.toggle_state([3, 3])
    api["toggle_dup_flips_once"] = bool(s.world[3] != w0[3]) and int(np.sum(s.world != w0)) == 1
    s.toggle_state(36)
    api["noop_changes_nothing"] = int(np.sum(s.world != w0)) == 1

    def exc(fn):
        try:
            fn()
            return None
        except Exception as e:  # noqa: BLE001
            return type(e).__name__

    api["toggle_37"] = exc(lambda: s.toggle_state(37))
    api["toggle_-1"] = exc(lambda: s.toggle_state(-1))
    api["toggle_list_invalid"] = exc(lambda: s.toggle_state([1, 99]))
    api["toggle_list_noop_single"] = exc(lambda: s.toggle_state([36]))
    api["toggle_empty_list"] = exc(lambda: s.toggle_state([]))
    api["ctor_side_str"] = exc(lambda: make_sim(side="3"))
    api["ctor_side_0"] = exc(lambda: make_sim(side=0))
    api["ctor_seed_neg"] = exc(lambda: make_sim(seed=-1))
    api["ctor_seed_float"] = exc(lambda: make_sim(seed=1.5))
    api["ctor_warp_neg"] = exc(lambda: make_sim(warp=-1))
    api["ctor_spawn_float"] = exc(lambda: make_sim(spawnStabilityFactor=1.0))
    api["ctor_stable_float"] = exc(lambda: make_sim(stableStabilityFactor=1.0))
    api["ctor_state_tuple"] = exc(lambda: make_sim(state=(1, 0)))
    api["ctor_state_empty"] = exc(lambda: make_sim(state=np.zeros(0)))
    api["ctor_gpu_int"] = exc(lambda: make_sim(gpu_select="x"))   # never type-checked (N5)
    api["ctor_gpu_mismatch"] = exc(lambda: CGL.sim(gpu=True))      # GPU_CAPABLE=false vs gpu=True
    api["ctor_spawn_overflow"] = exc(lambda: make_sim(spawnStabilityFactor=-200))
    s = make_sim(side=5, seed=1, spawnStabilityFactor=-2, stableStabilityFactor=2)
    api["getters"] = dict(side=s.get_side(), count=s.get_count(), seed=s.get_seed(), state_dim=s.get_state_dim(),
                          state_space_dim=str(s.get_state_space_dim()), action_space_dim=s.get_action_space_dim())
    s.step(); s.step()
    api["count_after_2"] = s.get_count()
    s.reset()
    api["count_after_reset"] = s.get_count()           # reset() does not touch count
    api["reset_restores"] = bool((s.world == s.initState).all() and (s.stable == s.initStable).all())
    api["update_state_bad_size"] = exc(lambda: s.update_state(np.zeros(9), 3))
    api["update_state_bad_side"] = exc(lambda: s.update_state(np.zeros(25), 4))
    new = np.arange(25) % 2
    st_before = s.stable.copy()
    s.update_state(new, 5)
    api["update_state_keeps_stable"] = bool((s.stable == st_before).all() and (s.world == new).all())
    api["match_true"] = bool(s.match(new.reshape(5, 5)))
    api["match_false"] = bool(s.match(1 - new))
    sv = s.save()
    api["save_len"] = len(sv)
    api["save_types"] = [type(v).__name__ for v in sv]
    api["load_bad_type"] = exc(lambda: s.load([1], np.zeros(1, np.int8), 1, 0, -1, 1))
    api["load_bad_side"] = exc(lambda: s.load(np.zeros(1), np.zeros(1), 0, 0, -1, 1))
    api["load_bad_count"] = exc(lambda: s.load(np.zeros(1), np.zeros(1), 1, -1, -1, 1))
    api["load_float_spawn"] = exc(lambda: s.load(np.zeros(1), np.zeros(1), 1, 0, -1.0, 1))
    # shallow views alias the live state (N1)
    s = make_sim(side=4, seed=2)
    v = s.get_stable(vector=True, shallow=True)
    s.step()
    api["shallow_stable_is_live"] = bool(v is s.get_stable(vector=True, shallow=True) and (v == s.stable).all())
    api["get_state_shape"] = list(s.get_state().shape)
    api["get_stable_dtype"] = str(s.get_stable().dtype)
    api["get_state_dtype"] = str(s.get_state().dtype)
    api["reward_type"] = type(s.reward()).__name__
    api["alive_type"] = type(s.alive()).__name__
    with open(os.path.join(HERE, "golden_api.json"), "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
    print(json.dumps({k: v for k, v in api.items() if k != "side6_seed3_world"}, indent=1))


if __name__ == "__main__":
    main()
