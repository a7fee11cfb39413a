"""GPU suite for SURVEY.md section 8 row f2: the CGL_action+ fork's env (dead-cell rules, masked toggle,
`empty` start) through BatchedSim / cgl_env_step_rule and through the fork facade.

  * every trace recorded from the fork's own CPU back end is replayed bit for bit (dead_rule="sat");
  * random batches with actions agree with the oracle for both dead-cell rules, on fused (in place, into a
    ring, chained) and generic sides;
  * the rule entry point with dead_rule="zero" and unmasked toggles is the base env.
"""
import importlib.util
import os

import numpy as np
import pytest

from conftest import FORK_TRACES, PKG
from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RULE_ID = {"zero": oracle.DEAD_ZERO, "decay": oracle.DEAD_DECAY, "sat": oracle.DEAD_SAT}


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import native
    native.load()
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def fork(cuda):
    spec = importlib.util.spec_from_file_location("cgl_fork_facade", os.path.join(PKG, "CGL_action+", "CGL.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _env(tr, **kw):
    from cgl_b200.batched import BatchedSim
    return BatchedSim(1, tr.side, states=tr.worlds[0][None, :], spawnStabilityFactor=tr.spawn,
                      stableStabilityFactor=tr.stable_max, dead_rule="sat", empty=tr.empty, empty_min=tr.empty_min,
                      masked_toggle=True, **kw)


@pytest.mark.parametrize("fused_toggle", [True, False])
@pytest.mark.parametrize("name", sorted(FORK_TRACES))
def test_batched_replays_fork_trace(cuda, name, fused_toggle):
    tr = FORK_TRACES[name]
    env = _env(tr)
    assert np.array_equal(env.stable.cpu().numpy()[0], tr.stables[0])
    for t in range(tr.T):
        a = tr.action(t)
        acts = None
        if a is not None:
            if isinstance(a, int) and fused_toggle:
                acts = torch.tensor([a], dtype=torch.int32, device=cuda)           # toggled inside the step kernel
            else:
                idx = [a] if isinstance(a, int) else a
                env.toggle(torch.tensor([idx], dtype=torch.int32, device=cuda))
                assert np.array_equal(env.stable.cpu().numpy()[0], tr.toggled_stables[t]), (name, t)
        obs, rew, _ = env.step(acts, want_alive=True)
        assert np.array_equal(env.get_state().cpu().numpy()[0], tr.worlds[t + 1]), (name, t)
        assert np.array_equal(obs.cpu().numpy()[0], tr.stables[t + 1]), (name, t)
        assert int(rew.item()) == tr.stability[t + 1] and int(env.last_alive().item()) == tr.alives[t + 1]
    env.check_actions()


@pytest.mark.parametrize("rule", ["decay", "sat", "zero"])
@pytest.mark.parametrize("side,n_envs", [(128, 70), (64, 130), (32, 260), (96, 9), (256, 3), (10, 40), (33, 12), (130, 3)])
def test_random_batches_vs_oracle(cuda, rule, side, n_envs):
    from cgl_b200.batched import BatchedSim
    size = side * side
    spawn, smax, empty, emin = -2, 3, -1, -7
    rng = np.random.RandomState(side + len(rule))
    cells = rng.randint(2, size=(n_envs, size)).astype(np.uint8)
    env = BatchedSim(n_envs, side, states=cells, spawnStabilityFactor=spawn, stableStabilityFactor=smax,
                     dead_rule=rule, empty=empty, empty_min=emin, masked_toggle=True)
    ring = torch.empty((3, n_envs, size), dtype=torch.int8, device=cuda)
    world = cells.copy()
    stable = np.stack([oracle.initial_stable_fork(w, spawn, empty) for w in world])
    assert np.array_equal(env.stable.cpu().numpy(), stable)
    sample = range(n_envs) if n_envs <= 40 else rng.choice(n_envs, 40, replace=False)
    for t in range(6):
        acts = rng.randint(0, size + 1, size=n_envs).astype(np.int32)
        if t == 2:
            acts[:] = size                                        # all no-ops
        a_dev = None if t == 4 else torch.from_numpy(acts).to(cuda)
        if t % 2:                                                 # alternate in-place and into-a-ring steps
            obs, rew, _ = env.step(a_dev, want_alive=True, obs_out=ring[t % 3])
        else:
            obs, rew, _ = env.step(a_dev, want_alive=True)
        obs_g, rew_g, w_g = obs.cpu().numpy(), rew.cpu().numpy(), env.get_state().cpu().numpy()
        for e in sample:
            if t != 4:
                oracle.toggle_masked(world[e], stable[e], int(acts[e]), spawn)
            oracle.step_rule(world[e], stable[e], side, spawn, smax, RULE_ID[rule], empty, emin)
            assert np.array_equal(w_g[e], world[e]) and np.array_equal(obs_g[e], stable[e]), (rule, side, t, e)
            assert rew_g[e] == int(oracle.reward(stable[e]))
    env.check_actions()


@pytest.mark.parametrize("side", [128, 64, 20])
def test_rule_entry_with_base_settings_is_the_base_env(cuda, side):
    from cgl_b200 import native
    from cgl_b200.batched import BatchedSim
    lib = native.load()
    n = 50
    a = BatchedSim(n, side, seed=1, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    b = BatchedSim(n, side, seed=1, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    g = torch.Generator(device=cuda); g.manual_seed(3)
    for _ in range(4):
        acts = torch.randint(0, side * side + 1, (n,), dtype=torch.int32, device=cuda, generator=g)
        oa, ra, _ = a.step(acts)
        rew = torch.zeros(n, dtype=torch.int32, device=cuda)
        native.check(lib.cgl_env_step_rule(native.dptr(b._wa), native.dptr(b._wb), native.dptr(b.stable),
                                           native.dptr(b.stable), n, side, native.dptr(acts), -2, 2, 0, 0, -128, 0,
                                           native.dptr(rew), None, native.dptr(b._err), None, 0, 0,
                                           native.current_stream()))
        b._wa, b._wb = b._wb, b._wa
        assert torch.equal(oa, b.stable) and torch.equal(ra, rew) and torch.equal(a.world, b.world)
    with pytest.raises(native.CglNativeError):
        native.check(lib.cgl_env_step_rule(native.dptr(b._wa), native.dptr(b._wb), native.dptr(b.stable),
                                           native.dptr(b.stable), n, side, None, -2, 2, 3, 0, -128, 0, None, None,
                                           None, None, 0, 0, native.current_stream()))


def test_block_action_matches_the_forks_helper_formula(cuda):
    from cgl_b200.batched import BatchedSim
    side = 12
    env = BatchedSim(6, side, rng="device", masked_toggle=True)
    centers = torch.tensor([0, 11, 143, 132, 144, 77], dtype=torch.int32, device=cuda)
    got = env.block_action(centers).cpu().numpy()
    for e, c in enumerate(centers.cpu().numpy()):
        if c < side * side:                                       # CGL_action+/helper.py:113-127
            x = c % side; y = c - x
            right = (x + 1) % side; down = (y + side) % (side * side)
            want = [x + y, right + y, x + down, right + down]
        else:
            want = [side * side] * 4
        assert got[e].tolist() == want
    # the recorded block trace used exactly such index lists
    tr = FORK_TRACES["blank12_blocks"]
    blocks = [tr.action(t) for t in range(tr.T) if isinstance(tr.action(t), list) and len(tr.action(t)) == 4]
    anchors = torch.tensor([b[0] for b in blocks[:6]], dtype=torch.int32, device=cuda)
    assert env.block_action(anchors).cpu().numpy().tolist() == blocks[:6]


@pytest.mark.parametrize("name", ["floor64", "blank12_blocks", "tiny7", "posempty10", "fused128"])
def test_fork_facade_replays_fork_trace(fork, name):
    tr = FORK_TRACES[name]
    kw = dict(gpu=True, spawnStabilityFactor=tr.spawn, stableStabilityFactor=tr.stable_max, empty=tr.empty,
              empty_min=tr.empty_min, dead_rule="sat")
    env = fork.sim(state=tr.worlds[0].copy(), **kw)
    assert env.get_max_density() == tr.max_density and env.max_density == tr.max_density
    obs = env.get_stable(vector=True, shallow=True)
    assert np.array_equal(obs, tr.stables[0]) and int(env.stability()) == tr.stability[0]
    for t in range(tr.T):
        a = tr.action(t)
        if a is not None:
            env.toggle_state(np.int32(a) if isinstance(a, int) else a)
        env.step()
        assert np.array_equal(env.get_state(vector=True), tr.worlds[t + 1]), (name, t)
        assert np.array_equal(env.get_stable(vector=True, shallow=True), tr.stables[t + 1]), (name, t)
        assert env.stability() == tr.stability[t + 1] and env.alive() == tr.alives[t + 1]
        assert isinstance(env.stability(), np.int32) and isinstance(env.alive(), np.int32)
    assert env.get_count() == tr.T


def test_fork_facade_api(fork):
    env = fork.sim(side=10, seed=3, gpu=True, spawnStabilityFactor=-2, stableStabilityFactor=2, empty=-1, empty_min=-4)
    assert env.dead_rule == "decay"                              # gpu=True is the fork's CUDA kernel
    w0 = oracle.initial_world(10, 3)
    s0 = oracle.initial_stable_fork(w0, -2, -1)
    assert np.array_equal(env.get_state(vector=True), w0) and np.array_equal(env.get_stable(vector=True), s0)
    w, s = w0.copy(), s0.copy()
    for a in (5, 100, [1, 2, 11, 12], 37):
        env.toggle_state(a if isinstance(a, list) else np.int32(a))
        oracle.toggle_masked(w, s, a, -2)
        env.step()
        oracle.step_rule(w, s, 10, -2, 2, oracle.DEAD_DECAY, -1, -4)
        assert np.array_equal(env.get_state(vector=True), w) and np.array_equal(env.get_stable(vector=True), s)
        assert int(env.stability()) == int(oracle.reward(s))
    assert env.breakdown_stable().tolist() == oracle.breakdown(s).tolist()
    # fresh(seed): new random world, stability re-initialised with `empty`
    env.fresh(seed=9)
    np.random.seed(9)
    wf = np.random.randint(2, size=100, dtype=np.uint8)
    assert np.array_equal(env.get_state(vector=True), wf)
    assert np.array_equal(env.get_stable(vector=True), oracle.initial_stable_fork(wf, -2, -1))
    env.reset()
    assert np.array_equal(env.get_state(vector=True), w0) and np.array_equal(env.get_stable(vector=True), s0)
    # update_state with a stability plane; save/load carry `empty`
    new_s = np.arange(100, dtype=np.int8) - 50
    env.update_state(1 - w0, new_s)
    assert np.array_equal(env.get_state(vector=True), 1 - w0) and np.array_equal(env.get_stable(vector=True), new_s)
    sv = env.save()
    assert len(sv) == 7 and sv[6] == -1
    other = fork.sim(side=10, seed=0, gpu=True, empty=0, empty_min=-4)
    other.load(*sv)
    assert other.empty == -1 and np.array_equal(other.get_stable(vector=True), new_s)
    other.fresh(seed=1)
    np.random.seed(1)
    wf = np.random.randint(2, size=100, dtype=np.uint8)
    assert np.array_equal(other.get_stable(vector=True), oracle.initial_stable_fork(wf, sv[4], -1))
    # runBlank, validation
    blank = fork.sim(side=6, seed=0, gpu=True, runBlank=True, empty=-3)
    assert int(blank.alive()) == 0 and int(blank.stability()) == -3 * 36
    assert fork.sim(side=70, seed=0, gpu=True).get_max_density() == np.floor(70 * 70 / 2 + 17 / 27 * 70 - 2)   # 70 % 54 = 16
    assert fork.sim(side=64, seed=0, gpu=True).get_max_density() == np.floor(64 * 64 / 2 + 17 / 27 * 64 - 1)
    for bad, exc in ((dict(runBlank=1), TypeError), (dict(empty=1.0), TypeError), (dict(empty_min="x"), TypeError),
                     (dict(empty=300), OverflowError), (dict(dead_rule="zero"), ValueError)):
        with pytest.raises(exc):
            fork.sim(side=4, gpu=True, **bad)


import hashlib  # noqa: E402
import json  # noqa: E402

from conftest import ROOT  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "golden_action_plus_converge.json")) as _f:
    FORK_RUNS = json.load(_f)["cases"]


@pytest.mark.parametrize("c", FORK_RUNS, ids=lambda c: f"{c['mode']}-side{c['side']}-seed{c['seed']}")
def test_run_on_the_fork_env_reproduces_the_forks_loops(cuda, c):
    from cgl_b200.batched import BatchedSim
    env = BatchedSim(1, c["side"], seed=c["seed"], spawnStabilityFactor=c["spawn"], stableStabilityFactor=c["stable"],
                     dead_rule="sat", empty=c["empty"], empty_min=c["empty_min"], masked_toggle=True)
    conv = c["mode"] == "converge"
    obs, rew, steps = env.run(c["limit"] + conv, until_fixed=conv, want_alive=True)
    assert int(steps.item()) == c["steps"] and int(rew.item()) == c["stability"]
    assert int(env.last_alive().item()) == c["alive"]
    assert hashlib.sha256(env.get_state().cpu().numpy().tobytes()).hexdigest() == c["world_sha"]
    assert hashlib.sha256(obs.cpu().numpy().tobytes()).hexdigest() == c["stable_sha"]
    hist = env.breakdown_stable()[0].cpu().numpy()
    vals = np.nonzero(hist)[0]
    assert [(vals - 128).tolist(), hist[vals].tolist()] == c["breakdown"]


@pytest.mark.parametrize("rule", ["decay", "sat"])
@pytest.mark.parametrize("side,k", [(128, 5), (64, 12), (32, 7), (96, 4), (20, 9)])
def test_fork_run_k_equals_k_steps(cuda, rule, side, k):
    from cgl_b200.batched import BatchedSim
    kw = dict(seed=side, spawnStabilityFactor=-2, stableStabilityFactor=3, rng="device", dead_rule=rule, empty=-1,
              empty_min=-7, masked_toggle=True)
    a, b = BatchedSim(9, side, **kw), BatchedSim(9, side, **kw)
    for _ in range(k):
        oa, ra, _ = a.step(None, want_alive=True)
    ob, rb, steps = b.run(k, want_alive=True)
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(a.world, b.world)
    assert torch.equal(a.last_alive(), b.last_alive()) and bool((steps == k).all())


def test_fork_facade_run(fork):
    c = next(x for x in FORK_RUNS if x["mode"] == "converge" and x["side"] == 10 and x["seed"] == 1)
    env = fork.sim(side=10, seed=1, gpu=True, spawnStabilityFactor=c["spawn"], stableStabilityFactor=c["stable"],
                   empty=c["empty"], empty_min=c["empty_min"], dead_rule="sat")
    assert env.run(c["limit"] + 1, until_fixed=True) == c["steps"]
    assert int(env.stability()) == c["stability"] and int(env.alive()) == c["alive"]
    assert env.breakdown_stable().tolist() == c["breakdown"]


@pytest.mark.parametrize("side,n_envs,limit", [(32, 20, 90), (64, 9, 40), (128, 4, 15), (96, 3, 12), (10, 12, 80)])
def test_fork_decay_run_until_fixed_vs_oracle(cuda, side, n_envs, limit):
    """The fork's default (CUDA-kernel) rule through the on-chip run kernels (bit-sliced for fused sides), with
    envs that stop after different numbers of steps."""
    from cgl_b200.batched import BatchedSim
    size = side * side
    rng = np.random.RandomState(side)
    dens = rng.choice([0.03, 0.08, 0.2, 0.5], size=n_envs)
    cells = (rng.random_sample((n_envs, size)) < dens[:, None]).astype(np.uint8)
    cells[0] = 0
    spawn, smax, empty, emin = -2, 3, -1, -9
    env = BatchedSim(n_envs, side, states=cells, spawnStabilityFactor=spawn, stableStabilityFactor=smax,
                     dead_rule="decay", empty=empty, empty_min=emin, masked_toggle=True)
    obs, rew, steps = env.run(limit, until_fixed=True, want_alive=True)
    w_g, o_g, r_g, s_g = env.get_state().cpu().numpy(), obs.cpu().numpy(), rew.cpu().numpy(), steps.cpu().numpy()
    for e in range(n_envs):
        w = cells[e].copy()
        s = oracle.initial_stable_fork(w, spawn, empty)
        n = oracle.run_rule(w, s, side, spawn, smax, limit, oracle.DEAD_DECAY, empty, emin, until_fixed=True)
        assert s_g[e] == n, (e, s_g[e], n)
        assert np.array_equal(w_g[e], w) and np.array_equal(o_g[e], s), e
        assert r_g[e] == int(oracle.reward(s))
    assert s_g[0] == 1
