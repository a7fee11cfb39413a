"""GPU suite for SURVEY.md section 8 row f3: cgl_env_run (many steps per launch, shared-memory resident,
optional stop at a fixed point) and cgl_breakdown_stable, through BatchedSim and the CGL facade.

  * every case recorded from the reference (golden_converge.json) is reproduced bit for bit;
  * batches whose envs converge after different numbers of steps (divergent trip counts inside one CTA for
    side <= 64) agree with the CPU oracle env by env;
  * run(k) == k x step() on every fused side and on generic sides.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden_converge.json")) as f:
    CASES = json.load(f)["cases"]


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import native
    native.load()
    return torch.device("cuda", 0)


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"{c['mode']}-side{c['side']}-seed{c['seed']}")
def test_run_reproduces_reference_vectors(cuda, c):
    from cgl_b200.batched import BatchedSim
    env = BatchedSim(1, c["side"], seed=c["seed"], spawnStabilityFactor=c["spawn"], stableStabilityFactor=c["stable"])
    if c["mode"] == "converge":
        obs, rew, steps = env.run(c["limit"] + 1, until_fixed=True, want_alive=True)
    else:
        obs, rew, steps = env.run(c["limit"], want_alive=True)
    assert int(steps.item()) == c["steps"]
    assert int(rew.item()) == c["reward"] and int(env.last_alive().item()) == c["alive"]
    assert int(env.reward().item()) == c["reward"] and int(env.alive().item()) == c["alive"]
    assert hashlib.sha256(env.get_state().cpu().numpy().tobytes()).hexdigest() == c["world_sha"]
    assert hashlib.sha256(obs.cpu().numpy().tobytes()).hexdigest() == c["stable_sha"]
    hist = env.breakdown_stable()[0].cpu().numpy()
    vals = np.nonzero(hist)[0]
    assert [(vals - 128).tolist(), hist[vals].tolist()] == c["breakdown"]
    assert env.breakdown_state()[0].tolist() == [c["side"] ** 2 - c["alive"], c["alive"]]


@pytest.mark.parametrize("side,n_envs,limit", [(32, 37, 260), (64, 9, 70), (10, 23, 200), (96, 3, 25), (128, 5, 20),
                                               (17, 11, 120), (160, 2, 6), (256, 2, 5)])
def test_converge_batch_matches_oracle_env_by_env(cuda, side, n_envs, limit):
    from cgl_b200.batched import BatchedSim
    size = side * side
    rng = np.random.RandomState(side)
    # sparse starts die out or freeze quickly; dense ones keep oscillating: a mix of trip counts
    dens = rng.choice([0.04, 0.08, 0.15, 0.5], size=n_envs)
    cells = (rng.random_sample((n_envs, size)) < dens[:, None]).astype(np.uint8)
    cells[0] = 0                                               # empty world: fixed after one step
    if n_envs > 2:
        cells[1] = 0
        cells[1, [0, 1, side, side + 1]] = 1                   # a block: still life from the start
        cells[2] = 0
        cells[2, [1, side + 1, 2 * side + 1]] = 1              # a blinker: never fixed
    env = BatchedSim(n_envs, side, states=cells, spawnStabilityFactor=-2, stableStabilityFactor=2)
    obs, rew, steps = env.run(limit, until_fixed=True, want_alive=True)
    world_g = env.get_state().cpu().numpy()
    obs_g, rew_g, steps_g, alv_g = obs.cpu().numpy(), rew.cpu().numpy(), steps.cpu().numpy(), env.last_alive().cpu().numpy()
    seen = set()
    for e in range(n_envs):
        w = cells[e].copy()
        s = oracle.initial_stable(w, -2)
        n = oracle.run(w, s, side, -2, 2, limit, until_fixed=True)
        seen.add(n)
        assert steps_g[e] == n, (e, steps_g[e], n)
        assert np.array_equal(world_g[e], w) and np.array_equal(obs_g[e], s), e
        assert rew_g[e] == int(oracle.reward(s)) and alv_g[e] == int(oracle.alive(w))
    assert steps_g[0] == 1
    if n_envs > 2:
        assert steps_g[1] == 1 and steps_g[2] == limit
        if n_envs >= 9:
            assert len(seen) >= 3                               # the batch really had different trip counts


@pytest.mark.parametrize("k", [3, 4, 9])       # 3: the byte kernel; 4, 9: the bit-sliced kernel (cgl_env_run.cu)
@pytest.mark.parametrize("side", [32, 64, 96, 128, 160, 192, 224, 256, 5, 33, 100])
def test_run_k_equals_k_steps(cuda, side, k):
    from cgl_b200.batched import BatchedSim
    n_envs = 7 if side <= 128 else 3
    a = BatchedSim(n_envs, side, seed=side, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    b = BatchedSim(n_envs, side, seed=side, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    for _ in range(k):
        oa, ra, _ = a.step(None, want_alive=True)
    ob, rb, steps = b.run(k, want_alive=True)
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(a.world, b.world)
    assert torch.equal(a.last_alive(), b.last_alive()) and bool((steps == k).all()) and a.count == b.count == k
    # and the env keeps stepping normally afterwards (chained launches after a plain one)
    acts = torch.randint(0, side * side + 1, (n_envs,), dtype=torch.int32, device=cuda)
    oa, ra, _ = a.step(acts); ob, rb, _ = b.step(acts)
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(a.world, b.world)
    # max_steps = 0: nothing moves, reductions of the current state are returned
    w0, s0 = b.world.clone(), b.stable.clone()
    _, r0, st0 = b.run(0, want_alive=True)
    assert torch.equal(b.world, w0) and torch.equal(b.stable, s0) and bool((st0 == 0).all())
    assert torch.equal(r0, b.reward()) and torch.equal(b.last_alive(), b.alive())


def test_run_rejects_sides_that_do_not_fit(cuda):
    from cgl_b200 import native
    from cgl_b200.batched import BatchedSim
    env = BatchedSim(1, 300, rng="device")
    with pytest.raises(native.CglNativeError):
        env.run(3)
    with pytest.raises(ValueError):
        BatchedSim(1, 32, rng="device").run(-1)


def test_breakdown_full_batch(cuda):
    from cgl_b200.batched import BatchedSim
    env = BatchedSim(300, 64, seed=3, spawnStabilityFactor=-2, stableStabilityFactor=5, rng="device")
    env.run(11)
    hist = env.breakdown_stable()
    st = env.stable.to(torch.int64) + 128
    want = torch.zeros((300, 256), dtype=torch.int64, device=cuda)
    want.scatter_add_(1, st, torch.ones_like(st))
    assert torch.equal(hist, want) and int(hist.sum()) == 300 * 64 * 64
    odd = BatchedSim(5, 13, seed=1, rng="device")               # size not a multiple of the block
    odd.run(3)
    st = odd.stable.to(torch.int64) + 128
    want = torch.zeros((5, 256), dtype=torch.int64, device=cuda).scatter_add_(1, st, torch.ones_like(st))
    assert torch.equal(odd.breakdown_stable(), want)


def test_facade_run_and_breakdown(cuda):
    import CGL
    c = next(x for x in CASES if x["mode"] == "converge" and x["side"] == 10 and x["seed"] == 0)
    env = CGL.sim(side=10, seed=0, gpu=True, spawnStabilityFactor=c["spawn"], stableStabilityFactor=c["stable"])
    n = env.run(c["limit"] + 1, until_fixed=True)
    assert n == c["steps"] and env.get_count() == c["steps"]
    assert int(env.reward()) == c["reward"] and int(env.alive()) == c["alive"]
    assert env.breakdown_stable().tolist() == c["breakdown"]
    ref = oracle.OracleSim(side=10, seed=0, spawnStabilityFactor=c["spawn"], stableStabilityFactor=c["stable"])
    for _ in range(c["steps"]):
        ref.step()
    assert np.array_equal(env.get_state(vector=True), ref.world) and np.array_equal(env.get_stable(vector=True), ref.stable)
    u, cnt = np.unique(ref.world, return_counts=True)
    assert env.breakdown_state().tolist() == np.asarray((u, cnt)).tolist()
    # a pending toggle is applied before the run, like before step()
    env.toggle_state(np.int32(55)); ref.toggle_state(np.int32(55))
    assert env.run(4) == 4
    for _ in range(4):
        ref.step()
    assert np.array_equal(env.get_state(vector=True), ref.world) and np.array_equal(env.get_stable(vector=True), ref.stable)
    assert int(env.reward()) == int(ref.reward())


@pytest.mark.parametrize("impl", ["bytes", "sliced"])
def test_both_run_kernels_reproduce_reference_vectors(cuda, impl):
    """The two in-SM kernels are selected by step count; here each one is forced for every recorded case."""
    import subprocess
    import sys
    env = dict(os.environ, CGL_RUN_IMPL=impl)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-x", "-k",
                        "test_run_reproduces_reference_vectors or test_converge_batch_matches_oracle_env_by_env"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
