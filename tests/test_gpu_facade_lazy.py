"""GPU suite: the facade's deferred plain steps (CGL.sim.step batches runs of plain steps into one on-chip
launch).  Whatever the interleaving of calls, every observation must equal the CPU oracle's -- with and without
live shallow views (which switch the batching off)."""
import numpy as np
import pytest

from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def CGL():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import CGL as mod
    return mod


def test_reference_bench_loop_is_one_launch(CGL):
    """CGL/bench.py:37-52 shape: construct, `iters` plain steps, print Stability and Life."""
    env = CGL.sim(side=64, seed=0, gpu=True, spawnStabilityFactor=-2, stableStabilityFactor=2)
    ref = oracle.OracleSim(side=64, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2)
    before = env._b.launches
    for _ in range(50):
        env.step()
        ref.step()
    assert env._b.launches == before                        # nothing has run yet
    assert int(env.reward()) == int(ref.reward()) and int(env.alive()) == int(ref.alive())
    assert env._b.launches - before <= 3                    # one run launch (+ the reductions asked for)
    assert env.get_count() == 50
    assert np.array_equal(env.get_state(vector=True), ref.world) and np.array_equal(env.get_stable(vector=True), ref.stable)


@pytest.mark.parametrize("side", [10, 32, 64, 33, 128])
@pytest.mark.parametrize("live_views", [False, True])
def test_random_call_interleavings_match_oracle(CGL, side, live_views):
    size = side * side
    env = CGL.sim(side=side, seed=7, gpu=True, spawnStabilityFactor=-2, stableStabilityFactor=3)
    ref = oracle.OracleSim(side=side, seed=7, spawnStabilityFactor=-2, stableStabilityFactor=3)
    views = (env.get_state(vector=True, shallow=True), env.get_stable(vector=True, shallow=True)) if live_views else None
    rs = np.random.RandomState(side + live_views)
    for it in range(160):
        op = rs.choice(["step", "step", "step", "step", "toggle", "toggle_list", "reward", "alive", "state", "stable",
                        "reset", "match", "update", "saveload", "run"])
        if op == "step":
            env.step(); ref.step()
        elif op == "toggle":
            a = np.int32(rs.randint(size + 1))
            env.toggle_state(a); ref.toggle_state(a)
        elif op == "toggle_list":
            a = [int(v) for v in rs.randint(size, size=3)]
            env.toggle_state(a); ref.toggle_state(a)
        elif op == "reward":
            assert int(env.reward()) == int(ref.reward()), it
        elif op == "alive":
            assert int(env.alive()) == int(ref.alive()), it
        elif op == "state":
            assert np.array_equal(env.get_state(vector=True), ref.world), it
        elif op == "stable":
            assert np.array_equal(env.get_stable(), ref.stable.reshape(side, side)), it
        elif op == "reset":
            env.reset(); ref.reset()
        elif op == "match":
            assert env.match(ref.world.reshape(side, side)), it
        elif op == "update":
            new = rs.randint(2, size=size).astype(np.uint8)
            env.update_state(new, side)
            ref.world = new.copy()
        elif op == "saveload":
            sv = env.save()
            assert np.array_equal(sv[0], ref.world) and np.array_equal(sv[1], ref.stable) and sv[3] == env.get_count()
            env.load(*sv)
        elif op == "run":
            n = int(rs.randint(1, 7))
            assert env.run(n) == n
            for _ in range(n):
                ref.step()
        if views is not None and op != "toggle":             # live views show the current state (a scalar toggle
            # is documented to appear with the next call into the env)
            assert np.array_equal(views[0], ref.world) and np.array_equal(views[1], ref.stable), (it, op)
    assert np.array_equal(env.get_state(vector=True), ref.world) and np.array_equal(env.get_stable(vector=True), ref.stable)
    assert int(env.reward()) == int(ref.reward())
