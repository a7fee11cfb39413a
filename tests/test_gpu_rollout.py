"""GPU suite: the host-driven, double-buffered rollout (cgl_rollout_*, cgl_b200/rollout.py) and the C launch loop
(cgl_env_step_seq, StepSequence) against the CPU oracle.  Reference loop: CGL/main.py:64-72."""
import numpy as np
import pytest

from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SPAWN, STABLE = -2, 2


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import batched, rollout
    return batched, rollout


@pytest.mark.parametrize("side,n_envs,groups,replicas", [(64, 24, 2, 1), (128, 12, 2, 3), (32, 30, 3, 2), (10, 8, 2, 1),
                                                         (96, 7, 1, 1)])
def test_rollout_with_reward_dependent_policy_matches_oracle(mods, side, n_envs, groups, replicas):
    _, rollout = mods
    size = side * side
    ro = rollout.HostRollout(n_envs, side, n_groups=groups, n_replicas=replicas, seed=5, spawnStabilityFactor=SPAWN,
                             stableStabilityFactor=STABLE, obs_to_host=True)
    per = n_envs // groups
    # oracle twin: replica r, group g
    cells = [[ro.sim(g, r).get_state().cpu().numpy() for r in range(replicas)] for g in range(groups)]
    stab = [[ro.sim(g, r).stable.cpu().numpy().copy() for r in range(replicas)] for g in range(groups)]
    last_rew = [np.zeros(per, np.int32) for _ in range(groups)]
    log = []

    def pick(rew, step, group):                             # an action that depends on the group's last rewards
        return ((rew.astype(np.int64) * 7 + step * 13 + group * 5 + np.arange(per) * 3) % (size + 1)).astype(np.int32)

    def policy(group, step, rewards, actions):
        log.append((group, step, rewards.copy()))
        actions[:] = pick(rewards, step, group)

    steps = 7
    ro.run(steps, policy)
    ro.run(2, policy)                                       # a second call continues the rotation
    for s in range(steps + 2):
        r = s % replicas
        for g in range(groups):
            acts = pick(last_rew[g], s, g)
            rew, _ = oracle.step_batch(cells[g][r], stab[g][r], side, acts, SPAWN, STABLE, threads=2)
            last_rew[g] = np.asarray(rew, np.int32)
    # the policy saw exactly the oracle's reward stream
    assert len(log) == (steps + 2) * groups
    for g in range(groups):
        assert np.array_equal(ro.rewards[g], last_rew[g])
        for r in range(replicas):
            sim = ro.sim(g, r)
            assert np.array_equal(sim.get_state().cpu().numpy(), cells[g][r]), (g, r)
            assert np.array_equal(sim.stable.cpu().numpy(), stab[g][r]), (g, r)
        assert np.array_equal(ro.obs[g], stab[g][(steps + 1) % replicas])
    ro.close()


@pytest.mark.parametrize("side,n_envs", [(128, 9), (64, 40), (10, 6)])
@pytest.mark.parametrize("chained", ["0", "1"])
def test_step_sequence_matches_single_steps(mods, side, n_envs, chained, monkeypatch):
    batched, _ = mods
    monkeypatch.setenv("CGL_ENV_CHAINED", chained)
    size = side * side
    R, A = 3, 4
    sims = [batched.BatchedSim(n_envs, side, seed=r * 100, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE) for r in range(R)]
    cells = [s.get_state().cpu().numpy() for s in sims]
    stab = [s.stable.cpu().numpy().copy() for s in sims]
    acts_h = np.random.RandomState(side).randint(size + 1, size=(A, n_envs)).astype(np.int32)
    seq = batched.StepSequence(sims, torch.from_numpy(acts_h).cuda())
    done = 0
    for k in (5, 24, 1, 7):
        seq.run(k)
        for i in range(k):
            oracle.step_batch(cells[i % R], stab[i % R], side, acts_h[i % A], SPAWN, STABLE, threads=2)
        done += k
        # single steps interleave freely with sequences
        sims[0].step(torch.from_numpy(acts_h[0]).cuda())
        oracle.step_batch(cells[0], stab[0], side, acts_h[0], SPAWN, STABLE, threads=2)
        for r in range(R):
            assert np.array_equal(sims[r].get_state().cpu().numpy(), cells[r]), (k, r)
            assert np.array_equal(sims[r].stable.cpu().numpy(), stab[r]), (k, r)
            assert np.array_equal(sims[r]._reward.cpu().numpy(), stab[r].astype(np.int32).sum(axis=1))
    for s in sims:
        s.check_actions()


def test_sequence_numbers_then_graph_capture_then_eager(mods, monkeypatch):
    """A batch steps with sequence-number tokens, is then captured in a CUDA graph (switches to plane ids for good),
    replayed, and stepped eagerly again: every transition re-initialises the tokens."""
    batched, _ = mods
    monkeypatch.setenv("CGL_ENV_CHAINED", "1")
    side, n = 128, 5
    env = batched.BatchedSim(n, side, seed=1, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
    cells, st = env.get_state().cpu().numpy(), env.stable.cpu().numpy().copy()

    def ref(k):
        for _ in range(k):
            oracle.step_batch(cells, st, side, None, SPAWN, STABLE, threads=2)

    for _ in range(5):
        env.step()
    ref(5)
    assert not env._chain_ids
    env.run(3); ref(3)
    env.step(); ref(1)
    torch.cuda.synchronize()
    stream, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(4):
                env.step()
    assert env._chain_ids
    graph.replay(); graph.replay()
    ref(8)
    for _ in range(3):
        env.step()
    ref(3)
    torch.cuda.synchronize()
    assert np.array_equal(env.get_state().cpu().numpy(), cells) and np.array_equal(env.stable.cpu().numpy(), st)
    env.check_actions()
