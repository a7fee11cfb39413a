"""GPU suite for SURVEY.md section 8 row f1: the out-of-place env step (cgl_env_step_io) and the batched
DQN loop built on it (cgl_b200/dqn.py).

  * stepping through a ring of observation slots gives bit-identical planes, rewards and worlds to the
    in-place step, leaves the source slot untouched, on fused (chained and not) and generic sides;
  * the replay ring filled by the CUDA env equals the ring filled by the CPU oracle env for the same
    actions, and what sample() returns are exactly recorded transitions;
  * the whole loop (select_action -> step -> learn -> target_update) runs on the device.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import native
    native.load()
    return torch.device("cuda", 0)


@pytest.mark.parametrize("side,n_envs", [(128, 300), (64, 515), (32, 130), (96, 40), (256, 9), (10, 33), (33, 17), (130, 5)])
def test_out_of_place_step_equals_in_place(cuda, side, n_envs):
    from cgl_b200.batched import BatchedSim
    size = side * side
    a = BatchedSim(n_envs, side, seed=2, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    b = BatchedSim(n_envs, side, seed=2, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    ring = torch.full((3, n_envs, size), 77, dtype=torch.int8, device=cuda)
    rew_ring = torch.zeros((3, n_envs), dtype=torch.int32, device=cuda)
    b.bind_observation(ring[0])
    assert b.stable.data_ptr() == ring[0].data_ptr() and torch.equal(a.stable, b.stable)
    g = torch.Generator(device=cuda); g.manual_seed(side)
    for t in range(7):
        acts = None if t == 3 else torch.randint(0, size + 1, (n_envs,), dtype=torch.int32, device=cuda, generator=g)
        before = ring[t % 3].clone()
        oa, ra, _ = a.step(acts, want_alive=True)
        ob, rb, _ = b.step(acts, want_alive=True, obs_out=ring[(t + 1) % 3], reward_out=rew_ring[t % 3])
        assert ob.data_ptr() == ring[(t + 1) % 3].data_ptr() and rb.data_ptr() == rew_ring[t % 3].data_ptr()
        assert torch.equal(oa, ob) and torch.equal(ra, rb), (side, t)
        assert torch.equal(a.world, b.world) and torch.equal(a.last_alive(), b.last_alive())
        if b.fused:
            assert torch.equal(ring[t % 3], before)           # the state slot is read-only for the fused kernel
    # mixing in-place and out-of-place steps on one env keeps working (plane tokens, cached arguments)
    oa, ra, _ = a.step(None)
    ob, rb, _ = b.step(None)
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and ob.data_ptr() == ring[(7) % 3].data_ptr()
    a.check_actions(); b.check_actions()


def test_out_of_place_step_rejects_bad_buffers(cuda):
    from cgl_b200.batched import BatchedSim
    env = BatchedSim(4, 32, rng="device")
    with pytest.raises(TypeError):
        env.step(None, obs_out=torch.zeros((4, 1024), dtype=torch.uint8, device=cuda))
    with pytest.raises(TypeError):
        env.step(None, obs_out=torch.zeros((4, 1000), dtype=torch.int8, device=cuda))
    with pytest.raises(TypeError):
        env.step(None, reward_out=torch.zeros(4, dtype=torch.int64, device=cuda))
    with pytest.raises(TypeError):
        env.step(None, obs_out=torch.zeros((4, 1024), dtype=torch.int8))


@pytest.mark.parametrize("side,B", [(64, 6), (10, 5), (128, 3)])
def test_replay_ring_filled_by_cuda_env_equals_oracle_env(cuda, side, B):
    from cgl_b200.batched import BatchedSim
    from cgl_b200.dqn import TrajectoryReplay
    from oracle_env import OracleBatchEnv
    size = side * side
    env = BatchedSim(B, side, seed=5, spawnStabilityFactor=-2, stableStabilityFactor=2)       # reference RNG start
    ref = OracleBatchEnv(B, side, seed=5, spawn=-2, stable_max=2)
    mem = TrajectoryReplay(env, max_size=4 * B, batch_size=32, seed=0)
    rmem = TrajectoryReplay(ref, max_size=4 * B, batch_size=32, seed=0)
    rng = np.random.RandomState(1)
    mem.reset(); rmem.reset()
    for t in range(14):
        if t == 8:
            mem.reset(); rmem.reset()
        acts = rng.randint(0, size + 1, size=B).astype(np.int32)
        n, r = mem.step(torch.from_numpy(acts).to(cuda))
        rn, rr = rmem.step(torch.from_numpy(acts))
        assert torch.equal(n.cpu(), rn) and torch.equal(r.cpu(), rr), t
        assert mem.valid_steps() == rmem.valid_steps() and mem.size == rmem.size and mem.t == rmem.t
    valid = [k % mem.slots for k in mem.valid_steps()]
    assert torch.equal(mem.obs.cpu(), rmem.obs)
    assert torch.equal(mem.action.cpu()[valid], rmem.action[valid])
    assert torch.equal(mem.reward.cpu()[valid], rmem.reward[valid])
    # samples are recorded transitions: state/next are consecutive slots of the same env
    slot, e = mem.sample_indices()
    assert set(int(s) for s in slot.cpu()) <= set(valid)
    s, a, r, n = mem.gather(slot, e)
    rs, ra, rr, rn = rmem.gather(slot.cpu(), e.cpu())
    assert torch.equal(s.cpu(), rs) and torch.equal(a.cpu(), ra) and torch.equal(r.cpu(), rr) and torch.equal(n.cpu(), rn)
    env.check_actions()


def test_batched_dqn_loop_on_device(cuda):
    from cgl_b200.batched import BatchedSim
    from cgl_b200.dqn import BatchedDQNAgent
    B, side = 256, 32
    env = BatchedSim(B, side, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    twin = BatchedSim(B, side, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    agent = BatchedDQNAgent(env, max_size=8 * B, batch_size=64, seed=1, hidden=256)
    assert agent.state_dim == 1024 and agent.action_dim == 1025
    w0 = agent.Q.l1.weight.clone()
    tgt0 = agent.Q_target.l1.weight.clone()
    losses = []
    real_learn = agent.learn
    agent.learn = lambda ex, d: losses.append(real_learn(ex, d))
    eps = 1.0
    for episode in range(2):
        state = agent.reset()
        twin.reset()
        total = torch.zeros(B, dtype=torch.int64, device=cuda)
        for _ in range(12):
            action = agent.select_action(state, eps, out=agent.memory.action_slot())
            state, reward = agent.step(action)
            # the env inside the loop is the same env: an in-place twin fed the same actions agrees
            to, tr, _ = twin.step(action)
            assert torch.equal(to, state) and torch.equal(tr, reward)
            total += reward
        eps *= 0.5
    env.check_actions()
    assert len(losses) == 24 and all(torch.isfinite(l) for l in losses)
    assert not torch.equal(agent.Q.l1.weight, w0) and not torch.equal(agent.Q_target.l1.weight, tgt0)
    assert agent.memory.size == 8 * B          # ring of 9 slots: 8 complete transitions per env


def test_two_gpu_data_parallel_loop(cuda):
    import os
    import subprocess
    import sys
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "run_dqn.py"), "--side", "32", "--envs", "512",
           "--steps", "12", "--hidden", "256", "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"replicas_identical": true' in r.stdout
