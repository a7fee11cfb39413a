"""GPU suite: the drop-in facade UNDER the reference's DQN loop shape (CGL/main.py:58-75).

A fixed random MLP on the GPU plays the agent (select_action = argmax Q(obs), CGL/dqn.py:127-134 with
eps = 0): obs -> action -> toggle_state -> step -> get_stable(shallow) -> reward.  Because the action is a
function of the observation, one wrong byte anywhere makes the trajectories diverge; the CPU oracle is
driven through the same loop and every observation / reward / world must match."""
import numpy as np
import pytest

from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("side,steps", [(10, 300), (64, 120), (33, 100)])
def test_facade_under_dqn_loop_matches_oracle(side, steps):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import CGL
    size = side * side
    torch.manual_seed(side)
    qnet = torch.nn.Sequential(torch.nn.Linear(size, 2 * (size + 1)), torch.nn.ReLU(),
                               torch.nn.Linear(2 * (size + 1), size + 1)).cuda()       # CGL/dqn.py:52-60 shape

    def select_action(obs_np):                      # CGL/dqn.py:127-134 with eps = 0
        with torch.no_grad():
            q = qnet(torch.tensor(obs_np, dtype=torch.int8).to(torch.float32).cuda())
        return np.int32(int(torch.argmax(q).item()))

    env = CGL.sim(side=side, seed=4, gpu=True, spawnStabilityFactor=-2, stableStabilityFactor=2)
    ref = oracle.OracleSim(side=side, seed=4, spawnStabilityFactor=-2, stableStabilityFactor=2)
    for episode in range(2):
        env.reset(); ref.reset()
        state = env.get_stable(vector=True, shallow=True)
        assert np.array_equal(state, ref.get_stable(vector=True))
        for t in range(steps):
            action = select_action(state)
            env.toggle_state(action); ref.toggle_state(action)
            env.step(); ref.step()
            n_state = env.get_stable(vector=True, shallow=True)
            assert n_state is state                                           # aliasing
            assert np.array_equal(n_state, ref.get_stable(vector=True)), (episode, t)
            assert env.reward() == ref.reward(), (episode, t)
            state = n_state
        assert np.array_equal(env.get_state(vector=True), ref.get_state(vector=True))
        assert env.alive() == ref.alive()
