"""Test helper: the BatchedSim protocol implemented on the CPU oracle (test infrastructure only)."""
import numpy as np
import torch

from oracle import oracle


class OracleBatchEnv:
    """The BatchedSim protocol on the CPU oracle (reference semantics, CGL/CGL.py:211-243 per env)."""

    def __init__(self, n_envs, side, seed=0, spawn=-2, stable_max=2):
        self.n_envs, self.side, self.size = n_envs, side, side * side
        self.device = torch.device("cpu")
        self.spawn, self.stable_max = spawn, stable_max
        self.world = np.stack([oracle.initial_world(side, seed + e) for e in range(n_envs)])
        self.stable = torch.from_numpy(np.stack([oracle.initial_stable(w, spawn) for w in self.world]))
        self._w0, self._s0 = self.world.copy(), self.stable.clone()

    def bind_observation(self, buf):
        buf.view(self.n_envs, self.size).copy_(self.stable)
        self.stable = buf.view(self.n_envs, self.size)

    def reset(self):
        self.world[...] = self._w0
        self.stable.copy_(self._s0)
        return self.stable

    def step(self, actions=None, obs_out=None, reward_out=None):
        if obs_out is not None:
            obs_out.copy_(self.stable)
            self.stable = obs_out.view(self.n_envs, self.size)
        acts = None if actions is None else actions.numpy()
        rew, _ = oracle.step_batch(self.world, self.stable.numpy(), self.side, acts, self.spawn, self.stable_max)
        rew = torch.from_numpy(rew)
        if reward_out is not None:
            reward_out.copy_(rew)
            rew = reward_out
        return self.stable, rew, None
