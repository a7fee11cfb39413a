"""HostRollout -- the env side of the reference's training loop driven from HOST buffers, double-buffered.

Reference loop body, /root/reference/CGL/main.py:64-72: `toggle_state(action)` (CGL/CGL.py:322-328) -> `step()`
(:247-252) -> `reward()` (:255-256), one env, everything synchronous, both planes over PCIe twice per step (:203-208).

Here B environments live on the device, split into `n_groups` groups that are stepped alternately on their own
streams by `cgl_rollout_run` (include/cgl_b200.h): while one group's step runs on the GPU, the host reads the other
group's rewards, lets `policy` write its next actions and enqueues its next step.  Per step and env 4 bytes go in
(the action, from a pinned buffer, H2D copy inside the step's CUDA graph) and 4 bytes come out (the reward, written
by the kernel straight into pinned host memory).  The loop itself runs in C; `policy` is the only Python in it.

    ro = HostRollout(4096, 128, spawnStabilityFactor=-2, stableStabilityFactor=2)
    def policy(group, step, rewards, actions):      # numpy int32 views of the group's pinned buffers
        actions[:] = my_policy(rewards)             # rewards: the group's previous step (zeros before the first)
    ro.run(1000, policy)
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import native
from .batched import BatchedSim


class HostRollout:
    def __init__(self, n_envs: int, side: int, n_groups: int = 2, n_replicas: int = 1, seed: int = 0,
                 spawnStabilityFactor: int = -1, stableStabilityFactor: int = 1, device="cuda", rng: str = "reference",
                 first_env: int = 0, obs_to_host: bool = False, zero_copy_actions: bool = False):
        if n_envs % n_groups:
            raise ValueError("n_envs must be divisible by n_groups")
        self._lib = native.load()
        self.n_envs, self.side, self.size = n_envs, side, side * side
        self.n_groups, self.n_replicas, self.per_group = n_groups, n_replicas, n_envs // n_groups
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise native.CglNativeError("HostRollout needs a CUDA device (there is no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # sims[g][r]: group g, replica r.  Env e of replica r is seeded seed + r * n_envs + first_env + e (global index).
        self.sims = [[BatchedSim(self.per_group, side, seed=seed + r * n_envs, first_env=first_env + g * self.per_group,
                                 spawnStabilityFactor=spawnStabilityFactor, stableStabilityFactor=stableStabilityFactor,
                                 device=self.device, rng=rng) for r in range(n_replicas)] for g in range(n_groups)]
        flat = [s for grp in self.sims for s in grp]
        V = ctypes.c_void_p
        n = len(flat)
        wa = (V * n)(*[s._wa.data_ptr() for s in flat])
        wb = (V * n)(*[s._wb.data_ptr() for s in flat])
        st = (V * n)(*[s.stable.data_ptr() for s in flat])
        self.obs = None
        obs_ptrs = None
        if obs_to_host:
            self._obs_t = [torch.empty((self.per_group, self.size), dtype=torch.int8).pin_memory() for _ in range(n_groups)]
            self.obs = [t.numpy() for t in self._obs_t]
            obs_ptrs = (V * n_groups)(*[t.data_ptr() for t in self._obs_t])
        self._h = V()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            native.check(self._lib.cgl_rollout_create(ctypes.byref(self._h), n_groups, n_replicas, wa, wb, st,
                                                      self.per_group, side, spawnStabilityFactor, stableStabilityFactor,
                                                      obs_ptrs, int(bool(zero_copy_actions))), "cgl_rollout_create")
        self.actions, self.rewards = [], []
        for g in range(n_groups):
            a = ctypes.POINTER(ctypes.c_int32)()
            r = ctypes.POINTER(ctypes.c_int32)()
            native.check(self._lib.cgl_rollout_buffers(self._h, g, ctypes.byref(a), ctypes.byref(r)), "cgl_rollout_buffers")
            self.actions.append(np.ctypeslib.as_array(a, shape=(self.per_group,)))
            self.rewards.append(np.ctypeslib.as_array(r, shape=(self.per_group,)))
        self.steps = 0
        self.zero_copy_actions = bool(zero_copy_actions)
        self.h2d_bytes_per_step = 4 * n_envs
        self.d2h_bytes_per_step = 4 * n_envs + (n_envs * self.size if obs_to_host else 0)

    def run(self, steps: int, policy=None) -> None:
        """`steps` steps of every group.  policy(group, step, rewards, actions) -> None fills `actions` (numpy int32
        view of the group's pinned action buffer) for the step about to be enqueued; `rewards` is the group's pinned
        reward buffer holding its previous step's rewards.  None: the action buffers are used as they are."""
        cb = None
        if policy is not None:
            acts, rews = self.actions, self.rewards

            def trampoline(_user, group, step, _r, _a):
                policy(group, step, rews[group], acts[group])
            cb = native.POLICY_FN(trampoline)
        with torch.cuda.device(self.device):
            rc = self._lib.cgl_rollout_run(self._h, steps, ctypes.cast(cb, ctypes.c_void_p) if cb else None, None)
        if rc:
            native.check(rc, "cgl_rollout_run")
        first, self.steps = self.steps, self.steps + steps
        R = self.n_replicas
        for g, grp in enumerate(self.sims):                 # keep the BatchedSim views of the planes in step
            for r, s in enumerate(grp):
                par = self._lib.cgl_rollout_parity(self._h, g, r)
                cur_is_b = s._plane_id[s._wa.data_ptr()] == 2
                if bool(par) != cur_is_b:
                    s._wa, s._wb = s._wb, s._wa
                c = sum(1 for t in range(first, first + steps) if t % R == r) if steps < 4 * R else \
                    (first + steps - r + R - 1) // R - (first - r + R - 1) // R
                s.count += c
                s.launches += c * self._lib.cgl_env_step_launches(self.side, 1)

    def sim(self, group: int, replica: int = 0) -> BatchedSim:
        """The BatchedSim holding (group, replica)'s planes (for get_state / checks between runs)."""
        return self.sims[group][replica]

    def close(self) -> None:
        if self._h:
            self._lib.cgl_rollout_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
