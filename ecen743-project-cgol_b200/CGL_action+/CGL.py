"""CGL -- drop-in replacement for the fork's module `CGL_action+/CGL.py`, B200-native.

Put THIS directory on sys.path (as the fork's scripts do with theirs) and `import CGL`:

    env = CGL.sim(side=10, seed=0, gpu=True, spawnStabilityFactor=-2, stableStabilityFactor=2, empty=-1, empty_min=-6)
    env.toggle_state(block); env.step(); obs = env.get_stable(vector=True, shallow=True); r = env.stability()

Same constructor keywords and methods as the fork's `class sim`
(/root/reference/CGL/CGL_action+/CGL.py:53-411).  It is the base facade (../CGL.py) with the fork's
differences switched on in the library (include/cgl_b200.h, cgl_env_step_rule):

  * dead cells start at `empty` (:122-126) and, after every step, follow the fork's dead-cell rule.  The fork's
    two back ends DISAGREE on that rule: its CUDA kernel (:190-193) lets them fall by one per step down to
    `empty_min`; its CPU step (:256) sets them to min(stable + empty, empty_min).  `gpu=True` is the fork's
    CUDA path, so the default here is dead_rule="decay"; pass dead_rule="sat" to reproduce its CPU step
    (that one is pinned by golden vectors recorded from the fork, tests/golden/golden_action_plus.npz).
  * toggle_state writes SPAWN only to cells that are alive after the toggle, 0 to the others (:382-384);
  * runBlank, fresh(seed), stability(), breakdown_stable/state, get_max_density, update_state(newState,
    newStability), save/load carrying `empty`.

There is no CPU path: gpu=False and step(forceCPU=True) raise.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
if _PKG not in sys.path:
    sys.path.insert(1, _PKG)                                  # for `cgl_b200`
_spec = importlib.util.spec_from_file_location("cgl_b200_base_facade", os.path.join(_PKG, "CGL.py"))
_base = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_base)
GPU_CAPABLE = _base.GPU_CAPABLE


class sim(_base.sim):
    def __init__(self, state=None, side=8, seed=8, gpu=False, gpu_select=0, warp=8, spawnStabilityFactor=-1,
                 stableStabilityFactor=1, runBlank=False, empty=0, empty_min=-128, dead_rule="decay"):
        # the fork's extra validation (CGL_action+/CGL.py:69-82); the rest is the base constructor's
        if not isinstance(runBlank, bool):
            raise TypeError("runBlank must be a bool!")
        if not isinstance(empty, int):
            raise TypeError("empty must be integer!")
        if not isinstance(empty_min, int):
            raise TypeError("empty_min must be integer!")
        _base._check_int8("empty", empty)
        _base._check_int8("empty_min", empty_min)
        if dead_rule not in ("decay", "sat"):
            raise ValueError('dead_rule must be "decay" (the fork\'s CUDA kernel) or "sat" (its CPU step)')
        self.empty, self.empty_min, self.dead_rule = empty, empty_min, dead_rule
        self._run_blank = runBlank
        super().__init__(state=state, side=side, seed=seed, gpu=gpu, gpu_select=gpu_select, warp=warp,
                         spawnStabilityFactor=spawnStabilityFactor, stableStabilityFactor=stableStabilityFactor)
        self.max_density = self.get_max_density()

    def _env_variant(self):
        return dict(dead_rule=self.dead_rule, empty=self.empty, empty_min=self.empty_min, masked_toggle=True)

    def _initial_world(self):
        if self._run_blank:                                   # CGL_action+/CGL.py:113-114
            np.random.seed(self.seed)
            return np.zeros(self.size, dtype=np.uint8)
        return None

    # ---- the fork's additions ---------------------------------------------------------------------
    def stability(self):
        """np.int32 sum of the stability vector (CGL_action+/CGL.py:272-273; the base env calls it reward)."""
        return self.reward()

    def alive(self):
        """np.int32 like the fork (CGL_action+/CGL.py:269-270)."""
        return np.int32(super().alive())

    def fresh(self, seed=0):
        """A new random world from `seed`, stability re-initialised (CGL_action+/CGL.py:282-287); count,
        initState and initStable are left alone like in the fork."""
        self._pending = None
        self._lazy_steps = 0
        self._wait()
        self._invalidate()
        np.random.seed(seed)
        cells = np.random.randint(2, size=self.size, dtype=np.uint8)
        self._b.set_state(self._torch.from_numpy(cells)[None, :].to(self._dev))
        self._b.init_stable()
        self._changed()

    def get_max_density(self):
        """Largest number of live cells a still life can have on a side x side board
        (CGL_action+/CGL.py:343-360).  Values for side <= 60 are the published optima of Chu & Stuckey,
        "A complete solution to the Maximum Density Still Life Problem", Artificial Intelligence 184-185
        (2012), table 7; above that their closed form floor(n^2/2 + 17n/27 - 2) with the -1 residue classes
        mod 54 of their theorem 6."""
        n = self.side
        if n <= 60:
            return MAX_DENSITY_STILL_LIFE[n]
        bump = 2 if n % 54 in MDSL_MINUS_TWO_RESIDUES else 1
        return np.floor((self.size / 2) + (17 / 27) * n - bump)

    def update_state(self, newState, newStability=None):
        """Replace the world and optionally the stability plane (CGL_action+/CGL.py:363-372)."""
        temp = newState.flatten().astype(np.uint8)
        if temp.size != self.size:
            raise ValueError("The new state must have the same size and side as the original state!\n"
                             f"Was given size={temp.size} but was expecting size={self.size} and side={self.side}.")
        cells = self._validated_cells(temp)
        self._flush()
        self._wait()
        self._alive_valid = False
        self._b.set_state(self._torch.from_numpy(cells)[None, :].to(self._dev))
        if newStability is not None:
            stab = np.ascontiguousarray(newStability.flatten().astype(np.int8))
            if stab.size != self.size:
                raise ValueError(f"newStability must have {self.size} cells")
            self._b.stable.copy_(self._torch.from_numpy(stab)[None, :])
            self._reward_valid = False
        self._changed(world=True, stable=newStability is not None)

    def save(self):
        """(world, stable, side, count, spawn, stable_max, empty)  (CGL_action+/CGL.py:388-389)."""
        return super().save() + (self.empty,)

    def load(self, newState, newstable, side, count, spawnStabilityFactor, stableStabilityFactor, empty):
        """CGL_action+/CGL.py:392-411."""
        if not isinstance(empty, int):
            raise TypeError("empty must be integer!")
        _base._check_int8("empty", empty)
        self.empty = empty
        super().load(newState, newstable, side, count, spawnStabilityFactor, stableStabilityFactor)
        self._b.set_factors(spawnStabilityFactor, stableStabilityFactor, empty)
        self._fast_args.clear()


# Maximum-density still-life optima for n x n boards, n = 0..60 (Chu & Stuckey 2012, table 7).
MAX_DENSITY_STILL_LIFE = (
    0, 0, 4, 6, 8, 16, 18, 28, 36, 43, 54, 64, 76, 90, 104, 119, 136, 152, 171, 190, 210, 232, 253, 276, 302, 326,
    353, 379, 407, 437, 467, 497, 531, 563, 598, 633, 668, 706, 744, 782, 824, 864, 907, 949, 993, 1039, 1085, 1132,
    1181, 1229, 1280, 1331, 1382, 1436, 1490, 1545, 1602, 1658, 1717, 1776, 1835)
MDSL_MINUS_TWO_RESIDUES = frozenset((0, 1, 3, 8, 9, 11, 16, 17, 19, 25, 27, 31, 33, 39, 41, 47, 49))
