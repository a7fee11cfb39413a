"""pytest configuration: registers the `gpu` marker and puts the product directory on sys.path.

The product directory is `ecen743-project-cgol_b200/` (not an importable identifier on purpose:
like the reference's `CGL/` directory it is put on sys.path and its `CGL.py` is imported bare,
CGL/main.py:1, CGL/bench.py:5)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ecen743-project-cgol_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("CGL_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Trace:
    """One recorded run of the reference (see tests/golden/make_golden.py)."""

    def __init__(self, z, name):
        self.name = name
        self.side = int(z[f"{name}/side"])
        self.size = self.side * self.side
        self.spawn = int(z[f"{name}/spawn"])
        self.stable_max = int(z[f"{name}/stable_max"])
        self.worlds = np.unpackbits(z[f"{name}/worlds"], axis=1)[:, :self.size]
        self.stables = z[f"{name}/stables"]
        self.rewards = z[f"{name}/rewards"]
        self.alives = z[f"{name}/alives"]
        self.rewards_after_toggle = z[f"{name}/rewards_after_toggle"]
        self.actions = z[f"{name}/actions"]
        self.T = self.actions.shape[0]

    def action(self, t):
        """None | int | list[int] exactly as it was passed to toggle_state."""
        a = self.actions[t]
        a = a[a >= 0]
        if a.size == 0:
            return None
        if self.actions.shape[1] == 1:
            return int(a[0])
        return [int(v) for v in a]


def _load_traces():
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_traces.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: Trace(z, n) for n in names}


TRACES = _load_traces()


@pytest.fixture(scope="session")
def traces():
    return TRACES


class ForkTrace(Trace):
    """One recorded run of the CGL_action+ fork's CPU back end (tests/golden/make_golden_action_plus.py)."""

    def __init__(self, z, name):
        self.name = name
        self.side = int(z[f"{name}/side"])
        self.size = self.side * self.side
        self.spawn = int(z[f"{name}/spawn"])
        self.stable_max = int(z[f"{name}/stable_max"])
        self.empty = int(z[f"{name}/empty"])
        self.empty_min = int(z[f"{name}/empty_min"])
        self.worlds = np.unpackbits(z[f"{name}/worlds"], axis=1)[:, :self.size]
        self.stables = z[f"{name}/stables"]
        self.toggled_stables = z[f"{name}/toggled_stables"]
        self.stability = z[f"{name}/stability"]
        self.alives = z[f"{name}/alives"]
        self.actions = z[f"{name}/actions"]
        self.max_density = float(z[f"{name}/max_density"])
        self.T = self.actions.shape[0]


def _load_fork_traces():
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_action_plus.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: ForkTrace(z, n) for n in names}


FORK_TRACES = _load_fork_traces()
