"""Host side of the B200-native Game-of-Life env step.

Modules: native (ctypes binding of the C ABI), batched (BatchedSim: env-index sharded batch
driver: step / step into a replay ring / on-chip runs / fork rules / checkpoints), dqn (the batched
DQN loop: trajectory replay ring, agent, data-parallel learn), bands (RowBandLife: row-band sharded
large grids with NVLink halo exchange).
The reference-facing drop-in classes live one level up: CGL.py (`import CGL; CGL.sim(...)`) and
CGL_action+/CGL.py for the fork.
"""
from . import native  # noqa: F401
