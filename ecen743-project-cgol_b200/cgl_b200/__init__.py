"""Host side of the B200-native Game-of-Life env step.

Modules: native (ctypes binding of the C ABI), batched (BatchedSim: env-index sharded batch
driver), bands (RowBandLife: row-band sharded large grids with NVLink halo exchange).
The reference-facing drop-in class lives one level up in CGL.py (`import CGL; CGL.sim(...)`).
"""
from . import native  # noqa: F401
