#!/usr/bin/env python
"""Small single-GPU drivers for ncu captures (profiles/README.md lists the exact commands).

    python tools/prof.py env   [--side 128 --envs 4096 --steps 6]
    python tools/prof.py life  [--n 65536 --gens 6 --k 1]
    python tools/prof.py run   [--side 128 --envs 4096 --steps 3 --k 16]    k plain steps per launch, env in smem
    python tools/prof.py fork  [--side 128 --envs 4096 --steps 6]           the CGL_action+ rule (decay, masked toggle)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ecen743-project-cgol_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from cgl_b200 import native  # noqa: E402
from cgl_b200.batched import BatchedSim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["env", "life", "run", "fork"])
ap.add_argument("--side", type=int, default=128)
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--gens", type=int, default=6)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--replicas", type=int, default=4)
a = ap.parse_args()
lib = native.load()
dev = torch.device("cuda", 0)
if a.what == "env":
    sims = [BatchedSim(a.envs, a.side, seed=r * a.envs, spawnStabilityFactor=-2, stableStabilityFactor=2, device=dev,
                       rng="device") for r in range(a.replicas)]
    acts = torch.randint(0, a.side * a.side + 1, (a.envs,), dtype=torch.int32, device=dev)
    for i in range(a.steps):
        sims[i % a.replicas].step(acts)
elif a.what == "run":
    sim = BatchedSim(a.envs, a.side, seed=0, spawnStabilityFactor=-2, stableStabilityFactor=2, device=dev, rng="device")
    for i in range(a.steps):
        sim.run(a.k)
elif a.what == "fork":
    sims = [BatchedSim(a.envs, a.side, seed=r * a.envs, spawnStabilityFactor=-2, stableStabilityFactor=2, device=dev,
                       rng="device", dead_rule="decay", empty=-1, empty_min=-6, masked_toggle=True)
            for r in range(a.replicas)]
    acts = torch.randint(0, a.side * a.side + 1, (a.envs,), dtype=torch.int32, device=dev)
    for i in range(a.steps):
        sims[i % a.replicas].step(acts)
else:
    n = a.n
    x = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * (n // 32),), dtype=torch.int32, device=dev)
    y = torch.empty_like(x)
    res = native.ctypes.c_int(0)
    native.check(lib.cgl_life_run(native.dptr(x), native.dptr(y), n, n, 1, a.gens, a.k, native.ctypes.byref(res),
                                  native.current_stream()))
torch.cuda.synchronize()
print("done")
