"""How does the host-driven rollout's time per full step depend on the number of env groups?  (C2 shape, 4 rotating
replicas like bench.py's e2e.)  Diagnostic for DESIGN.md section 4.8."""
import os, sys, time, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ecen743-project-cgol_b200"))
import numpy as np
import torch
from cgl_b200.rollout import HostRollout

B, SIDE, R = 4096, 128, 4
size = SIDE * SIDE
def policy(group, step, rewards, actions):                  # bench.py's policy: reads a reward, writes an action
    actions[step & 255] = (int(rewards[step & 255]) + step) % (size + 1)


for groups in (1, 2, 4, 8):
    for zero_copy in (False, True):
        ro = HostRollout(B, SIDE, n_groups=groups, n_replicas=R, seed=7, spawnStabilityFactor=-2, stableStabilityFactor=2,
                         rng="device", zero_copy_actions=zero_copy)
        for g in range(groups):
            ro.actions[g][:] = np.random.RandomState(g).randint(size + 1, size=B // groups)
        out = {"groups": groups, "zero_copy": zero_copy}
        for name, pol in (("no_policy", None), ("python_policy", policy)):
            ro.run(8, pol)
            ts = []
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ro.run(200, pol)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) / 200 * 1e6)
            out[name + "_us"] = round(statistics.median(ts), 2)
        ro.close()
        print(json.dumps(out), flush=True)
