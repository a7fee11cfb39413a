import os, sys, json, statistics, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ecen743-project-cgol_b200"))
import torch
from cgl_b200.batched import BatchedSim, StepSequence
B, SIDE, R = 4096, 128, 4
dev = torch.device("cuda", 0)
sims = [BatchedSim(B, SIDE, seed=r * B, spawnStabilityFactor=-2, stableStabilityFactor=2, device=dev, rng="device") for r in range(R)]
acts = torch.randint(0, SIDE * SIDE + 1, (16, B), dtype=torch.int32, device=dev)
seq = StepSequence(sims, acts)
def region(k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); seq.run(k); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3
seq.run(16); torch.cuda.synchronize()
res = {}
for k in (8, 24, 400):
    ts = [region(k) for _ in range(9)]
    res[k] = round(statistics.median(ts) / k, 2)
print(res)
