"""Where does the time of a SHORT timed region go?  T(K) for K env steps (4096 x 128^2, 4 rotating replicas)
bracketed by synchronize on both sides: Python-eager launches, the C launch loop (cgl_env_step_seq) and CUDA-graph
replay; plus a per-step event timeline.  Diagnostic for bench.py (the driver times --steps 20)."""
import os, sys, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ecen743-project-cgol_b200"))
import torch
from cgl_b200.batched import BatchedSim, StepSequence

B, SIDE, R = 4096, 128, 4
dev = torch.device("cuda", 0)
sims = [BatchedSim(B, SIDE, seed=r * B, spawnStabilityFactor=-2, stableStabilityFactor=2, device=dev, rng="device") for r in range(R)]
acts = torch.randint(0, SIDE * SIDE + 1, (16, B), dtype=torch.int32, device=dev)
seq = StepSequence(sims, acts)


def eager(k):
    for i in range(k):
        sims[i % R].step(acts[i % 16])


def region(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


for k in (8, 8, 8):
    eager(k); seq.run(k)
torch.cuda.synchronize()
out = {}
for name, fn in (("python_eager", eager), ("c_seq", seq.run)):
    res = {}
    for k in (8, 16, 24, 40, 80, 160, 400):        # multiples of 2R: planes return to their start
        ts = [region(fn, k) for _ in range(9)]
        res[k] = {"median_us": round(statistics.median(ts), 1), "min_us": round(min(ts), 1), "per_step_median": round(statistics.median(ts) / k, 2)}
    out[name] = res
    print(name, json.dumps(res))
# timeline: one event after every step of a 24-step region issued by the C loop, one step per call
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(25)]
one = [StepSequence([sims[j]], acts) for j in range(R)]
for rep in range(3):
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(24):
        sims[i % R].step(acts[i % 16])
        evs[i + 1].record()
    torch.cuda.synchronize()
    print("timeline_us", [round(evs[i].elapsed_time(evs[i + 1]) * 1e3, 1) for i in range(24)])
