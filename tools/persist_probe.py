"""Timing of the life-mode kernels on one GPU: plain torus (C4 / C5 shapes) and ONE 8-GPU band of C4 as a ring of
one (8192 owned rows + 2 x 64 ghost rows, the in-kernel exchange with itself).  CGL_LIFE_PERSIST=0 gives the
launch-per-pass numbers."""
import os, sys, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ecen743-project-cgol_b200"))
import torch
from cgl_b200.bands import RowBandLife

def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)

out = {"persist": os.environ.get("CGL_LIFE_PERSIST", "1")}
for n, k, gens in ((65536, 8, 1000), (32768, 8, 1000), (32768, 4, 1000), (32768, 16, 992)):
    b = RowBandLife(n, n, k=k, kernel_k=k)
    b.randomize(1)
    b.run(2 * k)
    ms = timed(lambda: b.run(gens))
    out[f"torus{n}_k{k}"] = {"us_per_gen": round(ms / gens * 1e3, 2), "tcups": round(n * n * gens / ms / 1e9, 2)}
    del b
    torch.cuda.empty_cache()
if out["persist"] != "0":
    for ghost in (64, 32):
        b = RowBandLife(8192, 65536, k=ghost, kernel_k=8, exchange="persist")
        b.randomize(1)
        b.run(128)
        gens = 1024
        ms = timed(lambda: b.run(gens))
        out[f"band8192_ring_of_one_ghost{ghost}"] = {"us_per_gen": round(ms / gens * 1e3, 2), "x8_tcups": round(8 * 8192 * 65536 * gens / ms / 1e9, 1)}
        b.close()
        del b
print(json.dumps(out))
