"""Life-mode timing sweep over the temporal-blocking depth k (BASELINE config 5) on one GPU."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ecen743-project-cgol_b200")):
    sys.path.insert(0, p)
import torch
from cgl_b200 import native
lib = native.load()
n = int(os.environ.get("N", 32768))
rows = int(os.environ.get("ROWS", n))
wrap = int(os.environ.get("WRAP", 1))
ks = [int(x) for x in os.environ.get("KS", "1,2,4,8,16").split(",")]
dev = torch.device("cuda", 0)
a = torch.randint(-2 ** 31, 2 ** 31 - 1, (rows * (n // 32),), dtype=torch.int32, device=dev)
b = torch.empty_like(a)
res = native.ctypes.c_int(0)
for k in ks:
    gens = 48 if k > 1 else 40
    gens = (gens // k) * k
    def run():
        native.check(lib.cgl_life_run(native.dptr(a), native.dptr(b), rows, n, wrap, gens, k, native.ctypes.byref(res), native.current_stream()))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / gens * 1e3)
    print(json.dumps({"cols": n, "rows": rows, "k": k, "rows_knob": os.environ.get("CGL_TB_ROWS"), "us_per_gen": round(best, 2),
                      "gcups": round(rows * n / best / 1e3, 1), "hbm_frac_algorithmic": round(0.25 * rows * n / (best * 1e-6) / 1e9 / 6543.4, 3)}), flush=True)
