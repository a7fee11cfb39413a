#!/usr/bin/env python
"""Time the REFERENCE's own Python CPU step (/root/reference/CGL/CGL.py:211-243) in this container.

    python tools/time_reference_cpu.py            # writes profiles/reference_cpu_timing.json

The reference is pure Python and /root/reference does not exist on the GPU box, so this is the one number of the
report that is measured on the BUILD box: BASELINE.md section 3's recipe, verbatim -- `sys.path.insert(0,
'/root/reference/CGL')`, GPU_CAPABLE=false, `CGL.sim(side, seed, gpu=False, spawnStabilityFactor=-2,
stableStabilityFactor=2)`, the loop of CGL/main.py:64-72 with random actions (`toggle_state`, `step`,
`get_stable(vector=True, shallow=True)`, `reward`), time.perf_counter around the loop only, median of 3 repeats;
plus the plain-step loop of CGL/bench.py:39-40.  bench.py reads the JSON for `extras.c1...reference_python_*`."""
import json
import os
import platform
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CGL_REFERENCE", "/root/reference")
os.environ["GPU_CAPABLE"] = "false"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "CGL"))
import numpy as np  # noqa: E402
import CGL  # noqa: E402  (the reference's module)


def loop_with_actions(side, n):
    env = CGL.sim(side=side, seed=0, gpu=False, spawnStabilityFactor=-2, stableStabilityFactor=2)
    acts = np.random.RandomState(123).randint(side * side + 1, size=n).astype(np.int32)
    t0 = time.perf_counter()
    acc = 0
    for a in acts:
        env.toggle_state(a)
        env.step()
        env.get_stable(vector=True, shallow=True)
        acc += int(env.reward())
    return n / (time.perf_counter() - t0), acc


def loop_plain(side, n):
    env = CGL.sim(side=side, seed=0, gpu=False, spawnStabilityFactor=-2, stableStabilityFactor=2)
    t0 = time.perf_counter()
    for _ in range(n):
        env.step()
    return n / (time.perf_counter() - t0), int(env.reward()), int(env.alive())


def main():
    out = {"what": "reference CGL.sim(gpu=False), unmodified, imported from /root/reference/CGL",
           "host": {"cpu": platform.processor() or platform.machine(), "os_cpu_count": os.cpu_count(), "cores_used": 1,
                    "python": platform.python_version(), "numpy": np.__version__},
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "loops": {}}
    for side, n in ((64, 40), (10, 1500), (128, 10)):
        rates = [loop_with_actions(side, n) for _ in range(3)]
        plain = [loop_plain(side, n) for _ in range(3)]
        out["loops"][f"side{side}"] = {
            "steps_timed": n,
            "dqn_loop_env_steps_per_s": statistics.median(r[0] for r in rates), "reward_sum": rates[0][1],
            "plain_step_loop_steps_per_s": statistics.median(p[0] for p in plain),
            "plain_loop_reward_alive": list(plain[0][1:]),
            "cell_updates_per_s": statistics.median(r[0] for r in rates) * side * side}
        print(side, out["loops"][f"side{side}"], flush=True)
    with open(os.path.join(ROOT, "profiles", "reference_cpu_timing.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
