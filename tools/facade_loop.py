"""The reference's DQN-loop call sequence (CGL/main.py:59-72) against the drop-in facade, with random
actions instead of a Q-network: toggle_state -> step -> get_stable(shallow) -> reward.  Prints steps/s."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ecen743-project-cgol_b200"))
os.environ.setdefault("CGL_QUIET", "1")
import numpy as np
import CGL
for side in (10, 64, 128, 200):
    env = CGL.sim(side=side, seed=0, gpu=True, spawnStabilityFactor=-2, stableStabilityFactor=2)
    rs = np.random.RandomState(123)
    acts = rs.randint(side * side + 1, size=3000).astype(np.int32)
    env.reset()
    state = env.get_stable(vector=True, shallow=True)
    total = 0
    for i in range(200):
        env.toggle_state(acts[i]); env.step(); n_state = env.get_stable(vector=True, shallow=True); total += env.reward()
    t0 = time.perf_counter()
    n = 2000
    for i in range(200, 200 + n):
        env.toggle_state(acts[i]); env.step(); n_state = env.get_stable(vector=True, shallow=True); total += env.reward()
    dt = time.perf_counter() - t0
    out = {"side": side, "env_steps_per_s": round(n / dt, 1), "us_per_step": round(dt / n * 1e6, 1), "reward_sum": int(total)}
    if getattr(env, "_serving", False):                     # device-side share of the last served step (cgl_sim_serve)
        out["device_ns_cmd_to_step_done"] = int(env._res[5])
        out["device_ns_cmd_to_mirror_fenced"] = int(env._res[6])
    print(json.dumps(out))
