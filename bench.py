#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Game-of-Life env step.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (default N=1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's step on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU (driver does this)

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): 4096 envs x 128x128 per GPU, env mode --
per step and env: toggle_state(action) -> generation -> int8 stability update -> reward
(/root/reference/CGL/main.py:64-72).  Weak scaling: every rank owns its own 4096 envs (sharding by
env index, no data-path collective).  A "step" is one launch over all envs of the rank.

Prints ONE JSON line (rank 0).  `value` is in G cell-updates/s over all ranks with state resident in
HBM; `e2e` is the same metric through the host-buffer C-ABI call (actions H2D + reward D2H per step);
`roofline` is algorithmic bytes (2.25 B/cell-update) / measured launch time vs the measured HBM peak;
`cpu_baseline` is the oracle port of the reference's per-cell loop timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ecen743-project-cgol_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

SIDE, ENVS_PER_GPU = 128, 4096
SPAWN, STABLE = -2, 2                      # CGL/main.py:29, CGL/bench.py:12-13
BYTES_PER_CELL_ENV = 2.25                  # 1 bit R + 1 bit W + int8 R + int8 W  
BYTES_PER_CELL_LIFE = 0.25
L2_BYTES = 126e6


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's per-cell loop (the reference itself is pure Python
# and cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------
def cpu_arm(steps, warmup, budget_s=8.0, threads=None):
    from oracle import oracle
    threads = threads or oracle.max_threads()
    size = SIDE * SIDE
    # calibrate on a few envs, then size the sample so a step stays <= ~0.4 s
    n = 4 * threads
    w = np.stack([oracle.initial_world(SIDE, e) for e in range(n)])
    s = np.stack([oracle.initial_stable(w[e], SPAWN) for e in range(n)])
    rs = np.random.RandomState(10 ** 6)
    t0 = time.perf_counter()
    oracle.step_batch(w, s, SIDE, rs.randint(size + 1, size=n).astype(np.int32), SPAWN, STABLE, threads)
    per_env = (time.perf_counter() - t0) / n
    n_envs = int(max(threads, min(ENVS_PER_GPU, 0.4 / per_env)))
    max_steps = max(1, int(budget_s / (per_env * n_envs)))
    steps = min(steps, max_steps)
    w = np.ascontiguousarray(np.resize(w, (n_envs, size)))
    s = np.ascontiguousarray(np.resize(s, (n_envs, size)))
    acts = rs.randint(size + 1, size=(warmup + steps, n_envs)).astype(np.int32)
    for i in range(warmup):
        oracle.step_batch(w, s, SIDE, acts[i], SPAWN, STABLE, threads)
    t0 = time.perf_counter()
    for i in range(steps):
        oracle.step_batch(w, s, SIDE, acts[warmup + i], SPAWN, STABLE, threads)
    dt = time.perf_counter() - t0
    gcups = n_envs * size * steps / dt / 1e9
    return {"value": gcups, "unit": "Gcell-updates/s", "cores": threads, "kind": "port",
            "sample": f"{n_envs} envs x {SIDE}x{SIDE} x {steps} steps (toggle+step+reward), oracle/cgl_oracle.c "
                      f"= C port of CGL/CGL.py:211-243, {threads} host threads",
            "env_steps_per_s": n_envs * steps / dt, "ms_per_step": dt / steps * 1e3, "steps": steps}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_arm(args.steps, min(args.warmup, 2), budget_s=20.0)
    line = {"impl": "reference", "metric": "life_cell_updates_per_s", "value": base["value"], "unit": base["unit"],
            "n_gpus": args.gpus, "steps": base["steps"], "warmup": min(args.warmup, 2),
            "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args.gpus, None),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "env_steps_per_s": base["env_steps_per_s"], "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, replicas):
    return {"workload": f"C2: {ENVS_PER_GPU} envs x {SIDE}x{SIDE} per GPU, env mode (toggle + generation + int8 "
                        f"stability + reward), spawn {SPAWN} / stable {STABLE}",
            "envs_per_gpu": ENVS_PER_GPU, "side": SIDE, "total_envs": ENVS_PER_GPU * n_gpus,
            "parallelism": f"env-index sharding x{n_gpus}, no data-path collective",
            "l2": None if replicas is None else
            f"inputs larger than L2: {replicas} rotating replicas of the env batch "
            f"({replicas * ENVS_PER_GPU * SIDE * SIDE * 1.25 / 2**20:.0f} MiB touched round-robin > 126 MB L2)"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def time_steps(torch, fn, steps, barrier):
    """K calls of fn(i) bracketed by barrier + synchronize, CUDA events on the launching stream."""
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=4, help="rotating env batches (working set > L2)")
    ap.add_argument("--launch", default="eager", choices=["graph", "eager"],
                    help="how the K timed steps are launched: one C-ABI call per step (default; consecutive launches "
                         "overlap through programmatic dependent launch + per-env chaining) or CUDA-graph replay")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3/C4/C5 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from cgl_b200 import native
    from cgl_b200.batched import BatchedSim
    native.load()                                           # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    size = SIDE * SIDE
    B, R, K, W = ENVS_PER_GPU, max(1, args.replicas), args.steps, args.warmup
    sims = [BatchedSim(B, SIDE, seed=rank * 10 ** 5 + r * B, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE,
                       device=dev, rng="device") for r in range(R)]
    n_act = 16
    g = torch.Generator(device=dev); g.manual_seed(10 ** 6 + rank)
    actions = torch.randint(0, size + 1, (n_act, B), dtype=torch.int32, device=dev, generator=g)

    def step(i):
        sims[i % R].step(actions[i % n_act])

    for i in range(W):
        step(i)
    # Launch-bound inner loop -> CUDA graph: one graph = 2R consecutive steps (every replica stepped
    # twice, so the ping-pong planes are back in place); K steps = K // 2R replays + eager remainder.
    cycle = 2 * R
    graph = None
    if True:                                                # the graph variant is always measured too
        side_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side_stream):
            for i in range(cycle):
                step(i)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side_stream):
                for i in range(cycle):
                    step(i)
        graph.replay()
        torch.cuda.synchronize()
    per_step_launches = sims[0]._lib.cgl_env_step_launches(SIDE, 1)

    pos = [0]                                               # eager steps since the planes were last in captured position

    def timed_region(use_graph):
        if use_graph:
            # a captured graph bakes the plane pointers in: replay it only from the captured position
            for i in range((-pos[0]) % cycle):
                step(pos[0] + i)
            pos[0] = 0
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = 0
        if use_graph:
            for _ in range(K // cycle):
                graph.replay()
            done = (K // cycle) * cycle
        for i in range(done, K):
            step(pos[0] + i - done)
        pos[0] = (pos[0] + K - done) % cycle
        e1.record()
        torch.cuda.synchronize()
        barrier()
        return e0.elapsed_time(e1) / 1e3

    # clocks are sampled from here until the last timing of this workload (headline region, then
    # the eager and L2-resident variants of the same K steps) so that short regions still get samples
    sampler = ClockSampler(local).start() if rank == 0 else None
    dt = max_over_ranks(timed_region(args.launch == "graph"))
    gpu_launches = K * per_step_launches
    dt_other = max_over_ranks(timed_region(args.launch != "graph"))
    cells_per_step = B * size * world
    value = cells_per_step * K / dt / 1e9
    ms_per_step = dt / K * 1e3

    # L2-resident variant (one replica, 80 MiB working set inside the 126 MB L2) -- reported, not the headline
    dt_l2 = max_over_ranks(time_steps(torch, lambda i: sims[0].step(actions[i % n_act]), K, barrier))
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the host-buffer C-ABI call (actions H2D, reward D2H every step) ----
    e2e = None
    if not args.no_e2e:
        acts_h = [torch.randint(0, size + 1, (B,), dtype=torch.int32).pin_memory() for _ in range(4)]
        rew_h = torch.empty(B, dtype=torch.int32).pin_memory()
        ke = max(10, min(K, 100))
        for i in range(3):
            sims[i % R].step_host(acts_h[i % 4], rew_h)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(ke):
            sims[i % R].step_host(acts_h[i % 4], rew_h)     # returns after the D2H copy completed
        torch.cuda.synchronize()
        dte = max_over_ranks(time.perf_counter() - t0)
        barrier()
        obs_h = torch.empty((B, size), dtype=torch.int8).pin_memory()
        ko = 5
        sims[0].step_host(acts_h[0], rew_h, obs_h)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(ko):
            sims[i % R].step_host(acts_h[i % 4], rew_h, obs_h)
        torch.cuda.synchronize()
        dto = max_over_ranks(time.perf_counter() - t0)
        # double-buffered rollout: the same B envs as two groups on two streams; each group's step is enqueued
        # without a sync, and before a group is stepped again the host waits for ITS previous step, reads a
        # reward and writes an action -- one group's host round trip overlaps the other group's kernel
        Bg = B // 2
        groups = [[BatchedSim(Bg, SIDE, seed=rank * 10 ** 5 + 31 * (2 * r + gi) + 5,
                              spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE, device=dev, rng="device")
                   for r in range(R)] for gi in range(2)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        acts_g = [[torch.randint(0, size + 1, (Bg,), dtype=torch.int32).pin_memory() for _ in range(4)] for _ in range(2)]
        rew_g = [torch.zeros(Bg, dtype=torch.int32).pin_memory() for _ in range(2)]
        rew_np = [r.numpy() for r in rew_g]
        act_np = [[a.numpy() for a in acts] for acts in acts_g]
        torch.cuda.synchronize()
        seen = 0

        raw = [ctypes.c_void_p(st.cuda_stream) for st in streams]

        def pipelined(n_it):
            nonlocal seen
            for i in range(n_it):
                for gi in range(2):
                    sim = groups[gi][i % R]
                    sim.wait_host(raw[gi])                            # this group's previous step has landed
                    seen += int(rew_np[gi][0])                        # the host looks at a result ...
                    act_np[gi][i % 4][0] = (seen + i) % (size + 1)    # ... and decides an action
                    sim.step_host(acts_g[gi][i % 4], rew_g[gi], sync=False, stream=raw[gi])
            for gi in range(2):
                groups[gi][0].wait_host(raw[gi])

        pipelined(3)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipelined(ke)
        dtp = max_over_ranks(time.perf_counter() - t0)
        barrier()
        del groups
        e2e = {"value": cells_per_step * ke / dte / 1e9, "unit": "Gcell-updates/s",
               "double_buffered": {"value": cells_per_step * ke / dtp / 1e9, "unit": "Gcell-updates/s", "groups": 2,
                                   "api": "two BatchedSim groups of B/2 envs on two streams, step_host(sync=False) + "
                                          "wait_host(): same H2D / D2H bytes per step, per-group data dependence kept"},
               "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 4 * B, "steps": ke,
               "api": "BatchedSim.step_host -> cgl_env_step_host (pinned actions H2D, step, reward D2H, sync); "
                      "the observation stays device-resident for the GPU Q-network",
               "env_steps_per_s": B * world * ke / dte,
               "with_obs_to_host": {"value": cells_per_step * ko / dto / 1e9, "unit": "Gcell-updates/s",
                                    "d2h_bytes_per_step": 4 * B + B * size, "steps": ko}}

    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_CELL_ENV * B * size / (dt / K) / 1e9          # per GPU, per launch
    kernel = f"env_step_fused_kernel<{SIDE}>"
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": recorded_traffic(kernel), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BYTES_PER_CELL_ENV * B * size,
                "avg_launch_ms": dt / K * 1e3, "l2_resident_variant_gbs": BYTES_PER_CELL_ENV * B * size / (dt_l2 / K) / 1e9}

    extras = {}
    if not args.no_extras:
        try:
            extras = run_extras(torch, dev, rank, world, barrier, max_over_ranks, peak)
        except Exception as exc:  # noqa: BLE001
            extras = {"error": repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_arm(steps=10 ** 6, warmup=1, budget_s=8.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "life_cell_updates_per_s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(world, R),
                "env_steps_per_s": B * world * K / dt, "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches,
                "roofline": roofline, "cpu_baseline": cpu, "l2_resident_value": cells_per_step * K / dt_l2 / 1e9,
                "launch": args.launch,
                ("graph_value" if args.launch == "eager" else "eager_value"): cells_per_step * K / dt_other / 1e9,
                "extras": extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(torch, dev, rank, world, barrier, max_over_ranks, peak):
    """Side measurements of the other BASELINE configs (short; not the headline line)."""
    from cgl_b200 import native
    from cgl_b200.batched import BatchedSim
    lib = native.load()
    out = {}
    # C3: 65536 envs x 64x64, sharded by env index over the ranks (strong scaling over N)
    side, total = 64, 65536
    sim = BatchedSim.shard(total, side, rank, world, seed=0, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE,
                           device=dev, rng="device")
    acts = torch.randint(0, side * side + 1, (8, sim.n_envs), dtype=torch.int32, device=dev)
    for i in range(5):
        sim.step(acts[i % 8])
    k = 50
    dt = max_over_ranks(time_steps(torch, lambda i: sim.step(acts[i % 8]), k, barrier))
    out["c3_envs65536_side64"] = {"gcups": total * side * side * k / dt / 1e9, "env_steps_per_s": total * k / dt,
                                  "ms_per_step": dt / k * 1e3, "scaling": "strong",
                                  "hbm_frac": BYTES_PER_CELL_ENV * sim.n_envs * side * side / (dt / k) / 1e9 / peak,
                                  "working_set_mib_per_gpu": sim.n_envs * side * side * 1.25 / 2 ** 20}
    del sim, acts
    torch.cuda.empty_cache()
    # C4: 65536^2 torus, row bands over the ranks, k = 8 generations per launch / per halo exchange
    from cgl_b200.bands import RowBandLife
    # one GPU: plain torus, 8 generations per launch.  N > 1: ghost depth 64 between exchanges (8 GPUs measured
    # 243 TCUPS against 232 with depth 32, 219 with 16 generations per launch, 208 with the fused exchange)
    n, k, gens = 65536, (8 if world == 1 else 64), (224 if world == 1 else 256)
    band = RowBandLife(n, n, k=k, rank=rank, world_size=world, device=dev, kernel_k=8)
    band.randomize(1)
    band.run(2 * k)
    dt = max_over_ranks(time_steps(torch, lambda i: band.run(gens), 1, barrier))
    out["c4_life_65536_bands"] = {"gcups": n * n * gens / dt / 1e9, "ms_per_gen": dt / gens * 1e3, "k": 8, "ghost_depth": k, "gens": gens,
                                  "exchange": band.exchange, "scaling": "strong", "alive": band.alive(),
                                  "hbm_frac_algorithmic": BYTES_PER_CELL_LIFE * n * n / world / (dt / gens) / 1e9 / peak}
    band.close()
    del band
    torch.cuda.empty_cache()
    if world == 1:
        # C1: the reference's own single-env loop shape (CGL/main.py:64-72 with random actions) through the
        # drop-in facade, 64x64; beside it the oracle port running the same loop on one host core
        import CGL
        from oracle import oracle
        side, n = 64, 2000
        acts = np.random.RandomState(123).randint(side * side + 1, size=n + 100).astype(np.int32)
        env = CGL.sim(side=side, seed=0, gpu=True, gpu_select=dev.index, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
        obs = env.get_stable(vector=True, shallow=True)
        acc = 0
        for i in range(100):
            env.toggle_state(acts[i]); env.step(); obs = env.get_stable(vector=True, shallow=True); acc += int(env.reward())
        t0 = time.perf_counter()
        for i in range(100, 100 + n):
            env.toggle_state(acts[i]); env.step(); obs = env.get_stable(vector=True, shallow=True); acc += int(env.reward())
        dt_f = time.perf_counter() - t0
        ref = oracle.OracleSim(side=side, seed=0, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
        acc_o = 0
        for i in range(100):
            ref.toggle_state(acts[i]); ref.step(); acc_o += int(ref.reward())
        t0 = time.perf_counter()
        for i in range(100, 100 + n):
            ref.toggle_state(acts[i]); ref.step(); acc_o += int(ref.reward())
        dt_o = time.perf_counter() - t0
        # the reference's own bench loop (CGL/bench.py:39-40: plain steps, then Stability and Life are printed):
        # the facade defers the plain steps and runs them as one on-chip launch when the result is asked for
        # (a fresh env: the one above handed out a live shallow view, which switches the deferral off)
        env2 = CGL.sim(side=side, seed=0, gpu=True, gpu_select=dev.index, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
        ref.reset()
        env2.step(); env2.step(); env2.step(); env2.step(); env2.reward()
        t0 = time.perf_counter()
        for _ in range(n):
            env2.step()
        r_f, a_f = int(env2.reward()), int(env2.alive())
        dt_b = time.perf_counter() - t0
        for _ in range(n + 4):
            ref.step()
        del env2
        plain_equal = (r_f, a_f) == (int(ref.reward()), int(ref.alive()))
        out["c1_single_64x64_loop"] = {"facade_env_steps_per_s": n / dt_f, "oracle_port_1core_env_steps_per_s": n / dt_o,
                                       "facade_plain_step_loop_steps_per_s": n / dt_b, "plain_loop_equal": plain_equal,
                                       "reference_python_env_steps_per_s": 76.0,
                                       "reference_source": "BASELINE.md section 2 (measured in the survey container)",
                                       "rewards_equal": acc == acc_o}
        del env
        # C5: 32768^2 torus, sweep of the temporal-blocking depth k on one GPU
        n = 32768
        words = n * (n // 32)
        a = torch.randint(-2 ** 31, 2 ** 31 - 1, (words,), dtype=torch.int32, device=dev)
        b = torch.empty_like(a)
        res = native.ctypes.c_int(0)
        sweep = {}
        for k in (1, 2, 4, 8, 16):
            g = 48

            def run(_i):
                native.check(lib.cgl_life_run(native.dptr(a), native.dptr(b), n, n, 1, g, k, native.ctypes.byref(res),
                                              native.current_stream()))
            if k > 1:        # set-up: let the library time its strip lengths for this shape (RowBandLife does the same)
                native.check(lib.cgl_life_tune(native.dptr(a), native.dptr(b), n, n, 1, k, native.current_stream()))
            run(0)
            dt = time_steps(torch, run, 2, barrier) / 2
            sweep[f"k{k}"] = {"gcups": n * n * g / dt / 1e9, "us_per_gen": dt / g * 1e6,
                              "hbm_frac_algorithmic": BYTES_PER_CELL_LIFE * n * n / (dt / g) / 1e9 / peak}
        out["c5_life_32768_k_sweep"] = sweep
        del a, b
        torch.cuda.empty_cache()
        # f3: the reference's own bench loop (CGL/bench.py:39-40: plain steps, no actions) as ONE launch per k
        # steps with the environments resident in shared memory (cgl_env_run), config 2 shape
        from cgl_b200.batched import BatchedSim
        env = BatchedSim(4096, 128, seed=0, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE, rng="device")
        runk = {}
        for k in (8, 64):
            env.run(k)
            dt = time_steps(torch, lambda _i: env.run(k), 3, barrier) / 3
            runk[f"k{k}"] = {"gcups": 4096 * 128 * 128 * k / dt / 1e9, "us_per_step": dt / k * 1e6,
                             "env_steps_per_s": 4096 * k / dt}
        out["f3_run_in_smem_4096x128"] = runk
        del env
        # f2: the CGL_action+ fork's rule (dead cells decay to a floor, masked toggle) in the same fused kernel
        g = torch.Generator(device=dev)
        g.manual_seed(11)
        acts = torch.randint(0, 128 * 128 + 1, (4096,), dtype=torch.int32, device=dev, generator=g)
        fork_rates = {}
        for rule in ("decay", "sat"):
            envs = [BatchedSim(4096, 128, seed=i, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE, rng="device",
                               dead_rule=rule, empty=-1, empty_min=-6, masked_toggle=True) for i in range(4)]
            for i in range(8):
                envs[i % 4].step(acts)
            dt = time_steps(torch, lambda i: envs[i % 4].step(acts), 400, barrier) / 400
            fork_rates[rule] = {"gcups": 4096 * 128 * 128 / dt / 1e9, "us_per_step": dt * 1e6,
                                "hbm_frac": BYTES_PER_CELL_ENV * 4096 * 128 * 128 / dt / 1e9 / peak}
            del envs
        out["f2_fork_rule_4096x128"] = fork_rates
        # f1: config 2 under the batched DQN loop (the reference's main.py:58-75 for 4096 envs at once):
        # select_action -> toggle+step+reward into the replay ring -> learn -> target update, all on the device.
        # Network = dqn.py:41-59 at side 128 (16384 -> 32770 -> 32770 -> 16385, 2.15 G parameters, fp32).
        try:
            out["f1_dqn_loop_4096x128"] = dqn_loop_extra(torch, dev, barrier)
        except torch.OutOfMemoryError as exc:                    # another tenant on the GPU: report, do not fail the bench
            out["f1_dqn_loop_4096x128"] = {"skipped": f"out of memory: {str(exc)[:80]}"}
        torch.cuda.empty_cache()
    return out


def dqn_loop_extra(torch, dev, barrier, n_envs=4096, side=128, steps=3):
    from cgl_b200.batched import BatchedSim
    from cgl_b200.dqn import BatchedDQNAgent
    env = BatchedSim(n_envs, side, seed=0, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE, rng="device")
    agent = BatchedDQNAgent(env, max_size=int(1e5), batch_size=64, seed=0)
    res = {"n_envs": n_envs, "side": side, "parameters": sum(p.numel() for p in agent.Q.parameters()),
           "replay_slots": agent.memory.slots, "replay_bytes_moved_per_step": 0}
    state = [agent.reset()]

    def loop(_i):
        action = agent.select_action(state[0], 0.1, out=agent.memory.action_slot())
        state[0], _ = agent.step(action)

    def env_only(_i):
        agent.memory.step(agent.memory.action_slot())

    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        loop(0); loop(0)
        dt = time_steps(torch, loop, steps, barrier) / steps
        res[name] = {"env_steps_per_s": n_envs / dt, "ms_per_batched_step": dt * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = False
    agent.act_dtype = torch.bfloat16                     # opt-in: acting forward on the bf16 tensor cores
    loop(0); loop(0)
    dt = time_steps(torch, loop, steps, barrier) / steps
    res["bf16_acting"] = {"env_steps_per_s": n_envs / dt, "ms_per_batched_step": dt * 1e3}
    agent.act_dtype = None
    env_only(0)
    dt_env = time_steps(torch, env_only, 200, barrier) / 200
    res["env_step_into_ring_us"] = dt_env * 1e6
    res["env_share_of_loop_fp32"] = dt_env / (res["fp32"]["ms_per_batched_step"] / 1e3)
    env.check_actions()
    return res


if __name__ == "__main__":
    main()
