#!/usr/bin/env python
"""The batched DQN loop (CGL/main.py:58-75 for many environments) on N GPUs -- torchrun entry point.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 \
        tools/run_dqn.py --side 64 --envs 4096 --steps 50 [--check]

Environments are sharded by index (BatchedSim.shard, no communication); each rank keeps its own replay ring;
the gradient all-reduce in BatchedDQNAgent.learn is the only exchange.  Hyper-parameter flags are main.py's.
--check: after the run all ranks must hold bit-identical Q and Q_target.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ecen743-project-cgol_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from cgl_b200.batched import BatchedSim  # noqa: E402
from cgl_b200.dqn import BatchedDQNAgent  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--side", default=10, type=int)
ap.add_argument("--seed", default=0, type=int)
ap.add_argument("--envs", default=4096, type=int, help="total environments over all GPUs")
ap.add_argument("--steps", default=50, type=int, help="timed batched steps")
ap.add_argument("--warmup", default=3, type=int)
ap.add_argument("--batch-size", default=64, type=int)
ap.add_argument("--discount", default=0.99, type=float)
ap.add_argument("--lr", default=5e-4, type=float)
ap.add_argument("--tau", default=0.001, type=float)
ap.add_argument("--exp-size", default=int(1e5), type=int)
ap.add_argument("--update-freq", default=4, type=int)
ap.add_argument("--epsilon", default=0.1, type=float)
ap.add_argument("--hidden", default=0, type=int, help="hidden width (default 2*action_dim as dqn.py:52)")
ap.add_argument("--act-dtype", default="", choices=["", "bf16"])
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

env = BatchedSim.shard(a.envs, a.side, rank, world, seed=a.seed, spawnStabilityFactor=-2, stableStabilityFactor=2,
                       device=dev, rng="device")
agent = BatchedDQNAgent(env, discount=a.discount, tau=a.tau, lr=a.lr, update_freq=a.update_freq, max_size=a.exp_size // world,
                        batch_size=a.batch_size, seed=a.seed, hidden=a.hidden or None,
                        act_dtype=torch.bfloat16 if a.act_dtype == "bf16" else None, group=True if world > 1 else None)
state = agent.reset()
total = torch.zeros(env.n_envs, dtype=torch.int64, device=dev)


def loop():
    global state
    action = agent.select_action(state, a.epsilon, out=agent.memory.action_slot())
    state, reward = agent.step(action)
    total.add_(reward)


for _ in range(a.warmup):
    loop()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loop()
e1.record()
torch.cuda.synchronize()
dt = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
env.check_actions()
result = {"side": a.side, "envs": a.envs, "n_gpus": world, "steps": a.steps, "ms_per_batched_step": float(dt) / a.steps * 1e3,
          "env_steps_per_s": a.envs * a.steps / float(dt), "parameters": sum(p.numel() for p in agent.Q.parameters()),
          "mean_reward_per_env_step": float(total.float().mean()) / (a.steps + a.warmup)}
ok = True
if a.check and world > 1:
    for net in (agent.Q, agent.Q_target):
        for p in net.parameters():
            ref = p.detach().clone()
            dist.broadcast(ref, src=0)
            ok = ok and bool(torch.equal(ref, p.detach()))
    t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = bool(t.item())
    result["replicas_identical"] = ok
if rank == 0:
    print(json.dumps(result))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
