"""CPU suite: the data-parallel batched DQN agent with world_size 2 on gloo (SURVEY.md section 8 rows e + f1).

Each rank owns its own environments (driven by the CPU oracle) and its own replay ring; the only exchange is
the gradient all-reduce inside learn().  Checked: (1) the replicas stay identical while acting and sampling
differently, (2) one learn() on two half batches equals one learn() of a single process on the union batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, out_dir):
    for p in (PKG, ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cgl_b200.dqn import BatchedDQNAgent
    from oracle_env import OracleBatchEnv
    torch.set_num_threads(1)
    env = OracleBatchEnv(4, 5, seed=10 * rank)                       # different envs per rank
    agent = BatchedDQNAgent(env, lr=1e-3, tau=0.05, max_size=64, batch_size=6, seed=3, hidden=24, group=True)
    state = agent.reset()
    actions = []
    for _ in range(8):
        a = agent.select_action(state, 0.5)
        actions.append(a.clone())
        state, _ = agent.step(a)
    # (2) a controlled learn(): rank r contributes half r of a fixed union batch
    g = torch.Generator().manual_seed(99)
    S = torch.randint(-3, 4, (12, 25), dtype=torch.int8, generator=g)
    N = torch.randint(-3, 4, (12, 25), dtype=torch.int8, generator=g)
    A = torch.randint(0, 26, (12, 1), dtype=torch.int32, generator=g)
    R = torch.randint(-50, 50, (12, 1), dtype=torch.int32, generator=g)
    before = {k: v.clone() for k, v in agent.Q.state_dict().items()}
    qt_before = {k: v.clone() for k, v in agent.Q_target.state_dict().items()}
    import copy
    opt_before = copy.deepcopy(agent.optimizer.state_dict())
    sl = slice(6 * rank, 6 * rank + 6)
    agent.learn((S[sl], A[sl], R[sl], N[sl]), 0.9)
    torch.save({"Q": agent.Q.state_dict(), "Qt": agent.Q_target.state_dict(), "before": before, "opt": opt_before,
                "Qt_before": qt_before, "actions": torch.stack(actions), "union": (S, A, R, N)},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_data_parallel_agent_world_size_2(tmp_path):
    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"), weights_only=False)
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"), weights_only=False)
    # replicas identical after 8 loop steps + the controlled update, although they acted differently
    for k in r0["Q"]:
        assert torch.equal(r0["Q"][k], r1["Q"][k]) and torch.equal(r0["Qt"][k], r1["Qt"][k]), k
        assert torch.equal(r0["before"][k], r1["before"][k]), k
    assert not torch.equal(r0["actions"], r1["actions"])
    # the controlled update == a single process learning on the union batch from the same starting point
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cgl_b200.dqn import BatchedDQNAgent
    from oracle_env import OracleBatchEnv
    solo = BatchedDQNAgent(OracleBatchEnv(4, 5, seed=0), lr=1e-3, tau=0.05, max_size=64, batch_size=6, seed=3, hidden=24)
    solo.Q.load_state_dict(r0["before"])
    solo.Q_target.load_state_dict(r0["Qt_before"])
    solo.optimizer.load_state_dict(r0["opt"])
    solo.learn(r0["union"], 0.9)
    for k, v in solo.Q.state_dict().items():
        assert torch.allclose(v, r0["Q"][k], rtol=1e-5, atol=1e-6), k
    for k, v in solo.Q_target.state_dict().items():
        assert torch.allclose(v, r0["Qt"][k], rtol=1e-5, atol=1e-6), k
