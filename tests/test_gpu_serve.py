"""GPU suite: the resident single-env step server (cgl_sim_serve, csrc/cgl_sim1.cu) behind the `CGL.sim` facade.

The reference's training loop (CGL/main.py:64-72) calls toggle_state -> step -> get_stable -> reward on ONE env
thousands of times; the facade serves those steps from a kernel that stays resident between them.  What must hold:
  * every step equals the CPU oracle (= the reference's own step, tests/golden) bit for bit, on every side class
    (ragged rows, sides that are not multiples of 4, several row passes) and every dead-cell rule;
  * the kernel leaves when it is told to or when the host goes quiet, and the next step finds the state intact
    (idle exits, device-wide synchronisation in the middle of a loop, interleaved calls that use other kernels);
  * switching the server off (CGL_SIM_LINGER_US=0) gives the launch-per-step path with identical results.
"""
import ctypes
import importlib.util
import os
import time

import numpy as np
import pytest

from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SPAWN, STABLE = -2, 2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def CGL():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import CGL as mod
    return mod


def dqn_iteration(env, ref, a):
    env.toggle_state(a); env.step()
    ref.toggle_state(a); ref.step()
    obs = env.get_stable(vector=True, shallow=True)
    assert int(env.reward()) == int(ref.reward())
    return obs


@pytest.mark.parametrize("side", [1, 2, 3, 5, 10, 31, 33, 64, 70, 100, 130, 200, 256])
def test_served_loop_matches_oracle(CGL, side):
    size = side * side
    env = CGL.sim(side=side, seed=3, gpu=True, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
    assert env._serve_ok
    ref = oracle.OracleSim(side=side, seed=3, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
    rs = np.random.RandomState(side)
    obs = env.get_stable(vector=True, shallow=True)
    for i in range(60):
        a = np.int32(rs.randint(size + 1))
        obs = dqn_iteration(env, ref, a)
        if i % 7 == 0:
            assert np.array_equal(obs, ref.stable), (side, i)
            assert int(env.alive()) == int(ref.alive())
    assert env._serving                                       # the whole loop ran on the resident kernel
    assert np.array_equal(env.get_state(vector=True), ref.world)      # (another kernel reads the planes: server leaves)
    assert not env._serving
    assert np.array_equal(env.get_stable(vector=True), ref.stable)
    assert env.get_count() == 60


@pytest.mark.parametrize("side", [10, 64, 96])
def test_idle_exit_sync_and_interleaved_calls(CGL, side):
    """The server leaves on its own when the host goes quiet; a device-wide synchronise in the loop returns; calls that
    run other kernels (reset, multi-index toggle, run, update_state, match, save/load) interleave freely."""
    size = side * side
    env = CGL.sim(side=side, seed=5, gpu=True, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
    ref = oracle.OracleSim(side=side, seed=5, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
    rs = np.random.RandomState(11 + side)
    env.get_stable(vector=True, shallow=True)
    for rnd in range(6):
        for _ in range(9):
            dqn_iteration(env, ref, np.int32(rs.randint(size + 1)))
        if rnd == 0:
            time.sleep(0.01)                                  # >> linger: the kernel has left by itself
            assert int(env._res[4]) == env._launch_id
        elif rnd == 1:
            t0 = time.perf_counter()
            torch.cuda.synchronize()                          # must return once the server lingers out
            assert time.perf_counter() - t0 < 1.0
        elif rnd == 2:
            idx = rs.randint(size, size=5)
            env.toggle_state(idx); ref.toggle_state(idx)
        elif rnd == 3:
            assert env.run(5) == 5
            for _ in range(5):
                ref.step()
        elif rnd == 4:
            w = rs.randint(2, size=size).astype(np.uint8)
            env.update_state(w, side)
            ref.world = w.copy()
            assert env.match(w)
        else:
            saved = env.save()
            ref_saved = (ref.world.copy(), ref.stable.copy(), ref.count)
            env.reset(); ref.reset()
            dqn_iteration(env, ref, np.int32(0))
            env.load(*saved)
            ref.world, ref.stable, ref.count = ref_saved
    for _ in range(5):
        obs = dqn_iteration(env, ref, np.int32(rs.randint(size + 1)))
    assert np.array_equal(obs, ref.stable)
    assert np.array_equal(env.get_state(vector=True), ref.world)
    assert env.get_count() == ref.count


def test_two_envs_served_side_by_side(CGL):
    envs = [CGL.sim(side=s, seed=s, gpu=True, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE) for s in (12, 64)]
    refs = [oracle.OracleSim(side=s, seed=s, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE) for s in (12, 64)]
    rs = np.random.RandomState(1)
    for e in envs:
        e.get_stable(vector=True, shallow=True)
    for _ in range(40):
        for e, r in zip(envs, refs):
            obs = dqn_iteration(e, r, np.int32(rs.randint(e.size + 1)))
            assert np.array_equal(obs, r.stable)
    assert all(e._serving for e in envs)
    del envs                                                  # __del__ stops the servers before the buffers go


def test_fork_rules_are_served(CGL):
    spec = importlib.util.spec_from_file_location("cgl_fork_facade", os.path.join(ROOT, "ecen743-project-cgol_b200",
                                                                                  "CGL_action+", "CGL.py"))
    fork = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fork)
    side, size = 20, 400
    for rule, code in (("decay", oracle.DEAD_DECAY), ("sat", oracle.DEAD_SAT)):
        env = fork.sim(side=side, seed=4, gpu=True, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE, empty=-1,
                       empty_min=-6, dead_rule=rule)
        cells = env.get_state(vector=True).copy()
        st = env.get_stable(vector=True).copy()
        rs = np.random.RandomState(9)
        obs = env.get_stable(vector=True, shallow=True)
        for _ in range(50):
            a = int(rs.randint(size + 1))
            env.toggle_state(a); env.step()
            if a < size:
                oracle.toggle_masked(cells, st, a, SPAWN)
            oracle.step_rule(cells, st, side, SPAWN, STABLE, code, -1, -6)
            obs = env.get_stable(vector=True, shallow=True)
            assert int(env.stability()) == int(oracle.reward(st))
        assert env._serving
        assert np.array_equal(obs, st) and np.array_equal(env.get_state(vector=True), cells), rule


def test_server_switched_off_gives_the_same_results(CGL, monkeypatch):
    monkeypatch.setenv("CGL_SIM_LINGER_US", "0")
    side, size = 48, 48 * 48
    env = CGL.sim(side=side, seed=8, gpu=True, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
    assert not env._serve_ok
    ref = oracle.OracleSim(side=side, seed=8, spawnStabilityFactor=SPAWN, stableStabilityFactor=STABLE)
    rs = np.random.RandomState(2)
    env.get_stable(vector=True, shallow=True)
    for _ in range(30):
        obs = dqn_iteration(env, ref, np.int32(rs.randint(size + 1)))
    assert not env._serving and np.array_equal(obs, ref.stable)


def test_serve_c_abi_protocol():
    """cgl_sim_serve through the C ABI with a hand-driven mailbox: commands, answers, QUIT, idle exit, relaunch."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import native
    lib = native.load()
    side, size = 37, 37 * 37
    W = (side + 31) // 32
    rs = np.random.RandomState(0)
    cells = rs.randint(2, size=size).astype(np.uint8)
    st = rs.randint(-128, 128, size=size).astype(np.int8)
    wa = torch.zeros(side * W, dtype=torch.int32, device="cuda")
    wb = torch.zeros_like(wa)
    c_d, s_d = torch.from_numpy(cells).cuda(), torch.from_numpy(st.copy()).cuda()
    native.check(lib.cgl_pack(native.dptr(c_d), native.dptr(wa), 1, side, side, native.current_stream()))
    torch.cuda.synchronize()
    mirror = torch.zeros(size, dtype=torch.int8).pin_memory()
    res_t = torch.zeros(8, dtype=torch.int32).pin_memory()
    cmd_t = torch.zeros(2, dtype=torch.int64).pin_memory()
    res, cmd = res_t.numpy(), cmd_t.numpy().view(np.uint64)
    a = native.SimStepArgs()
    a.world_a, a.world_b, a.stable = wa.data_ptr(), wb.data_ptr(), s_d.data_ptr()
    a.side, a.spawn, a.stable_max = side, SPAWN, STABLE
    a.dead_rule = a.empty = a.empty_min = a.masked_toggle = 0
    a.obs_mirror, a.result = mirror.data_ptr(), res_t.data_ptr()
    stream = torch.cuda.Stream()

    def wait(cond, what):
        t0 = time.perf_counter()
        while not cond():
            assert time.perf_counter() - t0 < 5.0, what

    seq = 0
    for launch in (1, 2):
        native.check(lib.cgl_sim_serve(ctypes.byref(a), cmd_t.data_ptr(), seq, launch, 2000, stream.cuda_stream))
        for _ in range(8):
            act = int(rs.randint(size + 1))
            seq += 1
            cmd[0] = (seq << 32) | act
            wait(lambda: int(res[2]) == seq, "no answer")
            if act < size:
                oracle.toggle(cells, st, act, SPAWN)
            oracle.step(cells, st, side, SPAWN, STABLE)
            assert int(res[0]) == int(oracle.reward(st)) and int(res[1]) == int(oracle.alive(cells))
            assert np.array_equal(mirror.numpy(), st)
        if launch == 1:
            seq += 1
            cmd[0] = (seq << 32) | native.SIM_QUIT            # told to leave
        wait(lambda: int(res[4]) == launch, "did not leave")  # (second launch: leaves after 2 ms of silence)
        stream.synchronize()
        native.check(lib.cgl_unpack(native.dptr(wa), native.dptr(c_d), 1, side, side, native.current_stream()))
        assert np.array_equal(c_d.cpu().numpy(), cells) and np.array_equal(s_d.cpu().numpy(), st)
    assert lib.cgl_sim_serve(ctypes.byref(a), cmd_t.data_ptr(), 0, 1, 0, stream.cuda_stream) != 0     # linger 0 refused
