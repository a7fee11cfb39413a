"""GPU suite: the thin PyTorch C++ extension (cgl_b200/_cgl_ext.so) against the ctypes binding and the CPU oracle --
both are faces of the same C ABI, so results must be identical."""
import numpy as np
import pytest

from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import native
    return native.ext()


@pytest.mark.parametrize("side", [5, 10, 33, 64, 128])
def test_pack_unpack_reward_alive(ext, side):
    rs = np.random.RandomState(side)
    n, size = 4, side * side
    cells = rs.randint(2, size=(n, size)).astype(np.uint8)
    stab = rs.randint(-128, 128, size=(n, size)).astype(np.int8)
    world = ext.pack(torch.from_numpy(cells).cuda(), side, side)
    assert world.shape == (n, side, (side + 31) // 32) and world.dtype == torch.int32
    assert np.array_equal(ext.unpack(world, side, side).cpu().numpy(), cells)
    assert np.array_equal(ext.reward(torch.from_numpy(stab).cuda(), size).cpu().numpy(), stab.astype(np.int32).sum(1))
    alive = ext.alive(world, side * ((side + 31) // 32)).cpu().numpy().view(np.uint32)
    assert np.array_equal(alive, cells.astype(np.uint32).sum(1))


@pytest.mark.parametrize("side", [10, 64, 128])
@pytest.mark.parametrize("rule", [(0, 0, 0, False), (1, -1, -5, True), (2, -1, -5, True)])
def test_env_step_and_run_match_oracle(ext, side, rule):
    dead_rule, empty, emin, masked = rule
    rs = np.random.RandomState(side + dead_rule)
    n, size = 5, side * side
    cells = rs.randint(2, size=(n, size)).astype(np.uint8)
    stab = rs.randint(-128, 128, size=(n, size)).astype(np.int8)
    wa = ext.pack(torch.from_numpy(cells).cuda(), side, side)
    wb = torch.empty_like(wa)
    st = torch.from_numpy(stab.copy()).cuda()
    rew = torch.empty(n, dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t in range(4):
        acts = rs.randint(size + 1, size=n).astype(np.int32)
        ext.env_step(wa, wb, st, st, torch.from_numpy(acts).cuda(), rew, None, err, None, side, -2, 2, dead_rule, empty,
                     emin, masked)
        wa, wb = wb, wa
        for e in range(n):
            if acts[e] < size:
                (oracle.toggle_masked if masked else oracle.toggle)(cells[e], stab[e], int(acts[e]), -2)
            oracle.step_rule(cells[e], stab[e], side, -2, 2, dead_rule, empty, emin)
        assert np.array_equal(ext.unpack(wa, side, side).cpu().numpy(), cells), t
        assert np.array_equal(st.cpu().numpy(), stab) and np.array_equal(rew.cpu().numpy(), stab.astype(np.int32).sum(1))
    steps = ext.env_run(wa, st, side, 6, False, -2, 2, dead_rule, empty, emin, rew, None)
    for e in range(n):
        assert oracle.run_rule(cells[e], stab[e], side, -2, 2, 6, dead_rule, empty, emin, False) == int(steps[e])
    assert np.array_equal(ext.unpack(wa, side, side).cpu().numpy(), cells) and np.array_equal(st.cpu().numpy(), stab)
    assert int(err.item()) == 0


def test_life_run_and_argument_checks(ext):
    rows, cols = 192, 4096
    grid = np.random.RandomState(3).randint(2, size=(rows, cols)).astype(np.uint8)
    a = ext.pack(torch.from_numpy(grid.reshape(1, -1)).cuda(), rows, cols).reshape(-1)
    b = torch.empty_like(a)
    in_a = ext.life_run(a, b, rows, cols, True, 24, 8)
    out = ext.unpack(a if in_a else b, rows, cols).cpu().numpy().reshape(rows, cols)
    assert np.array_equal(out, oracle.life(grid, 24, threads=4))
    with pytest.raises(RuntimeError):
        ext.pack(torch.zeros(16, dtype=torch.uint8), 4, 4)              # CPU tensor: there is no CPU path
    with pytest.raises(RuntimeError):
        ext.reward(torch.zeros(16, dtype=torch.int32, device="cuda"), 16)   # wrong dtype
