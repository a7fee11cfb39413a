"""GPU suite at BASELINE.json's full sizes: size-independent properties + oracle checks on samples.

C2 4096 x 128^2 and C3 65536 x 64^2 (env mode): a spread sample of envs is replayed by the CPU oracle,
and for ALL envs reward == sum(stable) and alive == popcount(world).  C4 65536^2 / C5 32768^2 (life mode):
light-cone windows cut from the torus (incl. across the wrap seams) are advanced by the oracle on the CPU
and compared with the GPU result; k-blocked and streamed generations must agree bit for bit."""
import numpy as np
import pytest

from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import native
    return native.load(), native


@pytest.mark.parametrize("n_envs,side", [(4096, 128), (65536, 64)])
def test_env_mode_full_batch(gpu, n_envs, side):
    from cgl_b200.batched import BatchedSim
    size = side * side
    env = BatchedSim(n_envs, side, seed=5, spawnStabilityFactor=-2, stableStabilityFactor=2, rng="device")
    sample = np.unique(np.concatenate([[0, 1, n_envs - 1, n_envs // 2], np.random.RandomState(0).randint(n_envs, size=20)]))
    sidx = torch.from_numpy(sample).cuda()
    cells = env.get_state()[sidx].cpu().numpy()
    st = env.stable[sidx].cpu().numpy().copy()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for step in range(4):
        acts = torch.randint(0, size + 1, (n_envs,), dtype=torch.int32, device="cuda", generator=g)
        obs, rew, _ = env.step(acts, want_alive=True)
        rew_o, alv_o = oracle.step_batch(cells, st, side, acts[sidx].cpu().numpy(), -2, 2, threads=8)
        assert np.array_equal(env.get_state()[sidx].cpu().numpy(), cells), step
        assert np.array_equal(obs[sidx].cpu().numpy(), st), step
        assert np.array_equal(rew[sidx].cpu().numpy(), rew_o)
        # whole batch: the fused reductions agree with independent reductions of the planes
        assert torch.equal(rew.to(torch.int64), obs.to(torch.int64).sum(1))
        words = env.world.view(n_envs, -1)
        pop = torch.zeros(n_envs, dtype=torch.int64, device="cuda")
        for b in range(32):
            pop += ((words >> b) & 1).sum(1)
        assert torch.equal(env.last_alive(), pop)
        assert torch.equal(env.reward(), rew) and torch.equal(env.alive(), pop)
    env.check_actions()


def _window(words, n, r0, c0, h, w):
    """uint8 cells of the h x w window at (r0, c0) of the n x n torus held as packed int32 words [n, n/32]."""
    rows = (torch.arange(r0, r0 + h, device=words.device) % n)
    cols = (torch.arange(c0, c0 + w, device=words.device) % n)
    sel = words[rows][:, cols // 32]
    return ((sel >> (cols % 32).to(torch.int32)) & 1).to(torch.uint8).cpu().numpy()


@pytest.mark.parametrize("n,k,gens", [(65536, 8, 16), (32768, 16, 16), (32768, 4, 10), (32768, 1, 5)])
def test_life_mode_light_cone_windows(gpu, n, k, gens):
    lib, native = gpu
    W = n // 32
    g = torch.Generator(device="cuda"); g.manual_seed(n + k)
    a = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, W), dtype=torch.int32, device="cuda", generator=g)
    start = a.clone()
    b = torch.empty_like(a)
    res = native.ctypes.c_int(-1)
    native.check(lib.cgl_life_run(native.dptr(a), native.dptr(b), n, n, 1, gens, k, native.ctypes.byref(res),
                                  native.current_stream()))
    out = a if res.value == 1 else b
    h = w = 96
    for (r0, c0) in [(0, 0), (n - 40, n - 50), (n // 2 - 7, 31), (12345 % n, n - 20), (n - 3, 1000)]:
        before = _window(start, n, r0 - gens, c0 - gens, h + 2 * gens, w + 2 * gens)
        want = oracle.life_open(before, gens)[gens:gens + h, gens:gens + w]
        got = _window(out, n, r0, c0, h, w)
        assert np.array_equal(got, want), (n, k, r0, c0)
    # live-cell count of the result by the library == independent popcount
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    native.check(lib.cgl_alive(native.dptr(out), 1, out.numel(), native.dptr(cnt), native.current_stream()))
    pop = 0
    for bit in range(32):
        pop += int(((out >> bit) & 1).sum())
    assert (int(cnt.item()) & 0xFFFFFFFF) == pop & 0xFFFFFFFF


def test_temporal_blocking_depths_agree_at_full_size(gpu):
    lib, native = gpu
    n, gens = 32768, 24
    W = n // 32
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    start = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, W), dtype=torch.int32, device="cuda", generator=g)
    results = []
    for k in (1, 3, 8, 12):
        a, b = start.clone(), torch.empty_like(start)
        res = native.ctypes.c_int(-1)
        native.check(lib.cgl_life_run(native.dptr(a), native.dptr(b), n, n, 1, gens, k, native.ctypes.byref(res),
                                      native.current_stream()))
        results.append((a if res.value == 1 else b).clone())
        del a, b
    for r in results[1:]:
        assert torch.equal(r, results[0])
