"""CPU suite: the product's bit logic (cgl_bits.cuh compiled by g++ as tests/twin) vs the oracle.

Catches logic bugs in the bit-sliced rule, the torus word handling, the nibble/LUT plumbing and
the byte-SIMD stability update without a GPU.  The GPU suite then checks the kernels themselves."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, TRACES
from oracle import oracle

TWIN_DIR = os.path.join(ROOT, "tests", "twin")


@pytest.fixture(scope="module")
def twin():
    subprocess.run(["make", "-C", TWIN_DIR], check=True, stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(os.path.join(TWIN_DIR, "libcgl_twin.so"))
    vp, u64, u32, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    L.twin_pack.argtypes = [vp, vp, u64, u32, u32]
    L.twin_unpack.argtypes = [vp, vp, u64, u32, u32]
    L.twin_life_generic.argtypes = [vp, vp, u64, u32, u32, ci]
    L.twin_stable_generic.argtypes = [vp, vp, vp, u64, u32, ci, ci]
    L.twin_env_step_fused.argtypes = [vp, vp, vp, u64, u32, vp, ci, ci, vp, vp]
    L.twin_env_step_fused.restype = ci
    return L


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


def pack(twin, cells, n_envs, rows, cols):
    W = (cols + 31) // 32
    out = np.zeros(n_envs * rows * W, np.uint32)
    cells = np.ascontiguousarray(cells, np.uint8)
    twin.twin_pack(P(cells), P(out), n_envs, rows, cols)
    return out


def unpack(twin, world, n_envs, rows, cols):
    out = np.zeros(n_envs * rows * cols, np.uint8)
    twin.twin_unpack(P(world), P(out), n_envs, rows, cols)
    return out


def test_pack_unpack_roundtrip_and_bit_order(twin):
    rs = np.random.RandomState(0)
    for rows, cols in ((1, 1), (3, 5), (7, 32), (5, 33), (4, 64), (9, 100)):
        cells = rs.randint(2, size=(2, rows * cols)).astype(np.uint8)
        w = pack(twin, cells, 2, rows, cols)
        assert np.array_equal(unpack(twin, w, 2, rows, cols), cells.reshape(-1))
    w = pack(twin, np.array([[1, 0, 0, 1, 0]], np.uint8), 1, 1, 5)
    assert w.tolist() == [0b01001]            # bit j of word w = column 32*w + j


@pytest.mark.parametrize("name", sorted(TRACES))
def test_generic_path_replays_reference_trace(twin, name):
    tr = TRACES[name]
    side = tr.side
    w = pack(twin, tr.worlds[0], 1, side, side)
    s = tr.stables[0].copy()
    for t in range(tr.T):
        a = tr.action(t)
        if a is not None:                      # host-side toggle on the unpacked copy, as the oracle does
            cells = unpack(twin, w, 1, side, side)
            oracle.toggle(cells, s, a, tr.spawn)
            w = pack(twin, cells, 1, side, side)
        nxt = np.zeros_like(w)
        twin.twin_life_generic(P(w), P(nxt), 1, side, side, 1)
        twin.twin_stable_generic(P(w), P(nxt), P(s), 1, side, tr.spawn, tr.stable_max)
        w = nxt
        assert np.array_equal(unpack(twin, w, 1, side, side), tr.worlds[t + 1]), (name, t)
        assert np.array_equal(s, tr.stables[t + 1]), (name, t)


@pytest.mark.parametrize("name", [n for n in sorted(TRACES) if TRACES[n].side % 32 == 0 and TRACES[n].actions.shape[1] == 1])
def test_fused_path_replays_reference_trace(twin, name):
    tr = TRACES[name]
    side = tr.side
    w = pack(twin, tr.worlds[0], 1, side, side)
    s = tr.stables[0].copy()
    rew, alv = np.zeros(1, np.int32), np.zeros(1, np.uint32)
    for t in range(tr.T):
        a = tr.action(t)
        acts = np.array([side * side if a is None else a], np.int32)
        nxt = np.zeros_like(w)
        assert twin.twin_env_step_fused(P(w), P(nxt), P(s), 1, side, P(acts), tr.spawn, tr.stable_max, P(rew), P(alv)) == 0
        w = nxt
        assert np.array_equal(unpack(twin, w, 1, side, side), tr.worlds[t + 1]), (name, t)
        assert np.array_equal(s, tr.stables[t + 1]), (name, t)
        assert rew[0] == tr.rewards[t + 1] and alv[0] == tr.alives[t + 1]


@pytest.mark.parametrize("side,spawn,smax", [(32, -2, 2), (64, -128, 127), (96, 5, 2), (128, -1, 1), (64, 127, -128), (32, 0, 0)])
def test_fused_path_random_states_vs_oracle(twin, side, spawn, smax):
    """Arbitrary stability bytes (incl. s > MAX, wrap at 127) and random actions, 6 envs, 4 steps."""
    rs = np.random.RandomState(side * 1000 + spawn + 500)
    B, size = 6, side * side
    cells = rs.randint(2, size=(B, size)).astype(np.uint8)
    st = rs.randint(-128, 128, size=(B, size)).astype(np.int8)
    w = pack(twin, cells, B, side, side)
    s = st.copy()
    for _ in range(4):
        acts = rs.randint(size + 1, size=B).astype(np.int32)
        rew_o, alv_o = oracle.step_batch(cells, st, side, acts, spawn, smax, threads=2)
        nxt = np.zeros_like(w)
        rew, alv = np.zeros(B, np.int32), np.zeros(B, np.uint32)
        assert twin.twin_env_step_fused(P(w), P(nxt), P(s), B, side, P(acts), spawn, smax, P(rew), P(alv)) == 0
        w = nxt
        assert np.array_equal(unpack(twin, w, B, side, side), cells.reshape(-1))
        assert np.array_equal(s, st)
        assert np.array_equal(rew, rew_o) and np.array_equal(alv, alv_o)


@pytest.mark.parametrize("rows,cols,wrap", [(1, 1, 1), (2, 2, 1), (3, 3, 1), (5, 31, 1), (6, 32, 1), (7, 33, 1), (9, 64, 1),
                                            (4, 65, 1), (8, 100, 1), (5, 96, 0), (6, 40, 0)])
def test_generic_life_rectangles_vs_oracle(twin, rows, cols, wrap):
    rs = np.random.RandomState(rows * 1000 + cols)
    cells = rs.randint(2, size=(rows, cols)).astype(np.uint8)
    w = pack(twin, cells, 1, rows, cols)
    for _ in range(3):
        nxt = np.zeros_like(w)
        twin.twin_life_generic(P(w), P(nxt), 1, rows, cols, wrap)
        w = nxt
        if wrap:
            cells = oracle.life(cells, 1)
        else:  # open rows, torus columns: pad one dead row above and below, wrap columns
            padded = np.zeros((rows + 2, cols), np.uint8)
            padded[1:-1] = cells
            full = np.zeros_like(padded)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if dy or dx:
                        sh = np.roll(padded, dx, 1)
                        sh = np.vstack([np.zeros((1, cols), np.uint8), sh[:-1]]) if dy == 1 else (
                            np.vstack([sh[1:], np.zeros((1, cols), np.uint8)]) if dy == -1 else sh)
                        full = full + sh
            cells = (((full == 3) | ((full == 2) & (padded == 1))).astype(np.uint8))[1:-1]
        assert np.array_equal(unpack(twin, w, 1, rows, cols).reshape(rows, cols), cells)


@pytest.mark.parametrize("spawn,smax", [(-2, 2), (-128, 127), (5, 127), (100, 120), (0, 0), (-1, -1), (7, -3)])
def test_bit_sliced_stability_equals_scalar_rule(twin, spawn, smax):
    """cgl_bits.cuh: byte <-> bit-plane transposes are exact inverses and the 32-cells-per-op rule equals the
    scalar rule for every int8 value under every transition."""
    twin.twin_sliced_mismatches.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
    twin.twin_sliced_mismatches.restype = ctypes.c_uint64
    rs = np.random.RandomState(abs(spawn) + 3)
    # all 256 values x 3 transitions, each at a random position among random neighbours, plus pure random groups
    groups = 256 * 3 + 500
    cells = rs.randint(-128, 128, size=(groups, 32)).astype(np.int8)
    tr = rs.randint(0, 3, size=(groups, 32)).astype(np.uint8)
    for v in range(256):
        for t in range(3):
            g = v * 3 + t
            pos = rs.randint(32)
            cells[g, pos] = np.int8(v - 128)
            tr[g, pos] = t
    cells[768:800] = np.int8(smax)                      # whole groups at the ceiling
    assert twin.twin_sliced_mismatches(P(cells), P(tr), groups, spawn, smax) == 0


@pytest.mark.parametrize("spawn,smax,emin", [(-2, 2, -6), (-128, 127, -128), (5, 127, -3), (-2, 3, -128), (0, 0, 0), (7, -3, 100)])
def test_bit_sliced_decay_rule_equals_scalar_rule(twin, spawn, smax, emin):
    twin.twin_sliced_decay_mismatches.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int]
    twin.twin_sliced_decay_mismatches.restype = ctypes.c_uint64
    rs = np.random.RandomState(abs(spawn) + abs(emin))
    groups = 256 * 3 + 400
    cells = rs.randint(-128, 128, size=(groups, 32)).astype(np.int8)
    tr = rs.randint(0, 3, size=(groups, 32)).astype(np.uint8)
    for v in range(256):
        for t in range(3):
            pos = rs.randint(32)
            cells[v * 3 + t, pos] = np.int8(v - 128)
            tr[v * 3 + t, pos] = t
    cells[768:790] = np.int8(emin)                      # whole groups on the floor / at the ceiling
    cells[790:812] = np.int8(smax)
    assert twin.twin_sliced_decay_mismatches(P(cells), P(tr), groups, spawn, smax, emin) == 0


def test_packed_byte_generation_equals_scalar_rule(twin):
    """life_next4_bytes (the single-env server's generation, 4 cells per word): EVERY 3 x 6 neighbourhood of a
    4-cell group against the scalar B3/S23 rule."""
    n = 1 << 18
    bits = ((np.arange(n, dtype=np.uint32)[:, None] >> np.arange(18, dtype=np.uint32)[None, :]) & 1).reshape(n, 3, 6)
    rows = bits.astype(np.uint32)                                  # rows[:, r, c]: column c = x0 - 1 + c
    words = [np.ascontiguousarray(sum(rows[:, r, 1 + k] << (8 * k) for k in range(4)).astype(np.uint32)) for r in range(3)]
    lc = np.ascontiguousarray(rows[:, :, 0].sum(axis=1).astype(np.uint32))
    rc = np.ascontiguousarray(rows[:, :, 5].sum(axis=1).astype(np.uint32))
    out = np.zeros(n, np.uint32)
    twin.twin_life_next4_bytes.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_uint64]
    twin.twin_life_next4_bytes(P(words[0]), P(words[1]), P(words[2]), P(lc), P(rc), P(out), n)
    for k in range(4):
        cnt = rows[:, :, k:k + 3].sum(axis=(1, 2)) - rows[:, 1, 1 + k]
        want = ((cnt == 3) | ((cnt == 2) & (rows[:, 1, 1 + k] == 1))).astype(np.uint32)
        assert np.array_equal((out >> (8 * k)) & 0xff, want), k


@pytest.mark.parametrize("spawn,smax,empty,emin", [(-2, 2, -1, -6), (-128, 127, -128, -128), (5, 127, 3, -3), (-2, 3, -1, 127),
                                                   (0, 0, 0, 0), (7, -3, 100, 100), (-2, 2, 127, 5), (-100, 20, -50, -120)])
def test_bit_sliced_saturating_rule_equals_scalar_rule(twin, spawn, smax, empty, emin):
    """stable_update_sliced_sat (the fork's CPU rule on absolute bit planes): every int8 value under every transition,
    including sums that wrap before the signed minimum is taken."""
    twin.twin_sliced_sat_mismatches.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int]
    twin.twin_sliced_sat_mismatches.restype = ctypes.c_uint64
    rs = np.random.RandomState(abs(spawn) + abs(emin) + abs(empty))
    groups = 256 * 3 + 400
    cells = rs.randint(-128, 128, size=(groups, 32)).astype(np.int8)
    tr = rs.randint(0, 3, size=(groups, 32)).astype(np.uint8)
    for v in range(256):
        for t in range(3):
            pos = rs.randint(32)
            cells[v * 3 + t, pos] = np.int8(v - 128)
            tr[v * 3 + t, pos] = t
    cells[768:790] = np.int8(emin)
    cells[790:812] = np.int8(smax)
    assert twin.twin_sliced_sat_mismatches(P(cells), P(tr), groups, spawn, smax, empty, emin) == 0
