"""CPU suite: the N>1 row-band path (ghost-zone schedule + ring halo exchange) with world_size 2 on gloo.

The stepping function is swapped for the host twin of the product's bit logic (tests/twin) because
there is no GPU here; everything else -- band partition, k-row ghost zones, exchange order, block
schedule, checksum -- is the product code in cgl_b200/bands.py."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT
from oracle import oracle

TWIN_DIR = os.path.join(ROOT, "tests", "twin")


def _twin():
    subprocess.run(["make", "-C", TWIN_DIR], check=True, stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(os.path.join(TWIN_DIR, "libcgl_twin.so"))
    L.twin_life_generic.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32,
                                    ctypes.c_uint32, ctypes.c_int]
    L.twin_pack.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
    L.twin_unpack.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
    return L


def _cpu_step_factory():
    L = _twin()

    def block(lib, a, b, rows, cols, wrap_rows, gens, k, stream):
        src, dst = a, b
        for _ in range(gens):
            L.twin_life_generic(src.data_ptr(), dst.data_ptr(), 1, rows, cols, wrap_rows)
            src, dst = dst, src
        return src is a
    return block


def _grid(rows, cols, seed):
    cells = np.random.RandomState(seed).randint(2, size=(rows, cols)).astype(np.uint8)
    L = _twin()
    words = np.zeros(rows * cols // 32, np.uint32)
    L.twin_pack(cells.ctypes.data, words.ctypes.data, 1, rows, cols)
    return cells, torch.from_numpy(words.view(np.int32)).view(rows, cols // 32)


def _worker(rank, world, port, rows, cols, k, gens, out_dir):
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cgl_b200 import bands
    bands._life_block = _cpu_step_factory()
    _, words = _grid(rows, cols, 42)
    b = bands.RowBandLife(rows, cols, k=k, rank=rank, world_size=world, device="cpu", exchange="dist")
    b.set_owned(words[rank * b.band_rows:(rank + 1) * b.band_rows])
    b.run(gens)
    torch.save(b.owned.clone(), os.path.join(out_dir, f"band{rank}.pt"))
    cs = b.checksum()
    if rank == 0:
        torch.save(torch.tensor([cs]), os.path.join(out_dir, "checksum.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("rows,cols,k,gens", [(24, 64, 4, 13), (16, 96, 8, 8)])
def test_two_rank_bands_match_single_torus(tmp_path, rows, cols, k, gens):
    port = 29500 + (os.getpid() + rows) % 2000
    mp.start_processes(_worker, args=(2, port, rows, cols, k, gens, str(tmp_path)), nprocs=2, join=True,
                       start_method="spawn")
    cells, words = _grid(rows, cols, 42)
    want = oracle.life(cells, gens, threads=2)
    got_words = torch.cat([torch.load(os.path.join(tmp_path, f"band{r}.pt")) for r in range(2)], 0).contiguous()
    L = _twin()
    got = np.zeros(rows * cols, np.uint8)
    L.twin_unpack(got_words.data_ptr(), got.ctypes.data, 1, rows, cols)
    assert np.array_equal(got.reshape(rows, cols), want)
    # the checksum is partition-invariant: a single band over the same final grid gives the same value
    sys.path.insert(0, PKG)
    from cgl_b200 import bands
    single = bands.RowBandLife(rows, cols, k=k, rank=0, world_size=1, device="cpu")
    single.set_owned(got_words)
    assert single.checksum() == int(torch.load(os.path.join(tmp_path, "checksum.pt"))[0])


def test_local_bands_emulation_matches_oracle():
    """G bands in one process (local copies instead of peers): the ghost-zone arithmetic alone."""
    sys.path.insert(0, PKG)
    from cgl_b200 import bands
    old = bands._life_block
    bands._life_block = _cpu_step_factory()
    try:
        rows, cols, k, gens = 32, 64, 4, 11
        cells, words = _grid(rows, cols, 7)
        lb = bands.LocalBands(rows, cols, k, 4, device="cpu")
        lb.set_grid(words)
        lb.run(gens)
        L = _twin()
        got = np.zeros(rows * cols, np.uint8)
        g = lb.grid().contiguous()
        L.twin_unpack(g.data_ptr(), got.ctypes.data, 1, rows, cols)
        assert np.array_equal(got.reshape(rows, cols), oracle.life(cells, gens))
    finally:
        bands._life_block = old
