"""GPU suite: row-band life mode.  One GPU: G emulated bands == plain torus == oracle.  With >= 2 GPUs:
torchrun over NVLink with both exchange back ends, checked against the single-GPU torus."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bands():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cgl_b200 import bands as b
    return b


def _pack_np(cells):
    rows, cols = cells.shape
    bits = cells.reshape(rows, cols // 32, 32).astype(np.uint64)
    words = (bits << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)
    return torch.from_numpy(words.view(np.int32))


def _unpack_np(words, rows, cols):
    w = words.cpu().numpy().view(np.uint32).reshape(rows, cols // 32)
    return ((w[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8).reshape(rows, cols)


@pytest.mark.parametrize("rows,cols,k,G,gens", [(256, 4096, 8, 4, 21), (96, 2048, 4, 2, 9), (64, 128, 8, 8, 16)])
def test_emulated_bands_equal_torus_and_oracle(bands, rows, cols, k, G, gens):
    cells = np.random.RandomState(rows + cols).randint(2, size=(rows, cols)).astype(np.uint8)
    words = _pack_np(cells).cuda()
    single = bands.RowBandLife(rows, cols, k=k)
    single.set_owned(words)
    single.run(gens)
    lb = bands.LocalBands(rows, cols, k, G)
    lb.set_grid(words)
    lb.run(gens)
    want = oracle.life(cells, gens, threads=4)
    assert np.array_equal(_unpack_np(single.owned, rows, cols), want)
    assert np.array_equal(_unpack_np(lb.grid(), rows, cols), want)
    assert single.alive() == int(want.sum())


def test_randomize_is_partition_invariant(bands):
    a = bands.RowBandLife(1024, 2048, k=8)
    a.randomize(3)
    parts = []
    for r in range(4):
        b = bands.RowBandLife(1024, 2048, k=8, rank=r, world_size=4, exchange="local")
        b.randomize(3)
        parts.append(b.owned.clone())
    assert torch.equal(torch.cat(parts, 0), a.owned)


@pytest.mark.parametrize("exchange", ["fused", "p2p", "dist"])
def test_two_gpu_bands_match_single_gpu(bands, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "run_bands.py"), "--rows", "4096", "--cols",
           "8192", "--k", "8", "--gens", "43", "--exchange", exchange, "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout
