"""RowBandLife -- one huge toroidal grid, world plane only ("life mode"), sharded by ROW BANDS.

The reference cannot express these sizes (`unsigned int` index math overflows above side 46340,
/root/reference/CGL/CGL.py:149-159) and is single-GPU (:6); this is the B200-native extension
BASELINE.json configs[3]/[4] name: 65536^2 x 1000 generations on 1/2/4/8 GPUs.

Rank r of G owns rows [r*H/G, (r+1)*H/G) of the H x C torus and keeps k ghost rows above and
below them.  Every k generations each rank sends its top k and bottom k owned rows to its ring
neighbours (rank 0's upper neighbour is rank G-1: that IS the vertical torus wrap), then runs k
generations over the whole (H/G + 2k)-row buffer with dead rows outside: the garbage that creeps
in from the buffer edges advances one row per generation and never reaches an owned row.
Horizontal wrap is local.  Message = k * C/8 bytes per neighbour per k generations.

Exchange back ends
  "persist" ONE cooperative launch per run() call (`cgl_life_band_run`): a warp keeps its strip of
          rows for all sub-steps of kernel_k generations and waits only for its neighbour strips; every k generations
          the strips that own a rank's first / last k owned rows store them into the ring neighbours' landing zones
          over NVLink (CUDA-IPC mapped peer memory, two slots) and bump an arrival counter, and the strips that read
          ghost rows wait for the counter and copy what they need out of their own landing zone.  No exchange
          launch, no launch boundary between sub-steps, interior strips never wait for another GPU.  Correct and
          tested, but measured slower than "p2p" (the persistent kernel's row loop runs at 0.88 us per row step
          against 0.60: DESIGN.md section 4.6), so it is opt-in.
  "fused" the halo exchange lives INSIDE the k-generation kernel
          (`cgl_life_band_block`): the strips that produce a rank's first/last k owned rows store
          them straight into the neighbours' next input buffers over NVLink (CUDA-IPC mapped peer
          memory) and bump an arrival counter; only the strips that read ghost rows wait for the
          neighbours' previous block.  No exchange launches, no host sync, interior strips never wait.
  "p2p"   (default on GPUs) one exchange launch per block (`cgl_halo_exchange`, 8 CTAs per neighbour): writes the strip
          straight into the neighbour's landing zone and publishes a sequence flag;
          `cgl_halo_wait_copy` on the neighbour spins on the flag and moves the strip into its
          ghost rows.  Two landing slots alternate by block parity, so a rank may run one block
          ahead of its neighbour without a handshake.  No host synchronisation per exchange.
  "dist"  torch.distributed batch_isend_irecv (NCCL on GPUs, gloo in the CPU tests).
  "local" all G bands live in this process (one GPU emulating G ranks; tests the ghost-zone math).
"""
from __future__ import annotations

import ctypes

import torch

from . import native


def _life_block(lib, a, b, rows, cols, wrap_rows, gens, k, stream) -> bool:
    """`gens` generations starting from buffer `a`, ping-ponging with `b`, k generations per launch
    (temporal blocking, cgl_life_run).  Returns True if the result is in `a`.
    Tests monkeypatch this for CPU runs."""
    res = ctypes.c_int(-1)
    native.check(lib.cgl_life_run(native.dptr(a), native.dptr(b), rows, cols, wrap_rows, gens, k,
                                  ctypes.byref(res), stream), "cgl_life_run")
    return res.value == 1


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class RowBandLife:
    def __init__(self, rows: int, cols: int, k: int = 8, rank: int = 0, world_size: int = 1, device="cuda",
                 exchange: str | None = None, group=None, lib=None, kernel_k: int | None = None):
        """k = ghost depth = generations between halo exchanges; kernel_k = generations per launch
        (temporal-blocking depth, default min(k, 8)): k = 16, kernel_k = 8 exchanges every 16
        generations and runs two 8-generation launches in between."""
        if rows % world_size:
            raise ValueError("rows must be divisible by world_size")
        if cols % 32:
            raise ValueError("cols must be a multiple of 32 for row bands")
        self.rows, self.cols, self.k = rows, cols, int(k)
        self.kernel_k = int(kernel_k) if kernel_k else min(self.k, 8)
        self.rank, self.G = rank, world_size
        self.W = cols // 32
        self.band_rows = rows // world_size
        if self.G > 1 and self.band_rows < self.k:
            raise ValueError("band must have at least k rows")
        self.device = torch.device(device)
        self.group = group
        self.lib = lib if lib is not None else native.load()
        if exchange is None:
            exchange = "single" if world_size == 1 else ("p2p" if self.device.type == "cuda" else "dist")
        self.exchange = exchange
        # a ring of ONE with exchange="persist" keeps its ghost rows and exchanges with itself (its upper and lower
        # neighbour are the rank itself: the torus wrap) -- the in-kernel exchange on a single GPU (tests)
        self._self_ring = self.G == 1 and self.exchange == "persist"
        g = self.k if (self.G > 1 or self._self_ring) else 0
        self.ghost = g
        self.buf_rows = self.band_rows + 2 * g
        n = self.buf_rows * self.W
        self.generation = 0
        self.block = 0              # exchanges done so far
        self.launches = 0
        self._ipc = None
        self._fused = None
        self._ghosts_valid = False
        if self.exchange == "fused" and self.G > 1:
            self._setup_fused(n)    # band buffers must be whole cudaMalloc allocations (IPC)
        else:
            self._a = torch.zeros(n, dtype=torch.int32, device=self.device)
            self._b = torch.zeros(n, dtype=torch.int32, device=self.device)
        if self.exchange in ("p2p", "persist") and (self.G > 1 or self._self_ring):
            self._setup_p2p()

    # ---------------------------------------------------------------- state access
    def _stream(self):
        if self.device.type != "cuda":
            return None
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def owned(self) -> torch.Tensor:
        """View of this rank's owned rows: int32 [band_rows, W] (bit-packed uint32 words)."""
        g, W = self.ghost, self.W
        return self._a[g * W:(g + self.band_rows) * W].view(self.band_rows, W)

    def set_owned(self, words: torch.Tensor) -> None:
        self.owned.copy_(words.view(self.band_rows, self.W))
        self._ghosts_valid = False

    def randomize(self, seed: int = 0) -> None:
        """Synthetic Bernoulli(0.5) start: every packed word uniform, seeded by GLOBAL row so that the
        grid does not depend on the number of bands."""
        g = torch.Generator(device=self.device)
        chunk = 256 if self.band_rows % 256 == 0 else 1      # chunk starts are the same for every G
        r0 = self.rank * self.band_rows
        for s in range(0, self.band_rows, chunk):
            g.manual_seed(seed * 1000003 + (r0 + s))
            self.owned[s:s + chunk] = torch.randint(-2 ** 31, 2 ** 31 - 1, (chunk, self.W), dtype=torch.int32,
                                                    device=self.device, generator=g)
        self._ghosts_valid = False

    # ---------------------------------------------------------------- fused-exchange plumbing
    def _alloc_shared(self, nbytes):
        """cudaMalloc'ed (IPC-exportable) memory + its handle."""
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        with torch.cuda.device(self.device):
            native.check(self.lib.cgl_dev_alloc(nbytes, ctypes.byref(ptr)), "cgl_dev_alloc")
            native.check(self.lib.cgl_ipc_get_handle(ptr, handle), "cgl_ipc_get_handle")
        return ptr.value, bytes(handle)

    def _as_tensor(self, ptr, n_words):
        class _Raw:     # zero-copy int32 view of raw device memory
            __cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Raw(), device=self.device)

    def _setup_fused(self, n_words):
        import torch.distributed as dist
        mine = [self._alloc_shared(n_words * 4), self._alloc_shared(n_words * 4), self._alloc_shared(256)]
        self._a = self._as_tensor(mine[0][0], n_words)
        self._b = self._as_tensor(mine[1][0], n_words)
        handles = [None] * self.G
        dist.all_gather_object(handles, [h for _, h in mine], group=self.group)
        up, down = (self.rank - 1) % self.G, (self.rank + 1) % self.G
        peers = {}
        with torch.cuda.device(self.device):
            for p in {up, down}:
                ptrs = []
                for h in handles[p]:
                    ptr = ctypes.c_void_p()
                    native.check(self.lib.cgl_ipc_open_handle((ctypes.c_uint8 * 64).from_buffer_copy(h), ctypes.byref(ptr)),
                                 "cgl_ipc_open_handle")
                    ptrs.append(ptr.value)
                peers[p] = ptrs
        self._fused = dict(mine=[m[0] for m in mine], up=peers[up], down=peers[down], peers=peers, block_index=0)
        dist.barrier(group=self.group)

    # ---------------------------------------------------------------- P2P plumbing
    def _setup_p2p(self):
        """Allocate landing zones + flags with plain cudaMalloc, exchange IPC handles, map the neighbours'."""
        import torch.distributed as dist
        lib, k, W = self.lib, self.k, self.W
        strip = k * W * 4
        # layout: [slot0: from_up | from_down][slot1: from_up | from_down][flags: 4 x uint32 (+pad to 16)]
        total = 4 * strip + 64
        base = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            native.check(lib.cgl_dev_alloc(total, ctypes.byref(base)), "cgl_dev_alloc")
            handle = (ctypes.c_uint8 * 64)()
            native.check(lib.cgl_ipc_get_handle(base, handle), "cgl_ipc_get_handle")
        if self.G == 1:                                     # ring of one: my neighbours are me
            self._ipc = dict(base=base.value, strip=strip, up=base.value, down=base.value, flags_off=4 * strip, local=True)
            return
        handles = [None] * self.G
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        up, down = (self.rank - 1) % self.G, (self.rank + 1) % self.G
        peers = {}
        with torch.cuda.device(self.device):
            for p in {up, down}:
                ptr = ctypes.c_void_p()
                h = (ctypes.c_uint8 * 64).from_buffer_copy(handles[p])
                native.check(lib.cgl_ipc_open_handle(h, ctypes.byref(ptr)), "cgl_ipc_open_handle")
                peers[p] = ptr.value
        self._ipc = dict(base=base.value, strip=strip, up=peers[up], down=peers[down], flags_off=4 * strip)
        dist.barrier(group=self.group)

    def close(self):
        if self._fused:
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            f = self._fused
            self._a = self._b = None
            with torch.cuda.device(self.device):
                for ptrs in f["peers"].values():
                    for p in ptrs:
                        self.lib.cgl_ipc_close_handle(ctypes.c_void_p(p))
                dist.barrier(group=self.group)
                for p in f["mine"]:
                    self.lib.cgl_dev_free(ctypes.c_void_p(p))
            self._fused = None
        if self._ipc:
            torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):
                if not self._ipc.get("local"):
                    import torch.distributed as dist
                    dist.barrier(group=self.group)
                    for p in {self._ipc["up"], self._ipc["down"]}:
                        self.lib.cgl_ipc_close_handle(ctypes.c_void_p(p))
                self.lib.cgl_dev_free(ctypes.c_void_p(self._ipc["base"]))
            self._ipc = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            if self._ipc and self._ipc.get("local"):        # (peer mappings need the collective close())
                self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---------------------------------------------------------------- halo exchange
    def _exchange(self):
        """Fill the k ghost rows above/below the owned rows of the current buffer."""
        k, W, g = self.k, self.W, self.ghost
        if self.G == 1:
            return
        a = self._a
        top_owned = a[g * W:(g + k) * W]                                     # my first k owned rows
        bot_owned = a[(g + self.band_rows - k) * W:(g + self.band_rows) * W]  # my last k owned rows
        ghost_up = a[0:k * W]
        ghost_down = a[(g + self.band_rows) * W:(g + self.band_rows + k) * W]
        self.block += 1
        if self.exchange in ("dist", "fused"):      # fused: only to (re)fill the ghosts of a fresh grid
            import torch.distributed as dist
            up, down = (self.rank - 1) % self.G, (self.rank + 1) % self.G
            send_up, send_down = top_owned.clone(), bot_owned.clone()
            ops = [dist.P2POp(dist.isend, send_up, up, group=self.group, tag=0),
                   dist.P2POp(dist.isend, send_down, down, group=self.group, tag=1),
                   dist.P2POp(dist.irecv, ghost_down, down, group=self.group, tag=0),   # neighbour's top rows
                   dist.P2POp(dist.irecv, ghost_up, up, group=self.group, tag=1)]       # neighbour's bottom rows
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            return
        if self.exchange == "p2p":
            lib, ipc, st = self.lib, self._ipc, self._stream()
            strip, slot, seq = ipc["strip"], self.block & 1, self.block
            n_words = k * W
            # my top rows are the neighbour-above's "from_down" strip; my bottom rows the neighbour-below's "from_up"
            base, fo = ipc["base"], ipc["flags_off"]
            V = ctypes.c_void_p
            with torch.cuda.device(self.device):
                native.check(lib.cgl_halo_exchange(
                    native.dptr(top_owned), native.dptr(bot_owned),
                    V(ipc["up"] + (2 * slot + 1) * strip), V(ipc["down"] + (2 * slot) * strip),
                    V(ipc["up"] + fo + 4 * (2 * slot + 1)), V(ipc["down"] + fo + 4 * (2 * slot)),
                    V(base + (2 * slot) * strip), V(base + (2 * slot + 1) * strip),
                    V(base + fo + 4 * (2 * slot)), V(base + fo + 4 * (2 * slot + 1)),
                    native.dptr(ghost_up), native.dptr(ghost_down), n_words, seq, st), "cgl_halo_exchange")
            self.launches += 1
            return
        raise ValueError(f"unknown exchange {self.exchange!r}")

    # ---------------------------------------------------------------- stepping
    def run(self, gens: int) -> None:
        """Advance `gens` generations (blocks of k generations between exchanges)."""
        lib, st = self.lib, self._stream()
        if self.device.type == "cuda" and self.exchange not in ("fused", "persist") and not getattr(self, "_tuned", False):
            # measure the strip length of the k-blocked kernel for this band shape once (clobbers _b only)
            with torch.cuda.device(self.device):
                native.check(lib.cgl_life_tune(native.dptr(self._a), native.dptr(self._b),
                                               self.buf_rows, self.cols, 1 if self.ghost == 0 else 0, self.kernel_k,
                                               st), "cgl_life_tune")
            self._tuned = True
        if self.G == 1 and not self._self_ring:   # plain torus: one call, k generations per launch
            with torch.cuda.device(self.device) if self.device.type == "cuda" else _null():
                if not _life_block(lib, self._a, self._b, self.band_rows, self.cols, 1, gens, self.kernel_k, st):
                    self._a, self._b = self._b, self._a
            self.launches += -(-gens // self.kernel_k)
            self.generation += gens
            return
        if self.exchange == "fused":
            return self._run_fused(gens)
        if self.exchange == "persist":
            return self._run_persist(gens)
        done = 0
        while done < gens:
            kb = min(self.k, gens - done)
            self._exchange()
            with torch.cuda.device(self.device) if self.device.type == "cuda" else _null():
                if not _life_block(lib, self._a, self._b, self.buf_rows, self.cols, 0, kb, min(kb, self.kernel_k), st):
                    self._a, self._b = self._b, self._a
            self.launches += -(-kb // self.kernel_k)
            done += kb
            self.generation += kb

    def _run_persist(self, gens: int) -> None:
        lib, ipc, st, kk = self.lib, self._ipc, self._stream(), self.kernel_k

        def barrier():
            if self.G > 1:
                import torch.distributed as dist
                dist.barrier(group=self.group)
        if gens % kk:
            raise ValueError(f"exchange='persist' runs whole sub-steps: gens must be a multiple of kernel_k = {kk}")
        V = ctypes.c_void_p
        if not self._ghosts_valid:
            # fresh grid: zero the arrival counters on every rank, then the first call pushes the edge rows itself
            torch.cuda.synchronize(self.device)
            barrier()
            with torch.cuda.device(self.device):
                native.check(lib.cgl_dev_memset(V(ipc["base"] + ipc["flags_off"]), 0, 64, st), "cgl_dev_memset")
            torch.cuda.synchronize(self.device)
            barrier()
            self.block = 0
        strip, base, fo = ipc["strip"], ipc["base"], ipc["flags_off"]
        up, down = ipc["up"], ipc["down"]
        arr = V * 2
        # my first owned rows -> the upper neighbour's "from below" zone; my last owned rows -> the lower one's "from above"
        peer_up = arr(up + strip, up + 3 * strip)
        peer_dn = arr(down, down + 2 * strip)
        mine_up = arr(base, base + 2 * strip)
        mine_dn = arr(base + strip, base + 3 * strip)
        n_sub = gens // kk
        with torch.cuda.device(self.device):
            native.check(lib.cgl_life_band_run(
                native.dptr(self._a), native.dptr(self._b), self.buf_rows, self.cols, self.ghost, kk, n_sub, self.block,
                int(not self._ghosts_valid), peer_up, peer_dn, V(up + fo + 4), V(down + fo), mine_up, mine_dn,
                V(base + fo), V(base + fo + 4), st), "cgl_life_band_run")
        self._ghosts_valid = True
        per_block = self.k // kk
        self.block += -(-n_sub // per_block)
        if n_sub & 1:
            self._a, self._b = self._b, self._a
        self.launches += 1
        self.generation += gens

    _BLOCK_SIZES = (16, 12, 8, 6, 4, 3, 2, 1)       # generations one fused launch can do

    def _run_fused(self, gens: int) -> None:
        import torch.distributed as dist
        f, lib, st = self._fused, self.lib, self._stream()
        if not self._ghosts_valid:
            # fresh grid: fill the ghost rows once with a plain exchange and restart the counters
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self._exchange()
            torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):
                native.check(lib.cgl_dev_memset(ctypes.c_void_p(f["mine"][2]), 0, 256, st), "cgl_dev_memset")
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            f["block_index"] = 0
            self._ghosts_valid = True
        done = 0
        while done < gens:
            kb = next(s for s in self._BLOCK_SIZES if s <= min(self.k, gens - done))
            f["block_index"] += 1
            cur_is_a = self._a.data_ptr() == f["mine"][0]
            i_in, i_out = (0, 1) if cur_is_a else (1, 0)
            with torch.cuda.device(self.device):
                native.check(lib.cgl_life_band_block(
                    ctypes.c_void_p(f["mine"][i_in]), ctypes.c_void_p(f["mine"][i_out]), self.buf_rows, self.cols,
                    self.ghost, kb, ctypes.c_void_p(f["up"][i_out]), ctypes.c_void_p(f["down"][i_out]),
                    ctypes.c_void_p(f["up"][2]), ctypes.c_void_p(f["down"][2]), ctypes.c_void_p(f["mine"][2]),
                    f["block_index"], st), "cgl_life_band_block")
            self._a, self._b = self._b, self._a
            self.launches += 1
            done += kb
            self.generation += kb

    # ---------------------------------------------------------------- global reductions
    def alive(self) -> int:
        """Global live-cell count (popcount of every rank's owned rows, summed over ranks)."""
        out = torch.zeros(1, dtype=torch.int32, device=self.device)
        owned = self.owned.contiguous()
        with torch.cuda.device(self.device):
            native.check(self.lib.cgl_alive(native.dptr(owned), 1, owned.numel(), native.dptr(out), self._stream()),
                         "cgl_alive")
        total = (out.to(torch.int64) & 0xFFFFFFFF)
        if self.G > 1:
            import torch.distributed as dist
            dist.all_reduce(total, group=self.group)
        n = int(total.item())
        if self.device.type == "cuda":
            native.check_alarm()                            # a device-side wait that gave up (see cgl_alarm_words)
        return n

    def checksum(self) -> int:
        """Order-independent 64-bit hash of the whole grid (sum over words of word * f(global index)),
        equal for any number of bands holding the same grid."""
        owned = self.owned.reshape(-1).to(torch.int64) & 0xFFFFFFFF
        r0 = self.rank * self.band_rows
        idx = torch.arange(owned.numel(), device=self.device, dtype=torch.int64) + r0 * self.W
        mix = (idx * 0x9E3779B1 + 0x7F4A7C15) & 0x7FFFFFFF
        total = ((owned * mix) & 0x7FFFFFFFFFFF).sum().reshape(1)
        if self.G > 1:
            import torch.distributed as dist
            dist.all_reduce(total, group=self.group)
        n = int(total.item())
        if self.device.type == "cuda":
            native.check_alarm()
        return n


class LocalBands:
    """G row bands held by ONE process/GPU, exchanged by local copies: same ghost-zone schedule as the
    multi-GPU path without peers (tests; also how a 1-GPU box checks the N-rank arithmetic)."""

    def __init__(self, rows, cols, k, n_bands, device="cuda", lib=None):
        self.bands = [RowBandLife(rows, cols, k=k, rank=r, world_size=n_bands, device=device, exchange="local", lib=lib)
                      for r in range(n_bands)]
        self.G, self.k = n_bands, k

    def set_grid(self, words: torch.Tensor) -> None:
        for b in self.bands:
            r0 = b.rank * b.band_rows
            b.set_owned(words.view(-1, b.W)[r0:r0 + b.band_rows])

    def grid(self) -> torch.Tensor:
        return torch.cat([b.owned for b in self.bands], dim=0)

    def run(self, gens: int) -> None:
        done = 0
        while done < gens:
            kb = min(self.k, gens - done)
            tops = [b._a[b.ghost * b.W:(b.ghost + b.k) * b.W].clone() for b in self.bands]
            bots = [b._a[(b.ghost + b.band_rows - b.k) * b.W:(b.ghost + b.band_rows) * b.W].clone() for b in self.bands]
            for r, b in enumerate(self.bands):
                up, down = (r - 1) % self.G, (r + 1) % self.G
                b._a[0:b.k * b.W] = bots[up]
                b._a[(b.ghost + b.band_rows) * b.W:(b.ghost + b.band_rows + b.k) * b.W] = tops[down]
            for b in self.bands:
                if not _life_block(b.lib, b._a, b._b, b.buf_rows, b.cols, 0, kb, kb, b._stream()):
                    b._a, b._b = b._b, b._a
                b.generation += kb
            done += kb
