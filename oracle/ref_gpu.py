"""TEST INFRASTRUCTURE ONLY -- launch the REFERENCE's own CUDA kernels on the GPU through cuda-python.

Only tests/, __graft_entry__.smoke() and bench.py's reference arms may import this module (the product never does:
tests/test_abi_and_host.py greps for it).

The cubins under tests/golden/ref_kernels/ were compiled by tests/golden/make_ref_cubins.py from the kernel
strings inside the reference files (/root/reference/CGL/CGL.py:146-182, CGL/CGL_action+/CGL.py:159-196), formatted
with the listed constants, exactly as pycuda's SourceModule would compile them.  `RefKernel.run` is the
reference's launch (`run_gpu(world, result, stable, side, size, block=(32*warp,1,1), grid=(ceil(size/block),1))`,
CGL/CGL.py:193-194,206) on device buffers; `RefGpuStep` adds the four blocking PCIe copies of
`__step_state_gpu` (CGL/CGL.py:203-208) around it: the reference's whole GPU step, for timing it as a baseline.
"""
from __future__ import annotations

import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_DIR = os.path.join(ROOT, "tests", "golden", "ref_kernels")


def _ok(res):
    err, *rest = res
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return rest[0] if len(rest) == 1 else (rest or None)


def manifest() -> dict:
    with open(os.path.join(KERNEL_DIR, "manifest.json")) as f:
        return json.load(f)


def cubin_path(variant: str, stable_max: int, spawn: int, empty_min: int | None = None) -> str:
    consts = (stable_max, spawn) if variant == "base" else (stable_max, empty_min, spawn)
    tag = "_".join(str(c).replace("-", "m") for c in consts)
    return os.path.join(KERNEL_DIR, f"{variant}_{tag}.cubin")


def available(variant: str, stable_max: int, spawn: int, empty_min: int | None = None) -> bool:
    return os.path.exists(cubin_path(variant, stable_max, spawn, empty_min))


class RefKernel:
    """One compiled instance of the reference's kernel `run` (constants are baked into the source, as in the
    reference).  variant = "base" (CGL/CGL.py) or "fork" (CGL_action+/CGL.py, the dead-cell decay rule)."""

    def __init__(self, variant: str, stable_max: int, spawn: int, empty_min: int | None = None, warp: int = 8):
        import torch
        from cuda.bindings import driver
        self.driver = driver
        torch.cuda.init()
        torch.zeros(1, device="cuda")                       # make sure the primary context is current
        path = cubin_path(variant, stable_max, spawn, empty_min)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run tests/golden/make_ref_cubins.py with these constants")
        with open(path, "rb") as f:
            self._image = np.frombuffer(f.read(), dtype=np.uint8).copy()
        self.module = _ok(driver.cuModuleLoadData(self._image.ctypes.data))
        self.func = _ok(driver.cuModuleGetFunction(self.module, b"run"))
        self.block = 32 * warp                              # CGL/CGL.py:124,193 (default warp = 8)

    def run(self, world_ptr: int, result_ptr: int, stable_ptr: int, side: int, stream: int = 0) -> None:
        """run<<<ceil(size / block), block, 0, stream>>>(world, result, stable, side, size)  (CGL/CGL.py:206)."""
        size = side * side
        grid = (size + self.block - 1) // self.block        # CGL/CGL.py:194
        args = (np.array([world_ptr], np.uint64), np.array([result_ptr], np.uint64), np.array([stable_ptr], np.uint64),
                np.array([side], np.uint32), np.array([size], np.uint32))
        ptrs = np.array([a.ctypes.data for a in args], dtype=np.uint64)
        _ok(self.driver.cuLaunchKernel(self.func, grid, 1, 1, self.block, 1, 1, 0, stream, ptrs.ctypes.data, 0))

    def step_tensors(self, world, result, stable, side: int) -> None:
        """Same on torch CUDA tensors (uint8 [size], uint8 [size], int8 [size]) on torch's current stream."""
        import torch
        self.run(world.data_ptr(), result.data_ptr(), stable.data_ptr(), side,
                 torch.cuda.current_stream().cuda_stream)


class RefGpuStep:
    """The reference's `__step_state_gpu` (CGL/CGL.py:203-208) restated call for call: three device buffers
    allocated once (:188-190), per step H2D world, H2D stable, launch on a private stream (:195), D2H world (from
    the result buffer), D2H stable -- blocking copies from/to the caller's pageable numpy arrays."""

    def __init__(self, side: int, stable_max: int, spawn: int, variant: str = "base", empty_min: int | None = None):
        self.k = RefKernel(variant, stable_max, spawn, empty_min)
        d = self.driver = self.k.driver
        self.side, self.size = side, side * side
        self.world_gpu = _ok(d.cuMemAlloc(self.size))
        self.stable_gpu = _ok(d.cuMemAlloc(self.size))
        self.result_gpu = _ok(d.cuMemAlloc(self.size))
        self.stream = _ok(d.cuStreamCreate(0))              # pycuda's cuda.Stream(): default flags (blocking)

    def step(self, world: np.ndarray, stable: np.ndarray) -> None:
        d = self.driver
        _ok(d.cuMemcpyHtoD(self.world_gpu, world.ctypes.data, self.size))
        _ok(d.cuMemcpyHtoD(self.stable_gpu, stable.ctypes.data, self.size))
        self.k.run(int(self.world_gpu), int(self.result_gpu), int(self.stable_gpu), self.side, int(self.stream))
        _ok(d.cuMemcpyDtoH(world.ctypes.data, self.result_gpu, self.size))
        _ok(d.cuMemcpyDtoH(stable.ctypes.data, self.stable_gpu, self.size))

    def close(self):
        d = self.driver
        for p in (self.world_gpu, self.stable_gpu, self.result_gpu):
            d.cuMemFree(p)
        d.cuStreamDestroy(self.stream)
