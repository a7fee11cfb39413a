#!/usr/bin/env python
"""Compile the reference's OWN CUDA kernels for sm_100a, so that the GPU box can execute them.

    python tests/golden/make_ref_cubins.py            # needs /root/reference and cuda-python (NVRTC); no GPU

The reference's device path is a CUDA-C string JIT-compiled by PyCUDA at construction:
  base env   /root/reference/CGL/CGL.py:146-182              `.format(stableStabilityFactor, spawnStabilityFactor)`
  fork env   /root/reference/CGL/CGL_action+/CGL.py:159-196  `.format(stableStabilityFactor, empty_min, spawnStabilityFactor)`
PyCUDA is not installed here and cannot be (no network), and /root/reference does not exist on the GPU box.  This
script extracts the two strings from the reference files as they lie under /root/reference (nothing is copied into
the repository as source), formats them with the constant sets the tests and the bench use, wraps them in
`extern "C" { }` exactly as pycuda.compiler.SourceModule does by default, and compiles each with NVRTC to a cubin
under tests/golden/ref_kernels/ (binary fixtures + manifest.json with the SHA-256 of every formatted source).

The cubins are TEST INFRASTRUCTURE: oracle/ref_gpu.py loads and launches them through cuda-python so that
  * tests/test_gpu_ref_kernel.py can compare our kernels with the reference's kernel executing on the same B200,
    bit for bit -- this is what pins the fork's CUDA-kernel rule ("decay"), which the fork's CPU step contradicts;
  * bench.py --impl reference-gpu can time the reference's GPU step (kernel + its four PCIe copies,
    CGL/CGL.py:203-208) beside ours.
"""
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_kernels")
REF = os.environ.get("CGL_REFERENCE", "/root/reference")
SOURCES = {"base": os.path.join(REF, "CGL", "CGL.py"), "fork": os.path.join(REF, "CGL", "CGL_action+", "CGL.py")}

# (stable_max, spawn) for the base kernel; (stable_max, empty_min, spawn) for the fork's -- in .format() order
BASE_SETS = [(1, -1), (2, -2), (3, -2), (4, -3), (127, -128), (2, 5)]
FORK_SETS = [(2, -128, -2), (2, -5, -2), (2, -4, -2), (2, -3, -2), (2, -6, -2), (2, 7, -2), (3, -6, -2), (3, -6, -1),
             (3, -6, 0), (4, -90, -3), (1, -128, -1), (127, -128, -128)]


def kernel_template(path):
    text = open(path).read()
    m = re.search(r'SourceModule\("""(.*?)"""\s*\.format\(', text, re.S)
    if not m:
        raise SystemExit(f"no SourceModule string found in {path}")
    return m.group(1)


def cubin_name(variant, consts):
    tag = "_".join(str(c).replace("-", "m") for c in consts)
    return f"{variant}_{tag}.cubin"


def compile_cubin(src: str, name: str) -> bytes:
    from cuda.bindings import nvrtc

    def ok(res):
        err, *rest = res
        if int(err) != 0:
            raise RuntimeError(f"NVRTC error {err}")
        return rest[0] if len(rest) == 1 else rest

    prog = ok(nvrtc.nvrtcCreateProgram(src.encode(), name.encode(), 0, [], []))
    opts = [b"--gpu-architecture=sm_100a"]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(err) != 0:
        n = ok(nvrtc.nvrtcGetProgramLogSize(prog))
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError(f"NVRTC failed for {name}:\n{log.decode(errors='replace')}")
    n = ok(nvrtc.nvrtcGetCUBINSize(prog))
    buf = b" " * n
    ok(nvrtc.nvrtcGetCUBIN(prog, buf) + (None,))
    return buf


def main():
    from cuda.bindings import nvrtc
    os.makedirs(OUT, exist_ok=True)
    manifest = {"nvrtc": list(int(v) for v in nvrtc.nvrtcVersion()[1:]), "arch": "sm_100a", "entry": "run",
                "wrap": 'extern "C" { ... }  (pycuda.compiler.SourceModule default, no_extern_c=False)', "kernels": {}}
    for variant, sets in (("base", BASE_SETS), ("fork", FORK_SETS)):
        tmpl = kernel_template(SOURCES[variant])
        for consts in sets:
            src = 'extern "C" {\n' + tmpl.format(*consts) + "\n}\n"
            name = cubin_name(variant, consts)
            cubin = compile_cubin(src, name.replace(".cubin", ".cu"))
            with open(os.path.join(OUT, name), "wb") as f:
                f.write(cubin)
            keys = ("stable_max", "spawn") if variant == "base" else ("stable_max", "empty_min", "spawn")
            manifest["kernels"][name] = {"variant": variant, **dict(zip(keys, consts)),
                                         "source": os.path.relpath(SOURCES[variant], REF),
                                         "source_sha256": hashlib.sha256(src.encode()).hexdigest(), "bytes": len(cubin)}
            print(name, len(cubin), "bytes")
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
