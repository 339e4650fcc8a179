"""Builds libconstriction_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library travels to the GPU box with the repo snapshot (built `.so` files are git-ignored but not
gpurun-ignored).  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only
  -fmad=false / -ffp-contract=off           model tabulation must not fuse a*b+c (the reference is
                                            Rust, which never contracts); the integer coder kernels
                                            are unaffected
  -lineinfo                                 so that ncu's source page maps to the .cuh files
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libconstriction_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
]
# one translation unit per kernel family (they compile in parallel); capi.cu holds the C ABI
UNITS = ["capi", "container", "gather", "host_pipeline", "small_kernels", "chain_encode", "chain_decode", "ans_encode", "ans_decode", "range_encode", "range_decode"]
OBJ_DIR = os.path.join(HERE, "_obj")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [
        os.path.join(os.path.dirname(HERE), "include", "constriction_b200.h")]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libconstriction_b200.so")
    extra = os.environ.get("CTR_EXTRA_NVCC_FLAGS", "").split()  # experiments only (e.g. -DCTR_PF_BATCHES=8)
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_unit(unit):
        obj = os.path.join(OBJ_DIR, unit + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, unit + ".cu")]
        return unit, obj, subprocess.run(cmd, capture_output=True, text=True)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
        results = list(pool.map(compile_unit, UNITS))
    log = ""
    for unit, _, proc in results:
        log += f"==== {unit}.cu\n{proc.stderr}"
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError(f"nvcc failed compiling {unit}.cu")
    link = [nvcc, "--shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [
        obj for _, obj, _ in results]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed linking libconstriction_b200.so")
    if verbose:
        sys.stderr.write(log)
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:  # register / spill report (compile times vary: dropped)
        f.write("".join(line for line in log.splitlines(keepends=True) if "Compile time" not in line))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
