"""Build recipe for the CUDA extension (in-tree, sm_100a only).

``python -m dexterous_rl_manipulation_b200.build`` or ``build()``: one nvcc invocation that
produces ``dexterous_rl_manipulation_b200/libdexsim_b200.so`` next to this file.  nvcc
cross-compiles without a GPU.  ``-fmad=false`` is part of the numerics contract: the reference
rounds every product and sum separately (DESIGN.md "Numerics").
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "dexsim_kernels.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", h) for h in ("dexsim_core.cuh", "dexsim_step_tma.cuh", "dexsim_rollout_split.cuh")] \
    + [os.path.join(HERE, "..", "include", "dexsim.h")]
OUT = os.path.join(HERE, "libdexsim_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-cudart", "static",
    "-shared",
]


def nvcc_path():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC=...)")
    return cand


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libdexsim_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
