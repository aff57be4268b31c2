"""Build recipe for the CUDA extension (in-tree, sm_100a only).

``python -m dexterous_rl_manipulation_b200.build`` or ``build()``: one nvcc invocation that
produces ``dexterous_rl_manipulation_b200/libdexsim_b200.so`` next to this file.  nvcc
cross-compiles without a GPU.  ``-fmad=false`` is part of the numerics contract: the reference
rounds every product and sum separately (DESIGN.md "Numerics").
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "dexsim_kernels.cu")
HOST_SRC = os.path.join(HERE, "csrc", "dexsim_host_expand.cpp")     # plain C++ (host compiler): AVX2 expansion of packed contact rows
DEPS = [SRC, HOST_SRC] + [os.path.join(HERE, "csrc", h) for h in ("dexsim_core.cuh", "dexsim_step_tma.cuh", "dexsim_rollout_split.cuh")] \
    + [os.path.join(HERE, "..", "include", "dexsim.h")]
OUT = os.path.join(HERE, "libdexsim_b200.so")
INFO = os.path.join(HERE, "BUILD_INFO.json")       # written next to the library: source hash + the nvcc command line

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


def sources_sha256():
    """Hash of everything the library is compiled from (file names + contents, fixed order)."""
    h = hashlib.sha256()
    for d in DEPS:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


STEP_KERNEL_FILES = ("dexsim_step_tma.cuh", "dexsim_core.cuh")      # the device code of the pipelined step kernel


def step_kernel_sha256():
    """Hash of the step kernel's device code only (profiles/step_kernel_traffic.json is keyed by it: an ncu traffic figure
    stays quotable across edits of host-side code, and goes stale with any edit of the kernel itself)."""
    h = hashlib.sha256()
    for name in STEP_KERNEL_FILES:
        h.update(name.encode() + b"\0")
        with open(os.path.join(HERE, "csrc", name), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_info():
    """What the library next to this file was built from, and whether that is the tree's current source."""
    info = {}
    try:
        with open(INFO) as fh:
            info = json.load(fh)
    except (OSError, ValueError):
        pass
    try:
        info["tree_sha256"] = sources_sha256()
    except OSError:
        info["tree_sha256"] = None          # sources not shipped: nothing to compare with
    info["fresh"] = bool(info.get("sources_sha256")) and info.get("sources_sha256") == info["tree_sha256"]
    return info


def needs_build():
    if not os.path.exists(OUT):
        return True
    return not build_info()["fresh"]


def build(force=False, verbose=False, out=None, defines=()):
    """``out`` / ``defines``: experiment builds (``-DNAME`` variants written somewhere else, loaded through
    DEXSIM_LIB_PATH by the timing tools); the product is the default call."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC, HOST_SRC]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libdexsim_b200.so")
    if out == OUT:
        with open(INFO, "w") as fh:
            json.dump({"sources_sha256": sources_sha256(), "step_kernel_sha256": step_kernel_sha256(), "nvcc": " ".join(cmd[:1] + [c for c in cmd[1:] if c not in (out, SRC, HOST_SRC)]),
                       "flags": NVCC_FLAGS, "defines": list(defines)}, fh, indent=1)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
