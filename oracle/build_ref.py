"""Byte-compile the UNMODIFIED reference into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

The reference is pure Python, and /root/reference does not exist on the GPU box.  This
recipe compiles each module from the sources where they lie (read-only) into sourceless
byte-code files (``*.refbc``) under ``oracle/_ref/`` (git-ignored, shipped with the gpurun snapshot like a
built ``.so``).  No reference source text is copied into the repository.  ``bench.py --impl
reference`` and ``cpu_baseline`` import it (under oracle/gym_stub) to time the reference's
own Python loop on the GPU box's host cores; tests use it to re-validate goldens when present.
"""
import argparse
import os
import py_compile
import shutil
import sys

# gpurun's snapshot drops *.pyc, so the byte-code files carry their own suffix; oracle/ref_harness.py
# installs a finder that loads them with importlib's SourcelessFileLoader.
BC_SUFFIX = ".refbc"
PACKAGES = ["envs", "rewards", "policies", "experiments", "training", "evaluation"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))
    args = ap.parse_args()
    if not os.path.isdir(args.src):
        print(f"build_ref: {args.src} not present; nothing to do")
        return 0
    if os.path.isdir(args.out):
        shutil.rmtree(args.out)
    n = 0
    for pkg in PACKAGES:
        src_pkg = os.path.join(args.src, pkg)
        if not os.path.isdir(src_pkg):
            continue
        for root, _dirs, files in os.walk(src_pkg):
            rel = os.path.relpath(root, args.src)
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                dst_dir = os.path.join(args.out, rel)
                os.makedirs(dst_dir, exist_ok=True)
                py_compile.compile(os.path.join(root, f), cfile=os.path.join(dst_dir, f[:-3] + BC_SUFFIX),
                                   dfile=os.path.join("<reference>", rel, f), doraise=True,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
                n += 1
    with open(os.path.join(args.out, "BUILD_INFO.txt"), "w") as fh:
        fh.write(f"byte-compiled from {args.src} by oracle/build_ref.py with Python {sys.version.split()[0]}; {n} modules\n")
    print(f"build_ref: compiled {n} modules into {args.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
