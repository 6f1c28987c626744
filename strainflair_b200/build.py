"""Builds libsfb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "match.cu", "steps.cu", "group.cu", "scatter.cu", "tile.cu", "tilesets.cu", "reduce.cu", "expand.cu", "decode.cpp", "gfa.cpp", "wire.cpp", "unassigned.cpp", "locality.cpp")]
import glob
HEADERS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
LIB = os.path.join(HERE, "libsfb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-pthread", "-shared", "-cudart", "shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise FileNotFoundError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"), "-o", LIB] + SOURCES + ["-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
