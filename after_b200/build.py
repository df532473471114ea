"""Build libafter_b200.so (sm_100a only) in-tree with nvcc.

    python -m after_b200.build [--force] [--debug]

The library lands in ``after_b200/lib/`` (git-ignored, shipped to the GPU box by gpurun).
There is a single translation unit (``csrc/api.cu``) on purpose: the whole build is one nvcc
invocation of a few seconds and needs no build system.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libafter_b200.so")
STAMP = os.path.join(LIBDIR, "libafter_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC",
]


def _sources():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "after_b200.h"))
    return files


def _digest(debug: bool = False):
    h = hashlib.sha256()
    for f in _sources():
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + (["-DAFTER_DEBUG"] if debug else [])).encode())
    return h.hexdigest()


def is_fresh(debug: bool = False) -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest(debug)


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """Compile the library if sources changed; returns its path.  ``debug`` adds -DAFTER_DEBUG: the A/B environment
    knobs (AFTER_ATTN, AFTER_PDL, AFTER_NO_GRAPH, AFTER_MLP_KSPLIT, ...) and the in-kernel %globaltimer traces exist
    only in that build; the release library ignores the environment."""
    debug = debug or os.environ.get("AFTER_B200_DEBUG_BUILD") == "1"
    if not force and is_fresh(debug):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libafter_b200.so")
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-DAFTER_DEBUG"] if debug else []) + (["-Xptxas", "-v"] if verbose else []) + [
        os.path.join(CSRC, "api.cu"), "-o", LIB, "-lcuda"
    ]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    with open(STAMP, "w") as fh:
        fh.write(_digest(debug))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv)
    print(path)
