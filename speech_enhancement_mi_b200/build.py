"""Build libse_b200.so (the C-ABI of include/se_b200.h) in-tree with nvcc for sm_100a.

    python -m speech_enhancement_mi_b200.build [--force] [-v]

nvcc cross-compiles without a GPU.  Every `.cu` is compiled to its own object (in parallel) and the objects are linked
into the shared library.  What is rebuilt is decided by CONTENT, not by file times: each object carries the SHA-256 of
its source, of every header of the library and of the compiler flags; the library carries the hash of its objects'
keys (`libse_b200.so.hash`).  A shipped library whose recorded hash does not match the sources in the tree is rebuilt.
The built files are git-ignored but travel with the working tree.
"""
from __future__ import annotations

import concurrent.futures
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libse_b200.so")
LIB_HASH = LIB + ".hash"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC",
]
LINK_LIBS: list = []  # the driver entry points (cuTensorMapEncodeTiled) are fetched with cudaGetDriverEntryPoint


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def _sha(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _object_key(src: str, hdr_hash: str) -> str:
    return _sha([src], hdr_hash + " ".join(NVCC_FLAGS))


def source_hash() -> str:
    """Hash of everything the library is built from (sources, headers, flags)."""
    hdr = _sha(headers())
    return hashlib.sha256("".join(_object_key(s, hdr) for s in sources()).encode() + " ".join(LINK_LIBS).encode()).hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(LIB_HASH):
        return True
    with open(LIB_HASH) as f:
        return f.read().strip() != source_hash()


def _compile(nvcc, src, obj, key, verbose):
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(obj + ".key", "w") as f:
        f.write(key)
    return res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr = _sha(headers())
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        key = _object_key(src, hdr)
        old = open(obj + ".key").read() if os.path.exists(obj + ".key") and os.path.exists(obj) else ""
        if force or old != key:
            jobs.append((src, obj, key))
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        futs = [ex.submit(_compile, nvcc, s, o, k, verbose) for s, o, k in jobs]
        for (s, _, _), f in zip(jobs, futs):
            log = f.result()
            if verbose:
                print(f"== {os.path.basename(s)}\n{log}")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + LINK_LIBS
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(LIB_HASH, "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
