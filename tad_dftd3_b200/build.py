"""
Build `libd3b200.so` in-tree with nvcc for sm_100a.

    python -m tad_dftd3_b200.build [--force]

The shared library is a plain C-ABI library (no torch, no Python headers):
`include/d3b200.h` declares its entry points.  It is git-ignored but travels
to the GPU box with the working tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libd3b200.so")
STAMP = os.path.join(HERE, ".libd3b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xptxas=-v",            # register / spill report -> ptxas_info.log (no fast-math: fp64 parity)
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))
    ) + [os.path.join(os.path.dirname(HERE), "include", "d3b200.h")]


def _digest() -> str:
    h = hashlib.sha256()
    for f in sources():
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found; libd3b200.so cannot be built")
    return exe


def build_variant(out: str, extra_flags) -> str:
    """Kernel experiments: another build of the same library with extra nvcc flags (e.g. -DD3_V2_TMEM=0),
    selected at run time through $TAD_DFTD3_B200_LIB."""
    cus = [f for f in sources() if f.endswith(".cu")]
    res = subprocess.run([nvcc(), *NVCC_FLAGS, *extra_flags, "-o", out, *cus], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + out)
    return out


INFO = os.path.join(HERE, "build_info.json")


def _record(mode: str, digest: str, seconds: float = 0.0) -> None:
    """`build_info.json` next to the library: which build the loaded .so is (source digest, flags, nvcc) and whether
    the last `build()` call compiled it ("compiled") or found it up to date ("reused: source digest unchanged")."""
    import json
    import time

    try:
        ver = subprocess.run([nvcc(), "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    except (OSError, RuntimeError, IndexError):
        ver = "nvcc unavailable"
    info = {"library": os.path.basename(LIB), "mode": mode, "source_digest": digest, "nvcc": ver,
            "flags": NVCC_FLAGS, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
            "compile_seconds": round(seconds, 1), "lib_bytes": os.path.getsize(LIB) if os.path.isfile(LIB) else 0}
    try:
        with open(INFO, "w") as fh:
            json.dump(info, fh, indent=1)
    except OSError:  # pragma: no cover
        pass


def build(force: bool = False, verbose: bool = False) -> str:
    import time

    digest = _digest()
    force = force or os.environ.get("TAD_DFTD3_B200_FORCE_BUILD") == "1"
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP):
        if open(STAMP).read().strip() == digest:
            _record("reused: source digest unchanged", digest)
            return LIB
    t0 = time.time()
    cus = [f for f in sources() if f.endswith(".cu")]
    cmd = [nvcc(), *NVCC_FLAGS, "-o", LIB, *cus]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libd3b200.so")
    with open(os.path.join(HERE, "ptxas_info.log"), "w") as fh:
        fh.write(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    _record("compiled", digest, time.time() - t0)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:            # python -m tad_dftd3_b200.build --variant out.so -DFLAG=0 ...
        k = sys.argv.index("--variant")
        print(build_variant(sys.argv[k + 1], sys.argv[k + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
