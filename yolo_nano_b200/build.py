"""In-tree nvcc build of the C-ABI library (sm_100a only).

`python -m yolo_nano_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles
here without a GPU; the resulting `libyolonano_b200.so` sits next to this file,
is git-ignored and travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libyolonano_b200.so"
STAMP = PKG / ".libyolonano_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    # IEEE semantics everywhere: NMS / decode parity depends on it (no fast-math).
    "--fmad=true", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                   + [ROOT / "include" / "yolonano_b200.h", Path(__file__)])
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library.  Returns its path."""
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    objs = []
    for src in _sources():
        obj = PKG / "build" / (src.stem + ".o")
        obj.parent.mkdir(exist_ok=True)
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(CSRC),
               "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}")
        (obj.parent / (src.stem + ".ptxas.log")).write_text(res.stderr)
        objs.append(str(obj))
    # No -lcuda: the container that builds this has no libcuda.so.1; the one driver
    # entry point we need (cuTensorMapEncodeTiled) is resolved at run time through
    # cudaGetDriverEntryPoint, and the runtime is linked statically.
    cmd = [nvcc_path(), "-shared", "-o", str(LIB), *objs, "-cudart", "static",
           "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        raise RuntimeError("link failed")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
