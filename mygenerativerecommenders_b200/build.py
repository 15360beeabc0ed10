"""Build libgrb200.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python -m mygenerativerecommenders_b200.build [--force] [-v]

Objects are compiled in parallel (one nvcc per .cu) and linked into
mygenerativerecommenders_b200/libgrb200.so.  A source/header hash makes rebuilds incremental.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libgrb200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("GRB_NVCC_EXTRA", "").split()   # e.g. -DGRB_BWD_TIMELINE (developer probes)


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _dep_hash() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "grb200.h"]):
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: Path, dep: str, force: bool) -> tuple[Path, str]:
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".stamp")
    key = hashlib.sha256(src.read_bytes() + dep.encode()).hexdigest()
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == key:
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(key)
    (OBJ / (src.stem + ".ptxas.log")).write_text(r.stderr)
    return obj, r.stderr or " "


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    dep = _dep_hash()
    srcs = _sources()
    objs = []
    rebuilt = False
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(srcs)))) as ex:
        for obj, log in ex.map(lambda s: _compile(s, dep, force), srcs):
            objs.append(obj)
            if log:
                rebuilt = True
                if verbose:
                    print(log)
    stale = {p.stem for p in OBJ.glob("*.o")} - {s.stem for s in srcs}
    for name in stale:  # a deleted source must not linger in the link
        for ext in (".o", ".stamp", ".ptxas.log"):
            (OBJ / (name + ext)).unlink(missing_ok=True)
        rebuilt = True
    if rebuilt or not LIB.exists():
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode",
               "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
