"""Builds librsvdb.so (hand-written sm_100a CUDA behind the C ABI of include/rsvdb.h) in-tree with nvcc.

    python -m rsvd_kamaneh_raganato_terrana_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  There is no JIT and no fallback: if the
library is missing or cannot be loaded, importing the engine fails loudly (see capi.py).
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
LIB = PKG / "librsvdb.so"
SOURCES = sorted(p.name for p in CSRC.glob("*.cu"))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", str(ROOT / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper without OpenMP specs); nvcc only needs a plain host compiler
    objs = []
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("librsvdb.so build failed")
    cmd = [_nvcc(), "-shared", "-ccbin", "/usr/bin/g++", "-o", str(LIB), *objs, "-lcudart", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("librsvdb.so link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
