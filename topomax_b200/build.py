"""Builds ``libtopomax_b200.so`` in-tree with nvcc for sm_100a (no JIT cache: the built
library travels with the source tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["tm_engine.cu"]
HEADERS = ["tm_common.cuh", "tm_element.cuh", "tm_elast.cuh", "tm_p1.cuh", "tm_vec.cuh", "tm_mg.cuh",
           "tm_filter_pcg.cuh", "tm_p1mg.cuh", "tm_tail.cuh", "tm_comm.h", "tm_p2p.cuh", "tm_dem.cuh", "tm_fluid.cuh", "tm_fluid_cuda.cuh", "tm_trimg.cuh", "tm_trimg_cuda.cuh",
           "tm_tables.h", os.path.join("..", "..", "include", "topomax_b200.h")]
OUTPUT = os.path.join(HERE, "libtopomax_b200.so")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.isfile(OUTPUT):
        return True
    built = os.path.getmtime(OUTPUT)
    csrc = os.path.join(HERE, "csrc")
    return any(os.path.getmtime(os.path.join(csrc, f)) > built for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUTPUT
    csrc = os.path.join(HERE, "csrc")
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC", "-o", OUTPUT,
    ] + [os.path.join(csrc, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return OUTPUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
