#!/usr/bin/env python
"""Builds mujoco-maze_b200/libmmz_dbg.so with extra -D flags (development aid), e.g.
    python tools/build_debug.py -DMMZ_PHASE_TIMING      # cycles per phase of block 0 of the hybrid kernel
    python tools/build_debug.py -DMMZ_DEBUG_UNIFORM     # prints where a warp's control flow diverged (mmz_dyn.cuh)
Use: MMZ_LIB=libmmz_dbg.so python bench.py ..."""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mujoco-maze_b200")
sys.path.insert(0, PKG)
from build_native import INSTANCES  # noqa: E402

extra = [a for a in sys.argv[1:] if a.startswith("-D")] or ["-DMMZ_DEBUG_UNIFORM"]
flags = "-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -prec-div=false -prec-sqrt=false -ftz=true -Xcompiler -fPIC".split() + extra
jobs = [(os.path.join(PKG, "csrc/mmz_api.cu"), "/tmp/dbg_api.o", [])]
for g, n, f in INSTANCES:
    jobs.append((os.path.join(PKG, "csrc/mmz_inst.cu"), f"/tmp/dbg_{g}_{n}_{f}.o", [f"-DMMZ_G={g}", f"-DMMZ_NVP={n}", f"-DMMZ_FEAT={f}"]))
for n, box in ((14, 0), (16, 1), (4, 1)):
    jobs.append((os.path.join(PKG, "csrc/mmz_hinst.cu"), f"/tmp/dbg_h{n}.o", [f"-DMMZ_NVP={n}", f"-DMMZ_BOX={box}"]))


def run(j):
    subprocess.check_call(["nvcc", *flags, *j[2], "-c", j[0], "-o", j[1]])
    return j[1]


with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(run, jobs))
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(PKG, "libmmz_dbg.so"), *objs])
print("ok")
