"""Builds libmmz.so (the CUDA step engine behind include/mmz.h) in-tree with nvcc for sm_100a.

Each (lanes-per-env, padded-nv) kernel instance is its own translation unit so the instances
compile in parallel; objects are cached under csrc/_build by source mtime.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libmmz.so")
# (lanes per env, padded nv, FEAT bits: 1 box geoms, 2 fluid, 4 sphere pairs between moving bodies) - keep in step with MMZ_INSTANCES in csrc/mmz_api.cu
INSTANCES = ((8, 4, 1), (8, 4, 7), (8, 8, 2), (8, 8, 7), (16, 14, 0), (16, 16, 1), (16, 16, 7), (32, 20, 7))
# Division and square root use the approximate (2 ulp) instructions and denormals flush to zero: +3 % on the Ant step,
# errors against the fp64 oracle unchanged (profiles/r1_parity.md). powf / logf stay precise: -use_fast_math buys
# another 1.4 % but triples the median velocity error; only the joint-angle sine / cosine use the SFU (mmz_math.cuh).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-prec-div=false", "-prec-sqrt=false", "-ftz=true", *os.environ.get("MMZ_NVCC_EXTRA", "").split(),
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(HERE, "..", "include", f) for f in ("mmz.h", "mmz_model.h")]
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({' '.join(cmd)}):\n{r.stdout[-4000:]}")


def build(force: bool = False, verbose: bool = True) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _sources()
    jobs = []
    objs = []
    api_o = os.path.join(OBJ, "mmz_api.o")
    objs.append(api_o)
    if force or _stale(api_o, hdrs + [os.path.join(CSRC, "mmz_api.cu")]):
        jobs.append(([nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, "mmz_api.cu"), "-o", api_o], api_o + ".log"))
    for g, nvp, feat in INSTANCES:
        o = os.path.join(OBJ, f"mmz_inst_{g}_{nvp}_{feat}.o")
        objs.append(o)
        if force or _stale(o, hdrs + [os.path.join(CSRC, "mmz_inst.cu")]):
            jobs.append(([nvcc, *NVCC_FLAGS, f"-DMMZ_G={g}", f"-DMMZ_NVP={nvp}", f"-DMMZ_FEAT={feat}", "-c",
                          os.path.join(CSRC, "mmz_inst.cu"), "-o", o], o + ".log"))
    for nvp, box in ((14, 0), (16, 1), (4, 1)):
        h_o = os.path.join(OBJ, f"mmz_hinst_{nvp}.o")
        objs.append(h_o)
        if force or _stale(h_o, hdrs + [os.path.join(CSRC, "mmz_hinst.cu")]):
            jobs.append(([nvcc, *NVCC_FLAGS, f"-DMMZ_NVP={nvp}", f"-DMMZ_BOX={box}", "-c", os.path.join(CSRC, "mmz_hinst.cu"), "-o", h_o], h_o + ".log"))
    if jobs:
        if verbose:
            print(f"[build_native] compiling {len(jobs)} translation unit(s) for sm_100a ...", flush=True)
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    if jobs or force or _stale(LIB, objs):
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs], os.path.join(OBJ, "link.log"))
        if verbose:
            print(f"[build_native] linked {LIB}", flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
