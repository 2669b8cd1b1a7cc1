"""Build libstarr_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("ST_BUILD_DIR") or os.path.join(CSRC, "build")
LIB = os.environ.get("ST_BUILD_LIB") or os.path.join(HERE, "libstarr_b200.so")
SOURCES = ["st_tree_kernels.cu", "st_update.cu", "st_shard.cu", "st_mintree.cu", "st_peer.cu", "st_axis_fold.cu", "st_capi.cu"]
HEADERS = ["st_common.cuh", "st_internal.h", "st_ordered_fold.cuh", "st_partition.cuh", "st_partition_hybrid.cuh", "st_update_defs.cuh", "st_update_small.cuh", os.path.join("..", "..", "include", "starr_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # float parity: never contract a*b+c
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    *(["-DST_SMALL_TIMING"] if os.environ.get("ST_SMALL_TIMING") else []),
    *([f"-DST_RETIRE_SEG={os.environ['ST_RETIRE_SEG']}"] if os.environ.get("ST_RETIRE_SEG") else []),
    *([f"-DST_HUGE_SEG={os.environ['ST_HUGE_SEG']}"] if os.environ.get("ST_HUGE_SEG") else []),
    *([f"-DST_HY_SEG={os.environ['ST_HY_SEG']}"] if os.environ.get("ST_HY_SEG") else []),
    *([f"-DST_SMALL_RETIRE_SEG={os.environ['ST_SMALL_RETIRE_SEG']}"] if os.environ.get("ST_SMALL_RETIRE_SEG") else []),
    "-Xptxas", "-v" if os.environ.get("ST_PTXAS_V") else "-O3",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append([NVCC, *FLAGS, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            if l.strip():
                print(l)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
             "-o", LIB, *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
