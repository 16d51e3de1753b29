"""In-tree build of libb200q.so (sm_100a) and the host-only math check library.

    python -m quantizers_b200.build            # or __graft_entry__.build()

Explicit nvcc, no JIT cache: the built .so files sit next to the package and travel to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "lib", "libb200q.so")
HOSTMATH = os.path.join(HERE, "lib", "libb200q_hostmath.so")

SOURCES = ["abi.cu", "quant_group.cu", "quant_group_fast.cu", "quant_group_tma.cu", "quant_tile_fast.cu", "quant_channel_fast.cu", "decode_fast.cu", "quant_nvfp4_persistent.cu", "quant_elementwise.cu", "elementwise_fast.cu", "quant_tile.cu", "observers.cu", "pack.cu", "awq_stats.cu",
           "awq_gemm.cu", "awq_attn.cu", "awq_fq_fast.cu", "awq_attn_core.cu"]
HEADERS = ["qmath.cuh", "common.cuh", "kernels.cuh", "fastmath.cuh", "async.cuh", "fp4.cuh", os.path.join("..", "..", "include", "b200q.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# no --use_fast_math: IEEE division / no FMA contraction is part of the parity contract (qmath.cuh)
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--fmad=false", "-Xptxas", "-v"] + os.environ.get("B200Q_NVCC_DEFS", "").split()


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {' '.join(cmd)}")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append(([NVCC] + ARCH + FLAGS + ["-c", s, "-o", o], o + ".log"))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    if force or jobs or not os.path.exists(LIB):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"], os.path.join(OBJ, "link.log"))
    hm_src = os.path.join(CSRC, "hostmath.cu")
    if force or not _newer(HOSTMATH, [hm_src] + hdrs):
        _run([NVCC, "-O2", "-std=c++17", "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-cudart", "none",
              hm_src, "-o", HOSTMATH], os.path.join(OBJ, "hostmath.log"))
    if verbose:
        print(f"built {LIB}\nbuilt {HOSTMATH}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
