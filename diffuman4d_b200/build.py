"""Build libd4d.so (sm_100a) in-tree with nvcc.  No torch, no CPU fallback: the product is this library.

Two libraries come out of the same sources:
  libd4d.so       the product (include/d4d.h): no measurement kernels, no ablation switches, no environment reads
                  on the data path
  libd4d_test.so  tools build (-DD4D_TEST_KERNELS -DD4D_ABLATE): additionally exports the UMMA probe and the
                  microbenchmarks (include/d4d_test.h) and honours D4D_GEMM_ABLATE / D4D_ATTN_ABLATE

    python -m diffuman4d_b200.build          # incremental
    python -m diffuman4d_b200.build --force
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libd4d.so")
TEST_LIB = os.path.join(HERE, "libd4d_test.so")
SOURCES = ["tmap.cu", "gemm_umma.cu", "attention_umma.cu", "attention_d64.cu", "norm.cu", "elementwise.cu", "unet.cu", "d4d_api.cu"]
TEST_SOURCES = SOURCES + ["probe.cu", "microbench.cu"]
TEST_DEFS = ["-DD4D_TEST_KERNELS", "-DD4D_ABLATE", "-DD4D_SPIN_LIMIT=(1u<<21)"]  # tools build: a protocol bug traps within seconds
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-diag-suppress", "177"]


def _deps_mtime() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(src: str, force: bool, hdr_t: float, test: bool = False) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".test.o" if test else ".o"))
    s = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(s), hdr_t):
        return obj
    cmd = [NVCC, *FLAGS, *(TEST_DEFS if test else []), "-I", CSRC, "-c", s, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def _link(lib: str, objs, force: bool) -> bool:
    need_link = force or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs)
    if need_link:
        cmd = [NVCC, "-shared", "-o", lib, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xlinker", "--exclude-libs,ALL"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return need_link


def build(force: bool = False, verbose: bool = True, test_lib: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _deps_mtime()
    jobs = [(s, False) for s in SOURCES] + ([(s, True) for s in TEST_SOURCES] if test_lib else [])
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda j: _compile(j[0], force, hdr_t, j[1]), jobs))
    linked = _link(LIB, objs[:len(SOURCES)], force)
    if test_lib:
        linked = _link(TEST_LIB, objs[len(SOURCES):], force) or linked
    if verbose:
        print(f"[d4d] {'built' if linked else 'up to date'}: {LIB}" + (f" + {os.path.basename(TEST_LIB)}" if test_lib else ""))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
