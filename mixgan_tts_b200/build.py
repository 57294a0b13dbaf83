"""Build the shared libraries in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

  libmixgan_b200.so       the product library: every symbol include/mixgan_b200.h declares.  It never allocates, never
                          synchronises and reads no debug environment variables.
  libmixgan_b200_dbg.so   test/diagnostic superset (same sources compiled with -DMGB_DEBUG_BUILD, plus umma_probe.cu and
                          the symbols of include/mixgan_b200_probe.h): in-kernel role profiling (MGB_PROFILE), timing
                          experiments (MGB_DEBUG_MODE), launch tracing (MGB_TRAIN_TRACE) and the tcgen05 / bulk-copy
                          probes used by tests/test_umma_probe.py and scripts/*_rate.py.  Never loaded by the product path.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmixgan_b200.so")
LIB_DBG = os.path.join(HERE, "libmixgan_b200_dbg.so")
SOURCES = ["abi.cu", "fp32_path.cu", "train_fp32.cu", "train_bf16.cu", "fused_bf16.cu", "conv1d_f32.cu", "train_branch.cu", "tcnet.cu", "auxdec.cu", "hifigan.cu"]
DBG_ONLY_SOURCES = ["umma_probe.cu"]
HEADERS = ["common.cuh", "tc05.cuh", "small_ops.cuh", "gemm_fp32.cuh", "train_small.cuh", "tmap.cuh", "tcnet.cuh",
           os.path.join("..", "..", "include", "mixgan_b200.h"), os.path.join("..", "..", "include", "mixgan_b200_probe.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# --cudart shared: the artefact carries no copy of the runtime (and none of its symbol names)
LINK_FLAGS = ["-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out: str, deps: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(nvcc, jobs, verbose):
    def one(job):
        src, obj, extra = job
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        subprocess.run(cmd, check=True)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(one, jobs))


def build(force: bool = False, verbose: bool = False, debug: bool = True) -> str:
    nvcc = _nvcc()
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    variants = [("build", SOURCES, [], LIB)]
    if debug:
        variants.append(("build_dbg", SOURCES + DBG_ONLY_SOURCES, ["-DMGB_DEBUG_BUILD"], LIB_DBG))
    jobs, links = [], []
    for objdir_name, sources, extra, lib in variants:
        objdir = os.path.join(HERE, objdir_name)
        os.makedirs(objdir, exist_ok=True)
        objs = []
        for src in sources:
            s = os.path.join(CSRC, src)
            o = os.path.join(objdir, src.replace(".cu", ".o"))
            objs.append(o)
            if force or _stale(o, [s] + hdrs):
                jobs.append((s, o, extra))
        links.append((lib, objs))
    if jobs:
        _compile(nvcc, jobs, verbose)
    for lib, objs in links:
        if force or _stale(lib, objs):
            subprocess.run([nvcc, *LINK_FLAGS, "-o", lib, *objs], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--no-debug" not in sys.argv))
