"""Build libgenconvit_b200.so in-tree with nvcc for sm_100a (no torch headers, no JIT cache).

    python -m genconvit_b200.build            # incremental
    python -m genconvit_b200.build --force

The shared library travels to the GPU box with the repo snapshot; it is
git-ignored (built artefact) but not gpurun-ignored.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libgenconvit_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["api.cu", "gemm_tcgen05.cu", "mlp_fused.cu", "gemm_simt.cu", "convnext_ops.cu", "dwconv_mma.cu", "autoenc_ops.cu", "conv3x3_c16.cu", "conv3x3_c32.cu", "conv3x3_tc.cu", "convt_mma.cu", "stem_fused.cu", "swin_ops.cu", "preprocess.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
FLAGS += os.environ.get("GCV_NVCC_FLAGS", "").split()      # e.g. -DGCV_GEMM_WHATIF: the GEMM's timing-experiment switches


def _deps():
    return [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_ptx.cuh"), os.path.join(INCLUDE, "genconvit_b200.h")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _deps()):
        return obj, ""
    r = subprocess.run([NVCC, *FLAGS, "-c", path, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
        if os.path.exists(LIB):
            os.remove(LIB)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(_compile, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
