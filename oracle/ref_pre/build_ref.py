"""Build oracle/_ref/libdvins_refpre.so: the REFERENCE's own CUDA pre/post-processing kernels, compiled from the
sources where they lie under /root/reference (nothing is copied), plus oracle/ref_pre/ref_pre_wrap.cu.

    python oracle/ref_pre/build_ref.py

* only three reference files are needed and they have no external dependencies:
  loop_fusion/src/deep_net/tensorrt_tools/{preprocess_kernel.cu, cuda_tools.cpp, ilogger.cpp}
* the reference's own build system (catkin/cmake, TensorRT, OpenCV, faiss) is NOT run; the rest of the path
  (TensorRT engines) cannot be built here - see DESIGN.md §2.
* `-include cstdint`: ilogger.hpp uses uint8_t without including <cstdint> (fails on gcc 13); a command-line include,
  the sources are untouched.
* output is git-ignored but travels to the GPU box with the working-tree snapshot; /root/reference does not exist
  there, so tests only ever load the prebuilt .so.
TEST INFRASTRUCTURE: only tests/ may load the result.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
LIB = os.path.join(OUT_DIR, "libdvins_refpre.so")            # the reference's own nvcc flags (--use_fast_math)
LIB_IEEE = os.path.join(OUT_DIR, "libdvins_refpre_ieee.so")  # same sources, IEEE arithmetic (-fmad=false, no fast-math)
# loop_fusion/CMakeLists.txt:92: --default-stream per-thread -lineinfo --use_fast_math --disable-warnings
REF_NVCC_FLAGS = ["--default-stream", "per-thread", "--use_fast_math"]
IEEE_NVCC_FLAGS = ["--default-stream", "per-thread", "-fmad=false"]
REF = "/root/reference/loop_fusion/src/deep_net/tensorrt_tools"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "preprocess_kernel.cu"))


def build(force: bool = False) -> str | None:
    """Both variants: the reference's shipped flags, and IEEE (what the oracle's fp32 restatement states)."""
    _build_one(LIB_IEEE, IEEE_NVCC_FLAGS, force)
    return _build_one(LIB, REF_NVCC_FLAGS, force)


def _build_one(LIB: str, extra: list, force: bool) -> str | None:
    if not available():
        return LIB if os.path.exists(LIB) else None
    srcs = [os.path.join(REF, "preprocess_kernel.cu"), os.path.join(REF, "cuda_tools.cpp"),
            os.path.join(REF, "ilogger.cpp"), os.path.join(HERE, "ref_pre_wrap.cu")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s) + ".o")
        if s.endswith(".cu"):
            cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++14", "-w", "-Xcompiler", "-fPIC",
                   "--pre-include", "cstdint", "-I", REF, "-c", s, "-o", o] + extra
        else:
            cmd = ["g++", "-O2", "-std=c++14", "-w", "-fPIC", "-include", "cstdint", "-I", REF,
                   "-I", "/usr/local/cuda/include", "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("reference pre-processing build failed: " + s)
        objs.append(o)
    r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("reference pre-processing link failed")
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
