"""In-tree build of libdvins_b200.so (sm_100a only).  `python -m d_vins_b200.build [--force]`.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box with the
working-tree snapshot.  No torch involvement: the library is a plain C-ABI shared object.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libdvins_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
         "--expt-relaxed-constexpr", "-I", os.path.join(os.path.dirname(HERE), "include")]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")) and "shim" not in root:
                out.append(os.path.join(root, f))
    return out


def _headers_mtime():
    m = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".h", ".cuh", ".hpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    inc = os.path.join(os.path.dirname(HERE), "include")
    for f in os.listdir(inc):
        m = max(m, os.path.getmtime(os.path.join(inc, f)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    hm = _headers_mtime()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(BUILD, os.path.relpath(s, CSRC).replace(os.sep, "_") + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hm):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r.returncode, r.stdout + r.stderr

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, rc, log in ex.map(cc, jobs):
            if rc != 0 or verbose:
                sys.stderr.write("[build] %s\n%s\n" % (os.path.relpath(s, HERE), log))
            failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if jobs or not os.path.exists(LIB):
        # drop objects of deleted sources
        keep = set(objs)
        for f in os.listdir(BUILD):
            if f.endswith(".o") and os.path.join(BUILD, f) not in keep:
                os.remove(os.path.join(BUILD, f))
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


def build_shim_demo() -> str:
    """g++ build of the reference-facing C++ facade (csrc/shim) + the keyframe.cpp-order demo driver."""
    lib = LIB if os.path.exists(LIB) else build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "shim_demo")
    srcs = [os.path.join(CSRC, "shim", "deep_net_shim.cpp"),
            os.path.join(os.path.dirname(HERE), "tests", "cpp", "shim_demo.cpp")]
    deps = srcs + [lib, os.path.join(CSRC, "shim", "deep_net_shim.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(x) > os.path.getmtime(exe) for x in deps):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", exe] + srcs + [
            "-L" + HERE, "-ldvins_b200", "-Wl,-rpath," + HERE]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("shim build failed")
    return exe


def build_stream_demo() -> str:
    """g++ build of the ROS-free keyframe stream driver (csrc/shim/loop_closure.cpp) + tests/cpp/stream_demo.cpp."""
    lib = LIB if os.path.exists(LIB) else build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "stream_demo")
    srcs = [os.path.join(CSRC, "shim", "loop_closure.cpp"),
            os.path.join(os.path.dirname(HERE), "tests", "cpp", "stream_demo.cpp")]
    deps = srcs + [lib, os.path.join(CSRC, "shim", "loop_closure.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(x) > os.path.getmtime(exe) for x in deps):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", exe] + srcs + ["-L" + HERE, "-ldvins_b200", "-Wl,-rpath," + HERE]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("stream demo build failed")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
