"""Build the seeksv_b200 shared library and CLI in-tree (sm_100a only).

    python -m seeksv_b200.build            # libseeksv_b200.so + bin/seeksv

nvcc cross-compiles without a GPU; the products are git-ignored but travel to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libseeksv_b200.so")
CLI = os.path.join(HERE, "bin", "seeksv")
OBJ = os.path.join(HERE, "build")
CU = ["csrc/walk.cu", "csrc/getclip.cu", "csrc/getsv.cu", "csrc/inflate.cu", "csrc/gzip.cu", "csrc/clipjoin.cu", "csrc/api.cu"]
CPP = ["host/bamfile.cpp", "host/junction.cpp", "host/commands.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-pthread"]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(HERE, "..", "include", "seeksv_b200.h")]
    for d in ("csrc", "host"):
        for f in os.listdir(os.path.join(HERE, d)):
            if f.endswith((".h", ".cuh")):
                hs.append(os.path.join(HERE, d, f))
    return hs


def _run(cmd):
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[-1])
    return r


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    hdrs = _headers()
    jobs, objs = [], []
    for src in CU + CPP:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if _stale(obj, [os.path.join(HERE, src)] + hdrs):
            if src.endswith(".cu"):
                jobs.append([NVCC] + NVFLAGS + ["-c", src, "-o", obj])
            else:
                jobs.append(["g++"] + CXXFLAGS + ["-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(_run, jobs))
    if jobs or _stale(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lz", "-lpthread", "-Xlinker", "-rpath,$ORIGIN"])
    if _stale(CLI, [LIB, os.path.join(HERE, "host/main.cpp")]):
        _run(["g++"] + CXXFLAGS + ["host/main.cpp", "-o", CLI, "-L" + HERE, "-lseeksv_b200", "-Wl,-rpath,$ORIGIN/.."])
    build_tools()
    if verbose:
        print("built", LIB, "and", CLI)
    return LIB


def build_tools():
    """The test/bench helpers (workload simulator, stand-in aligner for the external bwa step) and the stand-alone call
    evaluator svcompare: plain host C++."""
    tools = os.path.join(HERE, "..", "tools")
    for tool in ("svsim", "minialign", "svcompare"):
        out = os.path.join(os.path.dirname(CLI), tool)
        src = os.path.join(tools, tool + ".cpp")
        if _stale(out, [src]):
            _run(["g++", "-O2", "-std=c++17", "-pthread", src, "-o", out, "-lz"])
    # the device join's rules on the CPU, against the host mirror (tests; needs the host layer's sources)
    out = os.path.join(os.path.dirname(CLI), "clipjoin_sim")
    deps = [os.path.join(tools, "clipjoin_sim.cpp"), os.path.join(HERE, "host/junction.cpp"), os.path.join(HERE, "host/bamfile.cpp")]
    if _stale(out, deps + _headers()):
        _run(["g++", "-O2", "-std=c++17", "-pthread"] + deps + ["-o", out, "-lz"])


if __name__ == "__main__":
    build(verbose=True)
