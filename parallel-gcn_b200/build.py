"""Builds libgcn_b200.so (hand-written sm_100a kernels + C++ host mirror of the reference API) in-tree.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgcn_b200.so")
SYNTH_LIB = os.path.join(HERE, "libgcn_synth.so")  # the synthetic-workload generator alone (host C++, no CUDA): synth.py
SYNTH_SRC = os.path.join(HERE, "host", "src", "synth.cpp")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]
# kernels + C ABI: only the GCNB_API symbols are exported
KERNEL_FLAGS = COMMON + ["-Xcompiler", "-fvisibility=hidden"]
# C++ mirror of the reference API (Variable, Module..., GCN, Parser): must be linkable by user code
HOST_FLAGS = COMMON


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "host", "src", "*.cpp"))
                  + glob.glob(os.path.join(HERE, "host", "src", "*.cu")))


def headers():
    root = os.path.dirname(HERE)
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "host", "include", "*"))
                  + glob.glob(os.path.join(root, "include", "*.h")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in sources() + headers() + [os.path.abspath(__file__)])


def build_synth(force=False):
    deps = [SYNTH_SRC, os.path.join(os.path.dirname(HERE), "include", "gcnb_engine.h")]
    if force or not os.path.exists(SYNTH_LIB) or any(os.path.getmtime(d) > os.path.getmtime(SYNTH_LIB) for d in deps):
        subprocess.check_call([os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread",
                               "-fvisibility=hidden", "-o", SYNTH_LIB, SYNTH_SRC])
    return SYNTH_LIB


def build(force=False, verbose=False):
    build_synth(force)
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        deps = headers() + [src, os.path.abspath(__file__)]
        if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in deps):
            continue
        flags = HOST_FLAGS if (os.sep + "host" + os.sep) in src else KERNEL_FLAGS
        cmd = [nvcc] + flags + ["-x", "cu", "-c", src, "-o", obj, "-I", os.path.join(HERE, "host", "include")]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
        if verbose:
            sys.stderr.write(out)
    subprocess.check_call([nvcc, "--shared"] + ARCH + ["-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
