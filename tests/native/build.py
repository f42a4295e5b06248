"""Builds the test-only native helpers (real cuRAND on device and on host) next to this file."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build():
    src = os.path.join(HERE, "curand_ref.cu")
    dev, host = os.path.join(HERE, "libcurand_ref_device.so"), os.path.join(HERE, "libcurand_ref_host.so")
    if not os.path.exists(dev) or os.path.getmtime(dev) < os.path.getmtime(src):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "--shared",
                               "-Xcompiler", "-fPIC", "-o", dev, src])
    if not os.path.exists(host) or os.path.getmtime(host) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-DCURAND_REF_HOST", "-x", "c++", "-I/usr/local/cuda/include",
                               "-o", host, src])
    return dev, host


def build_module_chain():
    """tests/native/module_chain.cpp (the product's C++ module classes composed by hand) against the product's headers
    and libgcn_b200.so; g++ only, so it can be (re)built on the GPU box as well"""
    root = os.path.dirname(os.path.dirname(HERE))
    pkg = os.path.join(root, "parallel-gcn_b200")
    src, exe = os.path.join(HERE, "module_chain.cpp"), os.path.join(HERE, "module_chain")
    lib = os.path.join(pkg, "libgcn_b200.so")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(pkg, "host", "include"), "-I/usr/local/cuda/include",
                               src, "-o", exe, "-L" + pkg, "-lgcn_b200", "-L/usr/local/cuda/lib64", "-lcudart",
                               "-Wl,-rpath," + pkg, "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


if __name__ == "__main__":
    print(build())
    print(build_module_chain())
