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


if __name__ == "__main__":
    print(build())
