"""Tuning helper: builds libgcn_b200 variants that differ in one -D define of one source (e.g. GCNB_DRAIN of spmm.cu).
usage: python scripts/build_variants.py spmm.cu GCNB_DRAIN 0 1 2  ->  parallel-gcn_b200/build/variants/libgcn_b200_GCNB_DRAIN<k>.so"""
import glob, os, subprocess, sys
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
b = pkg._build
b.build()
src_name, macro, values = sys.argv[1], sys.argv[2], sys.argv[3:]
objdir = os.path.join(b.HERE, "build")
outdir = os.path.join(b.HERE, "variants")
os.makedirs(outdir, exist_ok=True)
src = os.path.join(b.HERE, "csrc", src_name)
others = [o for o in glob.glob(os.path.join(objdir, "*.o")) if os.path.basename(o) != src_name + ".o"]
for v in values:
    obj = os.path.join(outdir, "%s_%s%s.o" % (src_name, macro, v))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + b.KERNEL_FLAGS + ["-D%s=%s" % (macro, v), "-x", "cu", "-c", src, "-o", obj,
                           "-I", os.path.join(b.HERE, "host", "include")])
    lib = os.path.join(outdir, "libgcn_b200_%s%s.so" % (macro, v))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "--shared"] + b.ARCH + ["-o", lib, obj] + others)
    print(lib)
