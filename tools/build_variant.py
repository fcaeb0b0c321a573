"""Builds a variant of the library into tests/_build/libub200_<name>.so with extra nvcc -D flags.
Usage: python tools/build_variant.py NAME -DFOO=1 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultra_pytorch_b200 import build as b
name, extra = sys.argv[1], sys.argv[2:]
outdir = os.path.join(ROOT, "tests", "_build")
os.makedirs(outdir, exist_ok=True)
out = os.path.join(outdir, "libub200_%s.so" % name)
objs = []
for src in b._sources():
    obj = os.path.join(outdir, os.path.basename(src)[:-3] + ".%s.o" % name)
    subprocess.check_call([b.NVCC] + b.FLAGS + extra + ["-c", src, "-o", obj])
    objs.append(obj)
for src in b._host_sources():
    obj = os.path.join(outdir, os.path.basename(src)[:-4] + ".%s.host.o" % name)
    subprocess.check_call([b.CXX] + b.CXXFLAGS + ["-c", src, "-o", obj])
    objs.append(obj)
subprocess.check_call([b.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs +
                      ["--cudart", "static", "-lcuda", "-Xcompiler", "-pthread"])
print(out)
