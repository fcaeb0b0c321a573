"""Builds the C-ABI shared library (CUDA kernels for sm_100a) in-tree with nvcc.

    python -m ultra_pytorch_b200.build            # -> ultra_pytorch_b200/lib/libultra_b200.so

The .so is git-ignored but travels to the GPU box with the gpurun snapshot; it only links libcudart
(statically), so it loads without torch.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libultra_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]


CXX = os.environ.get("CXX", "g++")
CXXFLAGS = ["-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-pthread", "-I/usr/local/cuda/include"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _host_sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cpp")))


def _fingerprint():
    h = hashlib.sha256()
    for p in _sources() + _host_sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
            [os.path.join(os.path.dirname(HERE), "include", "ultra_b200.h")]:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS + CXXFLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    fp = _fingerprint()
    if not force and os.path.isfile(LIBPATH) and os.path.isfile(stamp) and open(stamp).read().strip() == fp:
        return LIBPATH
    objs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        subprocess.check_call(cmd)
        objs.append(obj)
    for src in _host_sources():      # host-only code (the feed packer): plain g++
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-4] + ".host.o")
        subprocess.check_call([CXX] + CXXFLAGS + ["-c", src, "-o", obj])
        objs.append(obj)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIBPATH] + objs + \
        ["--cudart", "static", "-lcuda", "-Xcompiler", "-pthread"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(fp)
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
