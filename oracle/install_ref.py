"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Installs an unmodified copy of the Python reference (ULTR-Community/ULTRA_pytorch) from
/root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the copy travels to
the GPU box where /root/reference does not exist).  The copy is what `tests/golden/make_goldens.py`
imports to generate golden vectors and what `bench.py --impl reference` times on the host cores.
No reference source is ever committed.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")


def install(force=False):
    """Copy the reference tree (python package + main.py + example JSON fixtures + toy data)."""
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(os.path.join(REF_DST, "ultra"))
    if os.path.isdir(os.path.join(REF_DST, "ultra")) and not force:
        return True
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    for name in ("ultra", "main.py", "example", "tests", "libsvm_tools"):
        src = os.path.join(REF_SRC, name)
        dst = os.path.join(REF_DST, name)
        if os.path.isdir(src):
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(src):
            shutil.copy(src, dst)
    for root, dirs, files in os.walk(REF_DST):
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    return True


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("oracle/_ref installed" if ok else "reference unavailable (no /root/reference, no oracle/_ref)")
