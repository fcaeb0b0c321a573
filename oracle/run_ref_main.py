"""TEST INFRASTRUCTURE ONLY: runs the UNMODIFIED reference driver (oracle/_ref/main.py) with the import shims of
oracle/ref_shim.py.  Usage: python oracle/run_ref_main.py <main.py args...>   (cwd is switched to oracle/_ref)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

if __name__ == "__main__":
    ref_shim.load()
    os.chdir(ref_shim.REF_COPY)
    sys.argv = [os.path.join(ref_shim.REF_COPY, "main.py")] + sys.argv[1:]
    runpy.run_path(sys.argv[0], run_name="__main__")
