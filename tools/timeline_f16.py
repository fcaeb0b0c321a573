"""GPU debugging aid: SM-clock timeline of CTA 0 of fwd16_kernel.  Builds a -DUB200_F16_TIMELINE copy of the library
into tests/_build/ (run with --build here, where nvcc lives), then on the GPU: UB200_LIB=tests/_build/libub200_tl.so."""
import ctypes, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.environ.get("TL_VARIANT", "")          # e.g. "64_104": control / worker register caps
OUT = os.path.join(ROOT, "tests", "_build", "libub200_tl%s.so" % VAR)
if "--build" in sys.argv:
    from ultra_pytorch_b200 import build as b
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    for src in b._sources():
        obj = os.path.join(os.path.dirname(OUT), os.path.basename(src)[:-3] + ".tl%s.o" % VAR)
        extra = ["-DUB200_REGS_CTRL=%s" % VAR.split("_")[0], "-DUB200_REGS_WORK=%s" % VAR.split("_")[1]] if VAR else []
        subprocess.check_call([b.NVCC] + b.FLAGS + ["-DUB200_F16_TIMELINE"] + extra + ["-c", src, "-o", obj])
        objs.append(obj)
    for src in b._host_sources():
        obj = os.path.join(os.path.dirname(OUT), os.path.basename(src)[:-4] + ".tl.host.o")
        subprocess.check_call([b.CXX] + b.CXXFLAGS + ["-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call([b.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs +
                          ["--cudart", "static", "-lcuda", "-Xcompiler", "-pthread"])
    print(OUT)
    sys.exit(0)
os.environ["UB200_LIB"] = os.environ.get("UB200_TL_LIB", OUT)
import numpy as np, torch
from ultra_pytorch_b200 import _capi
from ultra_pytorch_b200.engine import RankerEngine
lib = _capi.lib
lib.ub200_f16_timeline.restype = ctypes.c_int
lib.ub200_f16_timeline.argtypes = [ctypes.c_void_p]
F, L, B = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (136, 40, 256)
hidden = [int(h) for h in sys.argv[4].split(",")] if len(sys.argv) > 4 else [256, 128, 64]
training = "--inference" not in sys.argv
M = L * B
eng = RankerEngine(F, hidden)
eng.params.normal_(0, 0.05)
feats = torch.rand(M + 1, F, device="cuda")
docid = torch.randint(0, M, (M,), dtype=torch.int32, device="cuda")
for rep in range(3):
    eng.forward(feats, docid, L, B, training=training)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * 192)()
lib.ub200_f16_timeline(buf)
t = np.array(list(buf), dtype=np.int64).reshape(3, 64)
t0 = t[0, 0]
us = lambda x: (x - t0) / 1965.0
names = {0: "start", 1: "setup done", 2: "worker begin", 5: "L0 input stats done", 3: "L0 A operand produced", 4: "end"}
for q in range(len(hidden)):
    names[8 + 4 * q] = "L%d accumulator ready" % q
    names[9 + 4 * q] = "L%d pass1 done (ELU, Y store)" % q
    names[10 + 4 * q] = "L%d row stats exchanged" % q
    names[11 + 4 * q] = "L%d pass2 done (next A operand)" % q
ev = [(us(t[0, i]), "worker: " + n) for i, n in names.items() if t[0, i] > 0]
for q in range(len(hidden)):
    for r, nm in ((1, "mma"), (2, "tma")):
        if t[r, 2 * q] > 0:
            ev.append((us(t[r, 2 * q]), "%s: L%d begin" % (nm, q)))
            ev.append((us(t[r, 2 * q + 1]), "%s: L%d all issued" % (nm, q)))
for c in range(12):
    if t[1, 16 + 2 * c] > 0:
        ev.append((us(t[1, 16 + 2 * c]), "mma:   chunk L%d.%d operands ready" % (c // 4, c % 4)))
        ev.append((us(t[1, 17 + 2 * c]), "mma:   chunk L%d.%d issued" % (c // 4, c % 4)))
def show(ev):
    for tt, n in sorted(ev):
        print("%8.2f us  %s" % (tt, n))
    print()
show(ev)
if "--bwd" in sys.argv:
    dsc = torch.randn(B, L, device="cuda")
    for rep in range(3):
        eng.forward(feats, docid, L, B, training=True)
        eng.backward(feats, docid, L, B, dsc)
        torch.cuda.synchronize()
    lib.ub200_f16_timeline(buf)
    t = np.array(list(buf), dtype=np.int64).reshape(3, 64)
    t0 = t[0, 20]
    nl = len(hidden)
    ev = [(us(t[0, 20]), "bwd worker: begin (after setup)"), (us(t[0, 36]), "bwd worker: final-layer sums exchanged"),
          (us(t[0, 21]), "bwd worker: final step done (dZ stored, A operand written)"), (us(t[0, 39]), "bwd worker: end")]
    for q in range(nl - 1, -1, -1):
        if t[0, 25 + 4 * q] > t0:
            ev.append((us(t[0, 25 + 4 * q]), "bwd worker: dZ_%d stores drained" % q))
    for q in range(nl - 1, 0, -1):
        ev.append((us(t[0, 22 + 4 * q]), "bwd worker: dgrad_%d accumulator ready" % q))
        ev.append((us(t[0, 24 + 4 * q]), "bwd worker: dgrad_%d LN-backward sums exchanged" % q))
        ev.append((us(t[0, 23 + 4 * q]), "bwd worker: dgrad_%d epilogue done" % q))
    show(ev)
if "--wgrad" in sys.argv:
    dsc = torch.randn(B, L, device="cuda")
    for rep in range(3):
        eng.forward(feats, docid, L, B, training=True)
        eng.backward(feats, docid, L, B, dsc)
        torch.cuda.synchronize()
    lib.ub200_f16_timeline(buf)
    t = np.array(list(buf), dtype=np.int64).reshape(3, 64)
    t0 = t[0, 38]
    ev = [(us(t[0, 38]), "wgrad worker: start"), (us(t[0, 39]), "wgrad worker: setup done"),
          (us(t[0, 54]), "wgrad worker: all chunks produced"), (us(t[0, 55]), "wgrad worker: accumulator ready"),
          (us(t[0, 56]), "wgrad worker: column sums written"), (us(t[0, 57]), "wgrad worker: end")]
    for c in range(6):
        if t[0, 40 + 2 * c] > t0:
            ev.append((us(t[0, 40 + 2 * c]), "wgrad worker: chunk %d slot free, loads of chunk %d issued" % (c, c + 1)))
            ev.append((us(t[0, 41 + 2 * c]), "wgrad worker: chunk %d stored" % c))
        if t[1, 40 + 2 * c] > t0:
            ev.append((us(t[1, 40 + 2 * c]), "wgrad mma: chunk %d operands ready" % c))
            ev.append((us(t[1, 41 + 2 * c]), "wgrad mma: chunk %d issued" % c))
    show(ev)
