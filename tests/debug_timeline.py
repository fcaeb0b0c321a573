"""GPU debugging aid: in-kernel timeline (globaltimer stamps of CTA 0) of the tensor-core GEMM kernel.
Run with UB200_LIB=tests/_build/libultra_b200_timeline.so (built with -DUB200_TC_TIMELINE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ultra_pytorch_b200 import _capi
from ultra_pytorch_b200.engine import RankerEngine
lib = _capi.lib
lib.ub200_tc_timeline.restype = ctypes.c_int
lib.ub200_tc_timeline.argtypes = [ctypes.c_void_p]
L, B = 40, 256
M = L * B
for F, hidden in ((136, [256]), (256, [128]), (128, [64])):
    eng = RankerEngine(F, hidden)
    eng.params.normal_(0, 0.05)
    feats = torch.rand(M + 1, F, device="cuda")
    docid = torch.randint(0, M, (M,), dtype=torch.int32, device="cuda")
    for rep in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(); eng.forward(feats, docid, L, B, training=True); ev1.record()
        torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 64)()
    lib.ub200_tc_timeline(buf)
    t = np.array(list(buf), dtype=np.int64)
    t0 = t[0]
    names = {0: "start", 1: "setup done", 2: "producer start", 3: "mma all issued", 4: "accum ready", 5: "epilogue done", 6: "end"}
    print("F=%d N=%d: forward() total %.1f us (prep+stats+gemm+final)" % (F, hidden[0], 1e3 * ev0.elapsed_time(ev1)))
    for i in range(7):
        print("   %-16s +%.2f us" % (names[i], (t[i] - t0) / 1e3))
    n_chunks = (F + 31) // 32
    print("   chunks (slot free -> stored):", " ".join("[%.2f,%.2f]" % ((t[8 + 2 * i] - t0) / 1e3, (t[9 + 2 * i] - t0) / 1e3) for i in range(min(n_chunks, 8))))
