"""GPU debugging aid: timeline of CTA 0 of the fused forward kernel (UB200_LIB=tests/_build/libultra_b200_timeline.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ultra_pytorch_b200 import _capi
from ultra_pytorch_b200.engine import RankerEngine
lib = _capi.lib
lib.ub200_fused_timeline.restype = ctypes.c_int
lib.ub200_fused_timeline.argtypes = [ctypes.c_void_p]
L, B, F, hidden = 40, 256, 136, [256, 128, 64]
M = L * B
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
lib.ub200_fused_timeline(buf)
t = np.array(list(buf), dtype=np.int64); t0 = t[0]
print("forward() total %.1f us" % (1e3 * ev0.elapsed_time(ev1)))
names = {0: "start", 1: "setup done", 2: "stats0 done", 5: "final done", 6: "end"}
for j in range(3):
    names[8 + 4 * j] = "L%d producers done" % j; names[9 + 4 * j] = "L%d accum ready" % j
    names[10 + 4 * j] = "L%d epilogue pass1" % j; names[11 + 4 * j] = "L%d stats done" % j
for i in sorted(names, key=lambda i: t[i]):
    print("   %-20s +%.2f us" % (names[i], (t[i] - t0) / 1e3))

ex = {40: "L0 ep blk0 start", 41: "  tmem ld x2 done", 42: "  elu done", 43: "  sums done", 44: "  tmem st issued", 45: "  Y stored",
      47: "L1 chunk2 arrived", 48: "L1 chunk3 slot free", 49: "  tmem ld8 done", 50: "  split+sts done", 51: "  fence done", 52: "  arrived"}
for i in sorted(ex):
    print("   %-20s +%.2f us" % (ex[i], (t[i] - t0) / 1e3))
