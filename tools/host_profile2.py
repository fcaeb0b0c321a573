"""Pieces of RankerEngine.stage() timed separately (rotating over 8 host feeds, like the e2e leg of bench.py)."""
import os, sys, time, types, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ultra_pytorch_b200 import synth, _capi
from ultra_pytorch_b200.engine import RankerEngine, column_ptrs, _stream
lib = _capi.lib
F, L, B = 136, 40, 256
eng = RankerEngine(F, [256, 128, 64])
feeds = [synth.make_feed(i, F, L, B, "click") for i in range(8)]
pc = time.perf_counter
def med(f, n=200):
    ts = []
    for i in range(n):
        t0 = pc(); f(i); ts.append(pc() - t0)
    return np.median(ts) * 1e6
D = lambda f: [f["docid_input%d" % l] for l in range(L)]
Y = lambda f: [f["label%d" % l] for l in range(L)]
print("list build      %.1f us" % med(lambda i: (D(feeds[i % 8]), Y(feeds[i % 8]))))
dl = [(D(f), Y(f)) for f in feeds]
print("column_ptrs x2  %.1f us" % med(lambda i: (column_ptrs(dl[i % 8][0], B), column_ptrs(dl[i % 8][1], B))))
n_docs = feeds[0]["letor_features"].shape[0]
total = lib.ub200_feed_bytes(n_docs, F, L, B)
pin = torch.empty(int(total * 1.5), dtype=torch.uint8, pin_memory=True)
dev = torch.empty(int(total * 1.5), dtype=torch.uint8, device="cuda")
cp = [(column_ptrs(d, B), column_ptrs(y, B)) for d, y in dl]
print("pack_ids_host   %.1f us" % med(lambda i: lib.ub200_pack_ids_host(cp[i % 8][0][0], cp[i % 8][1][0], L, B, n_docs, pin.data_ptr(), pin.numel())))
off_f = (8 * L * B + 255) // 256 * 256
for thr in (4, 8, 12, 16):
    print("convert rotating sources, %2d threads: %.1f us" % (thr, med(lambda i: lib.ub200_convert_f64_f32_host(
        feeds[i % 8]["letor_features"].ctypes.data, pin.data_ptr() + off_f, n_docs * F, thr))))
st = torch.cuda.current_stream().cuda_stream
for thr, grp in ((16, 6), (16, 12), (12, 6), (15, 6)):
    def call(i):
        f = feeds[i % 8]
        lib.ub200_stage_feed(f["letor_features"].ctypes.data, n_docs, F, cp[i % 8][0][0], cp[i % 8][1][0], L, B,
                             pin.data_ptr(), pin.numel(), dev.data_ptr(), thr, grp, st)
    def call_sync(i):
        call(i); torch.cuda.synchronize()
    print("stage_feed threads %2d groups %2d: call %.1f us, with sync %.1f us" % (thr, grp, med(call_sync, 100) * 0 + med(call, 100), med(call_sync, 100)))
print("torch.cuda.synchronize alone %.1f us" % med(lambda i: torch.cuda.synchronize()))
def memcpy_only(i):
    lib_rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), ctypes.c_void_p(pin.data_ptr()), ctypes.c_size_t(total), 1, ctypes.c_void_p(st))
try:
    lib_rt = ctypes.CDLL("libcudart.so.12")
    print("cudaMemcpyAsync call alone (5.6 MB) %.1f us (then sync %.1f)" % (med(memcpy_only, 50), med(lambda i: (memcpy_only(i), torch.cuda.synchronize()), 50)))
except Exception as e:
    print("no libcudart:", e)
