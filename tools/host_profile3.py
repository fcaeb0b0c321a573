"""Does the H2D copy engine make progress while the host threads convert?  (H2D of one pinned buffer timed alone and
while another thread keeps the f64 -> f32 conversion running on other buffers.)"""
import os, sys, time, threading, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ultra_pytorch_b200 import _capi
lib = _capi.lib
n = 10241 * 136
srcs = [np.random.rand(n) for _ in range(8)]
dst_pin = torch.empty(n * 4, dtype=torch.uint8, pin_memory=True)
pin = torch.empty(n * 4, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
pc = time.perf_counter
def h2d(reps=30, chunks=1):
    ts = []
    step = (n * 4 // chunks + 255) // 256 * 256
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = pc()
        for c in range(chunks):
            lo, hi = c * step, min(n * 4, (c + 1) * step)
            dev[lo:hi].copy_(pin[lo:hi], non_blocking=True)
        torch.cuda.synchronize()
        ts.append(pc() - t0)
    return np.median(ts) * 1e6
print("H2D alone: %.1f us (1 copy), %.1f us (6 copies)" % (h2d(), h2d(chunks=6)))
for thr in (2, 4, 8, 12, 16):
    stop = False
    count = [0]
    def conv():
        i = 0
        while not stop:
            lib.ub200_convert_f64_f32_host(srcs[i % 8].ctypes.data, dst_pin.data_ptr(), n, thr)
            i += 1
        count[0] = i
    th = threading.Thread(target=conv)
    t0 = pc()
    th.start()
    time.sleep(0.05)
    t = h2d()
    t6 = h2d(chunks=6)
    stop = True
    th.join()
    el = pc() - t0
    print("conversion on %2d threads running (%.1f us per conversion): H2D %.1f us (1 copy), %.1f us (6 copies)" % (
        thr, el / max(1, count[0]) * 1e6, t, t6))
