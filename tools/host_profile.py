"""Host-side breakdown of train(input_feed) at config 2: where do the wall-clock microseconds of a step go?"""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ultra_pytorch_b200 import synth
import ultra_pytorch_b200.learning_algorithm as la

la.B200Algorithm.VERBOSE = False
wl = sys.argv[1] if len(sys.argv) > 1 else "c2_ipw_mslr10k"
w = synth.WORKLOADS[wl]
F, L, B = w["F"], w["L"], w["B"]
settings = synth.exp_settings(wl)
model = getattr(la, w["algo"])(types.SimpleNamespace(feature_size=F), settings)
feeds = [synth.make_feed(i, F, L, B, w["labels"]) for i in range(8)]
for i in range(16):
    model.train(feeds[i % 8])
torch.cuda.synchronize()
N = 400
T = {k: [] for k in ("stage", "launch", "read", "total", "gpu_idle_wait")}
pc = time.perf_counter
for i in range(N):
    f = feeds[i % 8]
    t0 = pc()
    st = model._stage(f, model.rank_list_size)
    t1 = pc()
    out = model.run_step(st)
    t2 = pc()
    s = model._read_scalars(out)
    t3 = pc()
    T["stage"].append(t1 - t0); T["launch"].append(t2 - t1); T["read"].append(t3 - t2); T["total"].append(t3 - t0)
torch.cuda.synchronize()
for k in ("stage", "launch", "read", "total"):
    v = np.array(T[k]) * 1e6
    print("%-8s median %7.1f us   p10 %7.1f   p90 %7.1f" % (k, np.median(v), np.percentile(v, 10), np.percentile(v, 90)))
# the same loop through the public train()
t0 = pc()
for i in range(N):
    model.train(feeds[i % 8])
torch.cuda.synchronize()
print("train() loop: %.1f us/step" % ((pc() - t0) / N * 1e6))
import ctypes
from ultra_pytorch_b200 import _capi as _c
_c.lib.ub200_stage_timeline.argtypes = [ctypes.c_void_p]
tl = (ctypes.c_longlong * 32)()
acc = []
for i in range(200):
    model.train(feeds[i % 8])
    _c.lib.ub200_stage_timeline(tl)
    acc.append(list(tl))
a = np.median(np.array(acc), axis=0) / 1e3
print("inside ub200_stage_feed during train(): ids packed %.1f us | copies issued at %s | return %.1f us" % (
    a[0], " ".join("%.1f" % x for x in a[2:30] if x > 0), a[31]))
# stage only (no kernels): host pack + copies, synchronised each time
ts = []
for i in range(100):
    f = feeds[i % 8]
    t0 = pc()
    st = model._stage(f, model.rank_list_size)
    t1 = pc()
    torch.cuda.synchronize()
    t2 = pc()
    ts.append((t1 - t0, t2 - t0))
a = np.array(ts) * 1e6
print("stage alone: call returns after %.1f us, copies complete after %.1f us" % (np.median(a[:, 0]), np.median(a[:, 1])))
# pure H2D of the same bytes from pinned memory
nb = st.h2d_bytes
pin = torch.empty(nb, dtype=torch.uint8, pin_memory=True); dev = torch.empty(nb, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = pc()
for i in range(50):
    dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
dt = (pc() - t0) / 50
print("plain H2D of %d bytes: %.1f us (%.1f GB/s)" % (nb, dt * 1e6, nb / dt / 1e9))
# host conversion alone
from ultra_pytorch_b200 import _capi
lib = _capi.lib
src = feeds[0]["letor_features"]; dst = np.empty(src.shape, dtype=np.float32)
for thr in (1, 4, 8, 16):
    t0 = pc()
    for i in range(50):
        lib.ub200_convert_f64_f32_host(src.ctypes.data, dst.ctypes.data, src.size, thr)
    dt = (pc() - t0) / 50
    print("f64->f32 of %d values on %d threads: %.1f us (%.1f GB/s read)" % (src.size, thr, dt * 1e6, src.nbytes / dt / 1e9))
print("cpus", os.cpu_count(), "pack threads", model.engine._pack_threads, "chunks", model.engine._pack_chunks)
