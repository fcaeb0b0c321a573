"""Timing aid (run on the GPU box): host packer thread scaling, H2D copy time, per-phase cost of train()."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ultra_pytorch_b200 import _capi, synth
lib = _capi.lib
F, L, B = 136, 40, 256
f = synth.make_feed(0, F, L, B)
feats = f["letor_features"]; n = feats.shape[0]
nb = lib.ub200_feed_bytes(n, F, L, B)
pin = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(nb, dtype=torch.uint8, device="cuda")
PtrArr = ctypes.c_void_p * L
d = [f["docid_input%d" % l] for l in range(L)]; y = [f["label%d" % l] for l in range(L)]
print("cpus", os.cpu_count())
for nt in (1, 2, 4, 8, 16):
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        dptr = PtrArr(*[x.ctypes.data for x in d]); lptr = PtrArr(*[x.ctypes.data for x in y])
        lib.ub200_pack_feed_host(feats.ctypes.data, n, F, dptr, lptr, L, B, pin.data_ptr(), nb, nt)
        ts.append(time.perf_counter() - t0)
    print("pack nt=%2d median %.3f ms min %.3f" % (nt, 1e3 * sorted(ts)[15], 1e3 * min(ts)))
hf = pin.numpy()[(8 * L * B + 255) // 256 * 256:].view(np.float32).reshape(n + 1, F)
ts = []
for _ in range(30):
    t0 = time.perf_counter(); np.copyto(hf[:n], feats, casting="same_kind"); ts.append(time.perf_counter() - t0)
print("numpy copyto median %.3f ms" % (1e3 * sorted(ts)[15]))
ts = []
for _ in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(pin, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("H2D %d bytes median %.3f ms (%.1f GB/s)" % (nb, 1e3 * sorted(ts)[15], nb / sorted(ts)[15] / 1e9))
# pipelined pack + H2D in one call, on 8 rotating (cache-cold) feeds
cold = [synth.make_feed(100 + i, F, L, B) for i in range(8)]
st = torch.cuda.current_stream().cuda_stream
for nt in (4, 8, 16):
    for ng in (1, 4, 6, 12):
        ts, te = [], []
        for it in range(40):
            fd = cold[it % 8]
            fe = fd["letor_features"]
            dd = [fd["docid_input%d" % l] for l in range(L)]; yy = [fd["label%d" % l] for l in range(L)]
            torch.cuda.synchronize(); t0 = time.perf_counter()
            dptr = PtrArr(*[x.ctypes.data for x in dd]); lptr = PtrArr(*[x.ctypes.data for x in yy])
            lib.ub200_stage_feed(fe.ctypes.data, n, F, dptr, lptr, L, B, pin.data_ptr(), nb, dev.data_ptr(), nt, ng, st)
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append(t1 - t0); te.append(t2 - t0)
        print("stage_feed nt=%2d groups=%2d: enqueue %.3f ms, data on device after %.3f ms" %
              (nt, ng, 1e3 * sorted(ts)[20], 1e3 * sorted(te)[20]))

# ---- per-phase cost of train() ----
import types, cProfile, pstats
import ultra_pytorch_b200.learning_algorithm as la
la.B200Algorithm.VERBOSE = False
torch.manual_seed(0)
model = la.IPWrank(types.SimpleNamespace(feature_size=F), synth.exp_settings("c2_ipw_mslr10k"))
feeds = [synth.make_feed(i, F, L, B) for i in range(4)]
for i in range(8):
    model.train(feeds[i % 4])
torch.cuda.synchronize()
N = 100
t_stage = t_step = t_read = 0.0
for i in range(N):
    fd = feeds[i % 4]
    t0 = time.perf_counter()
    st = model._stage(fd, L)
    t1 = time.perf_counter()
    out = model.run_step(st)
    t2 = time.perf_counter()
    s = model._read_scalars(out)
    t3 = time.perf_counter()
    t_stage += t1 - t0; t_step += t2 - t1; t_read += t3 - t2
print("per step: stage %.3f ms | launch %.3f ms | read-back (incl. GPU wait) %.3f ms | total %.3f ms" %
      (1e3 * t_stage / N, 1e3 * t_step / N, 1e3 * t_read / N, 1e3 * (t_stage + t_step + t_read) / N))
t0 = time.perf_counter()
for i in range(N):
    model.train(feeds[i % 4])
print("train() %.3f ms/step" % (1e3 * (time.perf_counter() - t0) / N))
pr = cProfile.Profile(); pr.enable()
for i in range(50):
    model.train(feeds[i % 4])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
