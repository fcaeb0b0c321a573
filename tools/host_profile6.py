import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ultra_pytorch_b200 import synth, _capi
from ultra_pytorch_b200 import engine as E
import ultra_pytorch_b200.learning_algorithm as la
la.B200Algorithm.VERBOSE = False
w = synth.WORKLOADS["c2_ipw_mslr10k"]
F, L, B = w["F"], w["L"], w["B"]
model = la.IPWrank(types.SimpleNamespace(feature_size=F), synth.exp_settings("c2_ipw_mslr10k"))
feeds = [synth.make_feed(i, F, L, B, w["labels"]) for i in range(8)]
for i in range(16):
    model.train(feeds[i % 8])
torch.cuda.synchronize()
eng = model.engine
pc = time.perf_counter
def med(f, n=300):
    ts = []
    for i in range(n):
        t0 = pc(); f(i); ts.append(pc() - t0)
    return np.median(ts) * 1e6
g = model._feed_getters[L]
print("getters           %.1f" % med(lambda i: (g[0](feeds[i % 8]), g[1](feeds[i % 8]))))
dl = [(g[0](f), g[1](f)) for f in feeds]
print("np.asarray+shape  %.1f" % med(lambda i: (np.asarray(feeds[i % 8]["letor_features"]).shape, isinstance(feeds[i % 8]["letor_features"], E._resident_features_cls()))))
print("_stream           %.1f" % med(lambda i: E._stream()))
cs = E._stream()
print("_acquire_slot     %.1f" % med(lambda i: eng._acquire_slot(5653024, cs)))
print("column_ptrs x2    %.1f" % med(lambda i: (E.column_ptrs(dl[i % 8][0], B), E.column_ptrs(dl[i % 8][1], B))))
print("np.array(40 arrays) %.1f" % med(lambda i: np.array(dl[i % 8][0], dtype=np.float32)))
print("np.stack(40 arrays) %.1f" % med(lambda i: np.stack(dl[i % 8][0])))
out = np.empty((L, B), dtype=np.float32)
print("np.stack(out=)      %.1f" % med(lambda i: np.stack(dl[i % 8][0], out=out)))
def fast(x):
    return isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags.c_contiguous and x.shape == (B,)
print("_f32_vec x4       %.1f" % med(lambda i: (fast(dl[0][0][0]), fast(dl[0][0][-1]), fast(dl[0][1][0]), fast(dl[0][1][-1]))))
feats = feeds[0]["letor_features"]
print("array_interface   %.1f" % med(lambda i: feats.__array_interface__['data'][0]))
print("feats flags       %.1f" % med(lambda i: (feats.dtype == np.float64 and feats.flags.c_contiguous and feats.shape[1] == F)))
slot = eng._slots[0]
print("staged_views      %.1f" % med(lambda i: eng.staged_views(slot.dev, L, B, 10240)))
print("whole _stage      %.1f" % med(lambda i: (model._stage(feeds[i % 8], L), torch.cuda.synchronize())[0]))
print("whole _stage nosync %.1f" % med(lambda i: model._stage(feeds[i % 8], L), 50))
