"""Kernel timeline of graph-replayed training steps (run on the GPU box): torch.profiler (CUPTI) start/end of every
kernel inside the replayed CUDA graph -> gaps and overlaps on the critical path.  Prints one table; the chrome trace
goes to gpurun_out/."""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from ultra_pytorch_b200 import synth
import ultra_pytorch_b200.learning_algorithm as la

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_ipw_mslr10k"
la.B200Algorithm.VERBOSE = False
w = synth.WORKLOADS[wl]
F, L, B = w["F"], w["L"], w["B"]
torch.manual_seed(0)
model = getattr(la, w["algo"])(types.SimpleNamespace(feature_size=F), synth.exp_settings(wl))
eng = model.engine
f = synth.make_feed(0, F, L, B, w["labels"])
st = eng.stage(f["letor_features"], [f["docid_input%d" % l] for l in range(L)], [f["label%d" % l] for l in range(L)])
for _ in range(6):
    model.run_step(st)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        model.run_step(st)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/trace_%s.json" % wl
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# steps are separated by the prep kernel (first kernel of a step)
first = ev[0]["name"]
starts = [i for i, e in enumerate(ev) if e["name"] == first]
if len(starts) >= 4:
    a, b = starts[2], starts[3]
    t0 = ev[a]["ts"]
    print("step %d kernels, %.1f us from first start to next step's first start" % (b - a, ev[b]["ts"] - t0))
    last_end = t0
    for e in ev[a:b]:
        print("%8.1f +%6.1f us  gap-after-prev-end %6.1f  stream %s  %s" %
              (e["ts"] - t0, e["dur"], e["ts"] - last_end, e["args"].get("stream"), e["name"][:70]))
        last_end = max(last_end, e["ts"] + e["dur"])
