"""Runs a few EAGER training steps of one workload (no CUDA graph, one resident batch) so that ncu can attribute DRAM
traffic / durations to the individual kernels of a step with warm caches (--cache-control none).
Usage: python tools/ncu_step.py [workload] [n_steps] [batch]"""
import os, sys, types
os.environ["UB200_GRAPH"] = "0"
os.environ["UB200_EARLY_LOSS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultra_pytorch_b200 import synth
import ultra_pytorch_b200.learning_algorithm as la

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_ipw_mslr10k"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
la.B200Algorithm.VERBOSE = False
la.B200Algorithm.USE_GRAPH = False
w = dict(synth.WORKLOADS[wl])
if len(sys.argv) > 3:
    w["B"] = int(sys.argv[3])            # batch override (e.g. 16384 for the large-M captures)
F, L, B = w["F"], w["L"], w["B"]
torch.manual_seed(0)
model = getattr(la, w["algo"])(types.SimpleNamespace(feature_size=F), synth.exp_settings(w))
f = synth.make_feed(0, F, L, B, w["labels"])
st = model.engine.stage(f["letor_features"], [f["docid_input%d" % l] for l in range(L)], [f["label%d" % l] for l in range(L)])
for _ in range(n):
    model.run_step(st)
torch.cuda.synchronize()
