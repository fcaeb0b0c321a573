import os, sys, time, types, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ultra_pytorch_b200 import synth
import ultra_pytorch_b200.learning_algorithm as la
la.B200Algorithm.VERBOSE = False
w = synth.WORKLOADS["c2_ipw_mslr10k"]
F, L, B = w["F"], w["L"], w["B"]
model = la.IPWrank(types.SimpleNamespace(feature_size=F), synth.exp_settings("c2_ipw_mslr10k"))
feeds = [synth.make_feed(i, F, L, B, w["labels"]) for i in range(8)]
for i in range(16):
    model.train(feeds[i % 8])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(500):
    model.train(feeds[i % 8])
pr.disable()
torch.cuda.synchronize()
ps = pstats.Stats(pr).sort_stats("tottime")
ps.print_stats(22)
