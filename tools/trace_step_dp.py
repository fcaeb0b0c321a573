"""tools/trace_step.py under torchrun (data parallel): rank 0 prints the kernel timeline of one replayed step.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/trace_step_dp.py"""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from ultra_pytorch_b200 import synth
import ultra_pytorch_b200.learning_algorithm as la

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_ipw_mslr10k"
la.B200Algorithm.VERBOSE = False
w = synth.WORKLOADS[wl]
F, L, B = w["F"], w["L"], w["B"]
torch.manual_seed(0)
model = getattr(la, w["algo"])(types.SimpleNamespace(feature_size=F), synth.exp_settings(wl))   # joins the NCCL group
rank = dist.get_rank()
eng = model.engine
f = synth.make_feed(rank, F, L, B, w["labels"])
st = eng.stage(f["letor_features"], [f["docid_input%d" % l] for l in range(L)], [f["label%d" % l] for l in range(L)])
for _ in range(8):
    model.run_step(st)
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(8):
        model.run_step(st)
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    path = "gpurun_out/trace_dp_%s.json" % wl
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    first = ev[0]["name"]
    starts = [i for i, e in enumerate(ev) if e["name"] == first]
    a, b = starts[4], starts[5]
    t0 = ev[a]["ts"]
    print("N = %d: step of %d kernels, %.1f us from first start to next step's first start" % (dist.get_world_size(), b - a, ev[b]["ts"] - t0))
    last_end = t0
    for e in ev[a:b]:
        print("%8.1f +%6.1f us  gap-after-prev-end %6.1f  stream %s  %s" %
              (e["ts"] - t0, e["dur"], e["ts"] - last_end, e["args"].get("stream"), e["name"][:70]))
        last_end = max(last_end, e["ts"] + e["dur"])
dist.destroy_process_group()
