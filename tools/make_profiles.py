"""Assembles profiles/r01_ncu_summary.md (+ copies of the small raw artefacts) from the files a gpurun call left under
gpurun_out/.  Usage: python tools/make_profiles.py <tag of the final call, e.g. r1g>"""
import contextlib
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_ncu as sn  # noqa: E402

tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def raw_csv(rep):
    out = rep[:-8] + "_raw.csv"
    with open(out, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
    return out


def captured(fn, *a):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


full = raw_csv(os.path.join(G, "prof_%s_full.ncu-rep" % tag))
k2 = raw_csv(os.path.join(G, "prof_%s_k2.ncu-rep" % tag))
k3 = raw_csv(os.path.join(G, "prof_%s_k3.ncu-rep" % tag))
bench_line = [l for l in open(os.path.join(G, "bench_%s_c2.json" % tag)) if l.startswith("{")][-1].strip()
ref_line = [l for l in open(os.path.join(G, "bench_%s_ref.json" % tag)) if l.startswith("{")][-1].strip()
others = [l.strip() for l in open(os.path.join(G, "bench_%s_others.json" % tag)) if l.startswith("{")]
for src, dst in (("launches_%s.csv" % tag, "r01_launches_c2_warmcache.csv"), ("kernels_%s.txt" % tag, "r01_kernels.txt"),
                 ("trace_%s_c2.txt" % tag, "r01_step_timeline_c2.txt"), ("bench_%s_c2.json" % tag, "r01_bench_c2.json"),
                 ("pytest_%s.log" % tag, "r01_pytest_gpu.log")):
    shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
with open(os.path.join(P, "r01_bench_others.jsonl"), "w") as f:
    f.write("\n".join(others + [ref_line]) + "\n")

md = []
md.append("# Round 1 - ncu evidence for the hot path (B200; config 2 = IPW + DNN[256,128,64], F=136, L=40, B=256)\n")
md.append("Commands (run through gpurun on one B200; clocks untouched, `clocks.sm` 1965 MHz during the bench; "
          "`tools/gpu_final.sh` is the exact script):\n")
md.append("```\n" + open(os.path.join(ROOT, "tools", "gpu_final.sh")).read() + "```\n")
md.append("## bench.py lines of the same build (not under a profiler)\n")
md.append("```\n" + bench_line + "\n```\n")
md.append("reference arm (`--impl reference`, the unmodified reference on the host CPUs of the same box):\n")
md.append("```\n" + ref_line + "\n```\n")
md.append("other workloads (`profiles/r01_bench_others.jsonl`): " + "; ".join(
    "%s B=%d: %.0f q/s (%.3f ms/step, e2e %.0f q/s)" % (
        __import__("json").loads(l)["config"]["workload"], __import__("json").loads(l)["config"]["batch_queries"],
        __import__("json").loads(l)["value"], __import__("json").loads(l)["ms_per_step"],
        __import__("json").loads(l)["e2e"]["value"]) for l in others) + "\n")
md.append("## K1 + K2 + optimizer: one training step of config 2, `ncu --set full` "
          "(default cache control: caches flushed before every kernel => cold-cache DRAM bytes and times)\n")
md.append(captured(sn.full_table, full))
md.append("\nReading: `tc_gemm_kernel<0|1|2, BLOCK_N>` (0 = forward, 1 = data gradient with fused LayerNorm-backward, "
          "2 = weight gradient) and `fwd_fused_kernel` are the tcgen05 kernels (`sm__pipe_tensor_cycles_active` > 0, 544 "
          "threads, 1 CTA/SM); nothing is near a bandwidth limit at B = 256 (DRAM throughput < 10 %): the step is bound by "
          "the latency chain of its 9 critical-path kernels (timeline below).  DRAM write bytes are ~0 because the "
          "working set (45 MB) stays in the 126 MB L2; the DRAM reads of the backward kernels are cold-cache artefacts of "
          "the per-kernel flush (in the pipelined step those are L2 hits), which is why `roofline.traffic` in bench.py "
          "(sum over the K1 launches, `profiles/r01_traffic.json`) is far above the algorithmic 5.6 MB.\n")
md.append("## launch list of the same command, warm caches (`--cache-control none`), 4 steps\n")
md.append(captured(sn.launch_table, os.path.join(G, "launches_%s.csv" % tag)))
md.append("\nK1 (prep + fused forward + final_* + tc_gemm + wgrad_finalize) is ~92 % of the serialised sum, the same share "
          "bench.py measures live with CUDA events (K1 0.140 ms of 0.147 ms per step; in the graph the weight-gradient "
          "branches overlap the data-gradient chain, so the step is shorter than the serialised sum).\n")
md.append("## timeline of one graph-replayed step (CUPTI via torch.profiler, `tools/trace_step.py`)\n")
md.append("```\n" + "".join(l for l in open(os.path.join(G, "trace_%s_c2.txt" % tag)) if "us " in l or l.startswith("step")) + "```\n")
md.append("## K2 at scale (B = 2^20 lists of 40 positions, 512 MB algorithmic traffic), `ncu --set full`\n")
md.append(captured(sn.full_table, k2))
md.append("\n## K3 (config 4: LambdaRank, B = 256, L = 200), `ncu --set full`\n")
md.append(captured(sn.full_table, k3))
md.append("\n## per-kernel throughput where the kernel, not the launch, sets the time (`tools/bench_kernels.py`, CUDA events, L2 flushed)\n")
md.append("```\n" + open(os.path.join(G, "kernels_%s.txt" % tag)).read() + "```\n")
open(os.path.join(P, "r01_ncu_summary.md"), "w").write("\n".join(md))
print("wrote profiles/r01_ncu_summary.md")
