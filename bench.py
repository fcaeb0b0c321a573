#!/usr/bin/env python
"""bench.py - queries/sec of the ULTRA training hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py --gpus 1 --steps 1000 --warmup 20                    # the B200 arm (this repo's kernels)
    python bench.py --impl reference --gpus 1 --steps 10 --warmup 3      # the reference's own CPU path
    torchrun --nproc-per-node N ... bench.py --gpus N ...                # data-parallel, one rank per GPU

A "step" = one `train()` of the workload's learning algorithm on one batch of B ranked lists (queries):
DNN forward + loss + backward + clip_grad_norm + Adagrad.  Workload at N=1: BASELINE.json configs[1]
(IPW + DNN[256,128,64], 136 features, list length 40, B = 256, PBM clicks) - `--workload` selects the others.

  value : whole-job queries/s with the input batches already resident in HBM (a ring of distinct resident batches
          larger than the 126 MB L2), timed with CUDA events, max over ranks.
  e2e   : the same metric through the plugin's public `train(input_feed)` with HOST numpy feeds: pinned pack +
          H2D copy of the batch + kernels + D2H read of the loss inside the timed region.
  roofline : the DNN forward+backward kernels (K1) timed alone with CUDA events; algorithmic FLOPs per SURVEY 8(d).
  cpu_baseline : the reference's CPU path (oracle/_ref, else the numpy oracle port) on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec (IPW+DNN, MSLR-30K 136-feat, list_len 40) at 1/2/4/8 B200 vs CPU ref"
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_ipw_mslr10k")
    ap.add_argument("--batch", type=int, default=0, help="override the batch size B (default: workload's, 256)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the informational feed + train leg")
    ap.add_argument("--cpu-steps", type=int, default=0, help="timed steps of the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-all-configs", action="store_true", help="skip the per-config / list-length table")
    ap.add_argument("--min-seconds", type=float, default=0.5,
                    help="the K timed steps are repeated (>= 5 times) until this much device time has been measured")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# CPU arm (reference / oracle port) - always a child process with CUDA hidden
# ----------------------------------------------------------------------------------------------------------
def run_cpu_arm(workload, batch, steps, warmup, timeout=900):
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "time_ref_cpu.py"), "--workload", workload,
           "--steps", str(steps), "--warmup", str(warmup)]
    if batch:
        cmd += ["--batch", str(batch)]
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        for line in reversed(out.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def cpu_steps_for(workload):
    # ~10-30 s of CPU work per sample (survey-time probes)
    return {"c2_ipw_mslr10k": 40, "c3_dla_yahoo": 20, "c4_lambdarank_mslr30k": 6, "c4_pairdebias_mslr30k": 1,
            "c5_dla_istella": 10, "c1_na_toy": 40}.get(workload, 10)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ultra_pytorch_b200 import synth
    w = dict(synth.WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    r = run_cpu_arm(args.workload, args.batch, args.steps, args.warmup, timeout=3000)
    if "error" in r:
        print(json.dumps({"impl": "reference", "unavailable": r["error"].replace("\n", " ")[-300:]}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": r["queries_per_s"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "algo": w["algo"], "features": w["F"], "list_len": w["L"],
                   "batch_queries": w["B"], "hidden": w["hidden"], "device": "host CPU"},
        "cpu_baseline": {"value": r["queries_per_s"], "unit": "queries/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["queries_per_s"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if r.get("feed_ms_per_step") is not None:
        tot = r["feed_ms_per_step"] + r["ms_per_step"]
        line["pipeline"] = {"what": "reference feed get_batch + train() per step (what main.py's loop runs)",
                            "unit": "queries/s", "value": round(w["B"] / (tot / 1e3), 1), "ms_per_step": round(tot, 3),
                            "feed_ms_per_step": r["feed_ms_per_step"]}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------
def main_b200(args):
    import types

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ultra_pytorch_b200 import _capi, synth
    import ultra_pytorch_b200.learning_algorithm as la

    la.B200Algorithm.VERBOSE = False
    if args.no_graph:
        la.B200Algorithm.USE_GRAPH = False
    w = dict(synth.WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    F, L, B, hidden = w["F"], w["L"], w["B"], w["hidden"]
    torch.manual_seed(0)
    settings = synth.exp_settings(args.workload)
    model = getattr(la, w["algo"])(types.SimpleNamespace(feature_size=F), settings)
    eng = model.engine

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build_ring(mdl, wl, n_host=8):
        """host feeds + a ring of distinct device-resident batches larger than the L2 (rank-specific seeds: every rank
        trains on its own shard of queries)"""
        e = mdl.engine
        Fq, Lq, Bq = wl["F"], wl["L"], wl["B"]
        fds = [synth.make_feed(1000 * rank + i, Fq, Lq, Bq, wl["labels"]) for i in range(n_host)]
        sb = e.stage(fds[0]["letor_features"], [fds[0]["docid_input%d" % l] for l in range(Lq)],
                     [fds[0]["label%d" % l] for l in range(Lq)]).h2d_bytes
        n = max(4, int(1.25 * L2_BYTES / sb) + 1)
        rg = []
        for i in range(n):
            f = fds[i % n_host]
            st = e.stage(f["letor_features"], [f["docid_input%d" % l] for l in range(Lq)],
                         [f["label%d" % l] for l in range(Lq)])
            torch.cuda.synchronize()
            own = e._dev[:st.h2d_bytes].clone()
            rg.append(e.staged_views(own, Lq, Bq, st.n_docs))
        return fds, rg, sb

    def time_steps(mdl, rg, steps, warmup, min_seconds, graphs):
        """K steps per repeat, timed on the device (CUDA events on the launching stream, barrier + synchronize on both
        sides, max over ranks); repeated >= 5 times and until min_seconds of device time have been measured; the MEDIAN
        repeat is reported.  Every ring slot has its CUDA graph before the first timed step."""
        n = len(rg)
        for i in range(max(warmup, 3)):
            mdl.run_step(rg[i % n])
        if graphs:
            for i in range(3 * n):
                mdl.run_step(rg[i % n])
        times, total, k0 = [], 0.0, 0
        while len(times) < 5 or total < min_seconds * 1e3:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for k in range(steps):
                mdl.run_step(rg[(k0 + k) % n])
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
            total += times[-1]
            k0 += steps
            if len(times) >= 200:
                break
        times.sort()
        return times[len(times) // 2], times, total

    feeds, ring, step_bytes = build_ring(model, w)
    ring_n, n_host = len(ring), len(feeds)
    use_graph = la.B200Algorithm.USE_GRAPH and (world == 1 or la.B200Algorithm.USE_GRAPH_DP)

    # ---- value leg: device-resident batches ----
    launches0 = _capi.lib.ub200_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, rep_times, timed_total_ms = time_steps(model, ring, args.steps, args.warmup, args.min_seconds, use_graph)
    eager_launches = _capi.lib.ub200_launch_count() - launches0
    value = world * B * args.steps / (ms / 1e3)

    # kernels per step (graph replays launch the captured kernels without passing through the library counter)
    model.USE_GRAPH_saved = la.B200Algorithm.USE_GRAPH
    la.B200Algorithm.USE_GRAPH = False
    c0 = _capi.lib.ub200_launch_count()
    model.run_step(ring[0])
    per_step = _capi.lib.ub200_launch_count() - c0
    la.B200Algorithm.USE_GRAPH = model.USE_GRAPH_saved
    gpu_launches = per_step * args.steps * len(rep_times) if use_graph else eager_launches

    # ---- roofline leg: K1 (DNN forward + backward kernels) timed alone, graph-replayed over the SAME ring of distinct
    # resident batches as the value leg (inputs larger than the L2, so the features come from HBM every time while the
    # weights stay cache-resident, exactly as inside a training step) ----
    def time_k1(mdl, rg, Bq, Lq, n_iter):
        e = mdl.engine
        dq = e.dscores_buf(Bq, Lq)
        graphs = []
        torch.cuda.synchronize()
        for stq in rg:
            gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf):
                e.forward(stq.feats, stq.docid.view(-1), Lq, Bq, training=True)
            with torch.cuda.graph(gb):
                e.backward(stq.feats, stq.docid.view(-1), Lq, Bq, dq)
            graphs.append((gf, gb))
        for gf, gb in graphs[:3]:
            gf.replay(); gb.replay()
        evs_ = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_iter)]
        torch.cuda.synchronize()
        for it in range(n_iter):
            gf, gb = graphs[it % len(graphs)]
            evs_[it][0].record(); gf.replay(); evs_[it][1].record(); gb.replay(); evs_[it][2].record()
        torch.cuda.synchronize()
        f_ms = sorted(x[0].elapsed_time(x[1]) for x in evs_)[n_iter // 2]
        b_ms = sorted(x[1].elapsed_time(x[2]) for x in evs_)[n_iter // 2]
        return f_ms + b_ms, f_ms

    k1_ms, fwd_ms = time_k1(model, ring, B, L, max(40, 2 * ring_n))
    flops = synth.train_flops_per_query(F, L, hidden) * B
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))

    def dense_peak(dtype, tf32):
        """torch.matmul 8192^3, best of 10 (the way MEASURED_PEAKS.json measures its bf16 number)"""
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a_ = torch.randn(8192, 8192, device="cuda", dtype=dtype)
        b_ = torch.randn(8192, 8192, device="cuda", dtype=dtype)
        best = 1e9
        for _ in range(12):
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            torch.matmul(a_, b_)
            q1.record()
            torch.cuda.synchronize()
            best = min(best, q0.elapsed_time(q1))
        torch.backends.cuda.matmul.allow_tf32 = False
        return 2.0 * 8192 ** 3 / (best / 1e3) / 1e12
    tf32_peak = dense_peak(torch.float32, True) if rank == 0 else None
    f16_peak = dense_peak(torch.float16, False) if rank == 0 else None
    traffic, traffic_src = None, None
    try:   # warm-cache dram__bytes_read.sum + dram__bytes_write.sum of the K1 kernels of ONE replayed step (ncu)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tr.get("%s_B%d" % (args.workload, B))
        if ent:
            traffic, traffic_src = ent["k1_dram_bytes_per_step"], tr.get("source")
    except Exception:  # noqa: BLE001
        pass
    achieved_tf = flops / (k1_ms / 1e3) / 1e12
    alg_bytes = int(4 * L * B * F)
    roofline = {"bound": "tensor", "achieved": round(achieved_tf, 3), "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(achieved_tf / peak_tf, 5), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_hbm_bytes": alg_bytes,
                "traffic_over_algorithmic": round(traffic / alg_bytes, 2) if traffic else None,
                "kernel": "K1 DNN forward+backward (all launches of ub200_mlp_forward + ub200_mlp_backward), median "
                          "CUDA-event duration over graph replays on the ring of distinct batches (> L2)",
                "ms_per_launch_group": round(k1_ms, 4), "fwd_ms": round(fwd_ms, 4),
                "algorithmic_flops": flops, "peak_source": "MEASURED_PEAKS.json bf16 burst" if peaks else "fallback",
                # the arithmetic is fp32-accurate through THREE fp16 products per multiply (x = hi + lo), so the ceiling
                # of this formulation is a third of the 16-bit dense rate; 3xTF32 (round 1) had a sixth
                "f16_dense_peak_measured": round(f16_peak, 1) if f16_peak else None,
                "tf32_dense_peak_measured": round(tf32_peak, 1) if tf32_peak else None,
                "frac_of_split_ceiling": round(achieved_tf / (peak_tf / 3.0), 5),
                "frac_of_tf32_peak_over_3": round(achieved_tf / (tf32_peak / 3.0), 5) if tf32_peak else None,
                "fp32_ffma_peak_tflops": 72.0, "frac_of_fp32_ffma_peak": round(achieved_tf / 72.0, 4)}

    # ---- e2e leg: public train(input_feed) with host feeds (wall clock around K train() calls + a final synchronize;
    # >= 5 repeats, median) ----
    for i in range(12):              # two staging buffer pairs, each seen three times before its CUDA graph exists
        model.train(feeds[i % n_host])
    e2e_times = []
    while len(e2e_times) < 5 or sum(e2e_times) < args.min_seconds:
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            model.train(feeds[k % n_host])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_times.append(float(t.item()))
        if len(e2e_times) >= 100:
            break
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (value leg .. e2e leg)
    e2e_s = sorted(e2e_times)[len(e2e_times) // 2]
    e2e = {"value": round(world * B * args.steps / e2e_s, 1), "unit": "queries/s",
           "h2d_bytes_per_step": int(model.last_h2d_bytes), "d2h_bytes_per_step": int(model.last_d2h_bytes),
           "ms_per_step": round(1e3 * e2e_s / args.steps, 4), "repeats": len(e2e_times)}

    # ---- pipeline leg (informational): what main.py's loop does per step - feed.get_batch() + model.train() - with
    # the drop-in ClickSimulationFeed on a synthetic data set, (a) in the reference's feed format (feature rows copied
    # per batch), (b) with the data set resident in HBM (input_layer/resident.py: only ids + labels move) and (c) with
    # query sampling + click simulation + batch assembly on the device as well (csrc/sampling.cu) ----
    pipeline = None
    if w["labels"] == "click" and not args.no_pipeline:
        import random as _random
        from ultra_pytorch_b200.input_layer import ClickSimulationFeed
        _random.seed(1234 + rank)
        ds = synth.synthetic_dataset(2048, L, F, seed=7 + rank)
        pipeline = {"what": "ClickSimulationFeed.get_batch + train() per step, synthetic data set of 2048 queries",
                    "unit": "queries/s"}
        n_pipe = max(200, min(args.steps, 1000))
        for name, hp in (("host_rows", ""), ("resident", "resident_features=True"), ("device", "device_batches=True")):
            try:
                feeder = ClickSimulationFeed(model, B, "click_model_json=%s,%s" % (synth.PBM_JSON, hp))
                for _ in range(24):          # the device feed rotates 4 output buffers and a graph needs 3 visits of each
                    model.train(feeder.get_batch(ds, check_validation=True)[0])
                barrier()
                t0 = time.perf_counter()
                for _ in range(n_pipe):
                    model.train(feeder.get_batch(ds, check_validation=True)[0])
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                pipeline[name] = {"value": round(world * B * n_pipe / dt, 1), "ms_per_step": round(1e3 * dt / n_pipe, 4),
                                  "h2d_bytes_per_step": int(model.last_h2d_bytes)}
            except Exception as exc:          # informational leg: never at the cost of the headline line (N = 1 only)
                if world > 1:
                    raise
                pipeline[name] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    # ---- the north_star table: every BASELINE.json config + the list-length sweep of the headline net, same method
    # (resident ring > L2, CUDA-graph replay, device-timed, median of >= 5 repeats, max over ranks), shorter runs ----
    all_configs = None
    if not args.no_all_configs:
        all_configs = []
        all_configs_error = None
        table = [(n, dict(synth.WORKLOADS[n])) for n in ("c1_na_toy", "c3_dla_yahoo", "c4_lambdarank_mslr30k",
                                                         "c4_pairdebias_mslr30k", "c5_dla_istella")]
        for Lq in (10, 20, 100, 200):
            table.append(("c2net_ipw_L%d" % Lq, dict(algo="IPWrank", F=136, L=Lq, B=256, hidden=[256, 128, 64],
                                                      labels="click")))
        for name, wq in table:
            if name == args.workload or all_configs_error is not None:
                continue
            try:
                torch.manual_seed(0)
                mq = getattr(la, wq["algo"])(types.SimpleNamespace(feature_size=wq["F"]), synth.exp_settings(wq))
                fq, rq, sbq = build_ring(mq, wq, n_host=4)
                steps_q = 50
                ms_q, reps_q, _ = time_steps(mq, rq, steps_q, 3, 0.1, use_graph)
                k1_q, _ = time_k1(mq, rq, wq["B"], wq["L"], max(16, len(rq)))     # K1 alone, same ring (inputs > L2)
                fl_q = synth.train_flops_per_query(wq["F"], wq["L"], wq["hidden"]) * wq["B"]
                tf_q = fl_q / (k1_q / 1e3) / 1e12
                all_configs.append({"workload": name, "algo": wq["algo"], "features": wq["F"], "list_len": wq["L"],
                                    "hidden": wq["hidden"], "batch_queries": wq["B"],
                                    "value": round(world * wq["B"] * steps_q / (ms_q / 1e3), 1), "unit": "queries/s",
                                    "ms_per_step": round(ms_q / steps_q, 5), "repeats": len(reps_q),
                                    "k1_ms": round(k1_q, 4), "k1_tflops": round(tf_q, 2),
                                    "k1_frac_of_bf16_peak": round(tf_q / peak_tf, 5),
                                    "k1_frac_of_split_ceiling": round(tf_q / (peak_tf / 3.0), 5)})
                del mq, fq, rq
                torch.cuda.empty_cache()
            except Exception as exc:          # an informational leg must not cost the headline line (single process only:
                if world > 1:                 # under torchrun every rank has to take the same path through the collectives)
                    raise
                all_configs_error = "%s: %s" % (type(exc).__name__, str(exc)[:300])
                all_configs.append({"workload": name, "error": all_configs_error})

    # ---- data-parallel self-check (N > 1): replicas must stay bitwise equal, and the sharded step must equal a
    # single-GPU step on the merged batch (rank 0 re-runs the merged batches on one GPU) ----
    dp_check = None
    if world > 1:
        import numpy as np
        torch.manual_seed(0)
        mdp = getattr(la, w["algo"])(types.SimpleNamespace(feature_size=F), settings)
        init = {k: v.clone() for k, v in mdp.model.state_dict().items()}
        Bc = 64
        for step in range(3):
            mdp.train(synth.make_feed(7000 + 100 * step + rank, F, L, Bc, w["labels"]))
        flat = mdp.engine.params.clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g_) for g_ in gathered)
        err = None
        err_stats = None
        if rank == 0:
            ws_fn = la.B200Algorithm.world_size
            la.B200Algorithm.world_size = staticmethod(lambda: 1)      # a plain single-GPU model (no rendezvous)
            try:
                torch.manual_seed(0)
                ref = getattr(la, w["algo"])(types.SimpleNamespace(feature_size=F), settings)
                ref.model.load_state_dict(init)
                for step in range(3):
                    fs = [synth.make_feed(7000 + 100 * step + r, F, L, Bc, w["labels"]) for r in range(world)]
                    feats_m = np.concatenate([f["letor_features"] for f in fs], axis=0)
                    merged = {"letor_features": feats_m}
                    for l in range(L):
                        d_, y_, base = [], [], 0
                        for f in fs:
                            n_ = f["letor_features"].shape[0]
                            di = f["docid_input%d" % l].astype(np.int64)
                            d_.append(np.where(di == n_, feats_m.shape[0], di + base))
                            y_.append(f["label%d" % l])
                            base += n_
                        merged["docid_input%d" % l] = np.concatenate(d_).astype(np.float32)
                        merged["label%d" % l] = np.concatenate(y_).astype(np.float32)
                    ref.train(merged)
                a_, b_ = flat, ref.engine.params
                keep = torch.ones_like(a_, dtype=torch.bool)       # mathematically-zero gradients random-walk
                for nm, off, shape in mdp.engine.layer_slices():
                    if nm in ("layer_norm%d.bias" % len(hidden), "linear%d.bias" % len(hidden)):
                        keep[off:off + int(np.prod(shape))] = False
                rel = ((a_ - b_)[keep].abs() / b_[keep].abs().mean()).float()
                err = float(rel.max())
                err_stats = {"median": float(rel.median()), "p999": float(torch.quantile(rel, 0.999)),
                             "entries_over_1e-4": int((rel > 1e-4).sum()), "entries": int(rel.numel())}
            finally:
                la.B200Algorithm.world_size = ws_fn
        dist.barrier()
        dp_check = {"replicas_bitwise_equal": bool(same), "vs_single_gpu": err, "vs_single_gpu_distribution": err_stats,
                    "what": "3 train() steps of %d queries per rank; vs_single_gpu = max|param_dp - param_single| / "
                            "mean|param| against one GPU training on the merged batches (the shards are summed in a "
                            "different order than the merged batch, and Adagrad's first steps move a parameter by "
                            "lr * g / |g|: entries whose gradient is rounding noise flip sign - the distribution "
                            "shows how few they are)" % Bc}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "algo": w["algo"], "features": F, "list_len": L,
                       "batch_queries": B, "hidden": hidden, "parallelism": "dp%d" % world,
                       "cuda_graph": bool(use_graph),
                       "l2": "ring of %d distinct resident batches (%.0f MB) > 126 MB L2" %
                             (ring_n, ring_n * step_bytes / 1e6)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "kernels_per_step": int(per_step),
            "roofline": roofline,
            "repeats": len(rep_times), "timed_region_s": round(timed_total_ms / 1e3, 4),
            "timing": "value = K steps / the MEDIAN of `repeats` device-timed repetitions of the K-step loop "
                      "(min %.3f ms, max %.3f ms per repetition)" % (min(rep_times), max(rep_times)),
        }
        if all_configs is not None:
            line["all_configs"] = all_configs
        if dp_check is not None:
            line["dp_check"] = dp_check
        if pipeline is not None:
            line["pipeline"] = pipeline
        if world == 1 and not args.no_cpu_baseline:
            n = args.cpu_steps or cpu_steps_for(args.workload)
            r = run_cpu_arm(args.workload, args.batch, n, 3)
            if "error" in r:
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": None, "kind": "unavailable",
                                        "sample": r["error"][-200:]}
            else:
                line["cpu_baseline"] = {"value": r["queries_per_s"], "unit": "queries/s", "cores": r["cores"],
                                        "kind": r["kind"], "sample": r["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _JsonOnlyStdout(object):
    """stdout carries exactly ONE line, the JSON result.  Everything the plugin classes print while they are being
    constructed (they mirror the reference's prints) goes to stderr, and so does whatever C libraries write to file
    descriptor 1 directly (NCCL prints its version there): fd 1 is pointed at stderr for the whole run and the JSON
    line is written to a duplicate of the original stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        self.last = sys.stderr

    def write(self, text):
        if text.strip():
            t = text.lstrip()
            self.last = self.real if (t.startswith('{"metric"') or t.startswith('{"impl"')) else sys.stderr
        self.last.write(text)
        return len(text)

    def flush(self):
        self.real.flush()
        sys.stderr.flush()


if __name__ == "__main__":
    a = parse_args()
    sys.stdout = _JsonOnlyStdout()
    try:
        if a.impl == "reference":
            main_reference(a)
        else:
            main_b200(a)
    finally:
        sys.stdout.flush()
