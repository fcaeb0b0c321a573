"""TEST / BENCH INFRASTRUCTURE ONLY: times the reference's own CPU implementation of the hot path.

Run as a child process with CUDA_VISIBLE_DEVICES="" (the reference picks its device at import time,
SURVEY.md 0.4).  kind = "reference": the UNMODIFIED reference copy in oracle/_ref driven through its own public API
(ClickSimulationFeed.get_batch -> <Algorithm>.train) on a seeded synthetic Raw_data of the workload's shape;
kind = "port": the numpy oracle (oracle/ultra_oracle.py) when oracle/_ref is absent.
Prints one JSON line: queries_per_s (train() only, pre-built feeds), ms_per_step, cores, kind, sample.
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2_ipw_mslr10k")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    from ultra_pytorch_b200 import synth          # workload table only (pure python, no CUDA)
    w = dict(synth.WORKLOADS[a.workload])
    if a.batch:
        w["B"] = a.batch
    F, L, B, hidden = w["F"], w["L"], w["B"], w["hidden"]
    cores = a.threads or os.cpu_count()
    torch.set_num_threads(cores)
    from oracle import ref_shim
    n_feeds = 4
    if ref_shim.available():
        kind = "reference"
        ultra = ref_shim.load()
        random.seed(0)
        np.random.seed(0)
        torch.manual_seed(0)
        ds = ref_shim.synthetic_raw_data(ultra, 512, L, F, seed=1, max_label=4)
        settings = synth.exp_settings(a.workload)
        settings["learning_algorithm"] = "ultra.learning_algorithm.%s" % w["algo"]
        settings["ranking_model"] = "ultra.ranking_model.DNN"
        settings.update({"train_input_feed": "ultra.input_layer.ClickSimulationFeed", "train_input_hparams": ""})
        if w["labels"] != "click":
            settings["train_input_feed"] = "ultra.input_layer.DirectLabelFeed"
        with ref_shim.ref_cwd():
            import contextlib
            import io
            sink = io.StringIO()
            with contextlib.redirect_stdout(sink):
                model = ultra.utils.find_class(settings["learning_algorithm"])(ds, settings)
                feed_cls = ultra.utils.find_class(settings["train_input_feed"])
                feeder = feed_cls(model, B, settings["train_input_hparams"])
                feeds = [feeder.get_batch(ds, check_validation=True)[0] for _ in range(n_feeds)]
                for i in range(a.warmup):
                    model.train(dict(feeds[i % n_feeds]))
                t0 = time.perf_counter()
                for i in range(a.steps):
                    model.train(dict(feeds[i % n_feeds]))
                dt = time.perf_counter() - t0
                # what main.py's loop pays per step besides train(): the reference feed's get_batch (main.py:154)
                n_fb = 3
                t1 = time.perf_counter()
                for _ in range(n_fb):
                    feeder.get_batch(ds, check_validation=True)
                feed_ms = 1e3 * (time.perf_counter() - t1) / n_fb
        nb = len(feeds[0][model.labels_name[0]])
        assert nb == B or w["labels"] != "click", (nb, B)
        B = nb
    else:
        feed_ms = None
        kind = "port"
        from oracle import ultra_oracle as uo
        rs = np.random.RandomState(0)
        params = {}
        for j, (k, n) in enumerate(uo.layer_sizes(F, hidden)):
            params["sequential.layer_norm%d.weight" % j] = np.ones(k, np.float32)
            params["sequential.layer_norm%d.bias" % j] = np.zeros(k, np.float32)
            params["sequential.linear%d.weight" % j] = (rs.uniform(-1, 1, (n, k)) / np.sqrt(k)).astype(np.float32)
            params["sequential.linear%d.bias" % j] = np.zeros(n, np.float32)
        algo = {"IPWrank": "ipw", "DLA": "dla", "LambdaRank": "lambdarank", "PairDebias": "pairdebias",
                "NavieAlgorithm": "na"}[w["algo"]]
        table = json.load(open(synth.IPW_JSON))["IPW_list"]
        prop = {"linear_layer.weight": 0.1 * rs.randn(1, L), "linear_layer.bias": np.zeros(1)}
        tr = uo.OracleTrainer(algo, params, F, hidden, L, ipw_table=table, prop_params=prop, dt=np.float32)
        feeds = []
        for i in range(n_feeds):
            f = synth.make_feed(i, F, L, B, w["labels"])
            d = np.stack([f["docid_input%d" % l] for l in range(L)], axis=1).astype(np.int64)
            y = np.stack([f["label%d" % l] for l in range(L)], axis=1)
            feeds.append((f["letor_features"], d, y))
        cores = 1
        for i in range(a.warmup):
            tr.train(*feeds[i % n_feeds])
        t0 = time.perf_counter()
        for i in range(a.steps):
            tr.train(*feeds[i % n_feeds])
        dt = time.perf_counter() - t0
    cpu = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                cpu = ln.split(":", 1)[1].strip()
                break
    except Exception:  # noqa: BLE001
        pass
    print(json.dumps({
        "queries_per_s": round(B * a.steps / dt, 2), "ms_per_step": round(1e3 * dt / a.steps, 3), "cores": cores,
        "kind": kind, "cpu": cpu, "feed_ms_per_step": None if feed_ms is None else round(feed_ms, 3),
        "sample": "%d timed train() steps (+%d warm-up) of %s B=%d L=%d F=%d DNN%s on %d host threads (%s), "
                  "pre-built feeds, torch %s" % (a.steps, a.warmup, w["algo"], B, L, F, hidden, cores, cpu,
                                                 torch.__version__)}))


if __name__ == "__main__":
    main()
