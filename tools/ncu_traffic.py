"""Post-processes the ncu CSV of tools/ncu_step.py (metrics dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum; --cache-control none) into profiles/r02_traffic.json + a per-kernel table of the LAST step.
Usage: python tools/ncu_traffic.py gpurun_out/ncu_step_c2.csv c2_ipw_mslr10k 256 4"""
import csv, json, os, sys
path, wl, B, n_steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = []
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
# one record per (launch id): collect metrics
launch = {}
order = []
for r in rows:
    i = int(r["ID"])
    if i not in launch:
        launch[i] = {"name": r["Kernel Name"], "m": {}}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit in ("Kbyte", "KB"): v *= 1e3
    if unit in ("Mbyte", "MB"): v *= 1e6
    if unit in ("Gbyte", "GB"): v *= 1e9
    if unit in ("usecond", "us"): v *= 1e3
    if unit in ("msecond", "ms"): v *= 1e6
    if unit in ("second", "s"): v *= 1e9
    launch[i]["m"][r["Metric Name"]] = v
per_step = len(order) // n_steps
last = order[-per_step:]
K1 = ("prep16", "fwd16", "bwd16", "wgrad16", "final_bwd", "final_finalize", "wgrad_finalize", "final_fwd", "tc_gemm",
      "fwd_fused", "prep_weights", "row_stats", "ln_bwd")
tot = 0.0
print("%-60s %10s %10s %10s" % ("kernel (last of %d eager steps, warm caches)" % n_steps, "us", "rd KB", "wr KB"))
for i in last:
    m = launch[i]["m"]
    rd, wr, ns = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0), m.get("gpu__time_duration.sum", 0.0)
    nm = launch[i]["name"]
    if any(k in nm for k in K1):
        tot += rd + wr
    print("%-60s %10.1f %10.1f %10.1f" % (nm[:60], ns / 1e3, rd / 1e3, wr / 1e3))
print("K1 DRAM bytes per step (read + write): %.0f" % tot)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
d = json.load(open(out)) if os.path.isfile(out) else {}
d["source"] = ("ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over "
               "tools/ncu_step.py (eager steps on one resident batch, caches NOT flushed between kernels); sum over the K1 "
               "kernels (prep, forward, data-gradient chain, weight gradients, finalize) of the last step")
d["%s_B%d" % (wl, B)] = {"k1_dram_bytes_per_step": int(tot)}
json.dump(d, open(out, "w"), indent=1)
