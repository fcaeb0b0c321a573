"""Turns the ncu exports of a gpurun call into the markdown tables kept under profiles/ (run here, after the call).

    python tools/summarize_ncu.py <raw_full.csv> <launches.csv> [<raw_k3.csv>]  > profiles/rNN_tables.md
"""
import csv
import sys
from collections import OrderedDict

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum"]
SHORT = ["time us", "dram rd", "dram wr", "tensor pipe %", "SM thr %", "DRAM thr %", "warps act %", "issue act %", "regs",
         "warp inst"]


def short(name):
    return name.replace("ub200::", "").replace("tc::", "").split("(")[0][:44]


def full_table(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(m) if m in hdr else None for m in FULL]
    ik, ig, ib = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    print("| kernel | grid x block | " + " | ".join(SHORT) + " |")
    print("|---|---|" + "---|" * len(SHORT))
    tot_rd = tot_wr = 0.0
    for r in data:
        cells = []
        for m, i in zip(FULL, idx):
            if i is None:
                cells.append("-")
                continue
            v, u = r[i], units[i]
            try:
                f = float(v)
            except ValueError:
                cells.append(v)
                continue
            if m.startswith("dram__bytes"):
                mb = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                cells.append("%.2f MB" % mb)
                if m.endswith("read.sum"):
                    tot_rd += mb
                else:
                    tot_wr += mb
            elif m == "smsp__inst_executed.sum":
                cells.append("%.2fM" % (f / 1e6))
            elif m == "launch__registers_per_thread":
                cells.append("%d" % f)
            else:
                cells.append("%.1f" % f)
        print("| %s | %s x %s | %s |" % (short(r[ik]), r[ig], r[ib], " | ".join(cells)))
    return tot_rd, tot_wr


def launch_table(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = OrderedDict()
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        key = (short(x["Kernel Name"]), x["Grid Size"], x["Block Size"])
        v = float(x["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(x["Metric Unit"], 1.0)
        agg.setdefault(key, []).append(v)
    print("| kernel | grid x block | launches captured | avg us |")
    print("|---|---|---|---|")
    per_step = 0.0
    n_steps = max(len(v) for v in agg.values())
    for (k, g, b), v in agg.items():
        print("| %s | %s x %s | %d | %.1f |" % (k, g, b, len(v), sum(v) / len(v)))
        per_step += sum(v) / n_steps
    print("\nserialised sum per step: %.1f us over %d captured steps" % (per_step, n_steps))


if __name__ == "__main__":
    print("## `ncu --set full` (default cache control: caches flushed before every kernel -> cold-cache bytes / times)\n")
    rd, wr = full_table(sys.argv[1])
    print("\nDRAM bytes of the captured step, all kernels: read %.1f MB, write %.1f MB" % (rd, wr))
    print("\n## launch list, warm caches (`--cache-control none`)\n")
    launch_table(sys.argv[2])
    if len(sys.argv) > 3:
        print("\n## K3 pairwise kernel (config 4: LambdaRank, B = 256, L = 200), `ncu --set full`\n")
        full_table(sys.argv[3])
