"""Per-kernel roofline numbers for K2 / K3 at sizes where the kernel, not the launch, sets the time (run on the GPU
box).  K2 (listwise softmax cross-entropy, HBM-bound): algorithmic bytes/list = 12*L + 8.  K3 (pairwise,
issue-bound): L*(L-1)/2 unordered pair evaluations per list.  CUDA events, L2 flushed between iterations."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ultra_pytorch_b200.engine import RankerEngine

peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6500.0))
eng = RankerEngine(4, [])
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
rs = np.random.RandomState(0)


def timeit(fn, n=10):
    ts = []
    for it in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = []
for B, L in ((256, 40), (65536, 40), (1 << 20, 40), (1 << 18, 200)):
    s = torch.randn(B, L, device="cuda")
    y = (torch.rand(B, L, device="cuda") < 0.2).float()
    table = torch.linspace(1, 9, 40, device="cuda")
    d = torch.empty(B, L, device="cuda")
    sums = torch.zeros(2, device="cuda")
    ms = timeit(lambda: eng.softmax_ce(s, y, 1, table, d, sums))
    byts = B * (12 * L + 8)
    out.append({"kernel": "K2 softmax_ce<IPW>", "B": B, "L": L, "ms": round(ms, 4), "algorithmic_GBs": round(byts / ms / 1e6, 1),
                "hbm_peak_GBs": hbm, "frac": round(byts / ms / 1e6 / hbm, 4)})
    del s, y, d
for B, L in ((256, 20), (1 << 20, 20), (1 << 18, 100)):
    # DLA (both losses + DenoisingNet): bytes/list = 12 L + 8 as for K2 (the propensity logits are L shared floats)
    s = torch.randn(B, L, device="cuda")
    y = (torch.rand(B, L, device="cuda") < 0.2).float()
    pw = torch.randn(L, device="cuda") * 0.1
    pb = torch.zeros(1, device="cuda")
    d = torch.empty(B, L, device="cuda")
    dp = torch.zeros(L + 1, device="cuda")
    sums = torch.zeros(4, device="cuda")
    ms = timeit(lambda: eng.dla_loss(s, y, pw, pb, d, dp, sums))
    byts = B * (12 * L + 8)
    out.append({"kernel": "K2 dla_loss", "B": B, "L": L, "ms": round(ms, 4), "algorithmic_GBs": round(byts / ms / 1e6, 1),
                "hbm_peak_GBs": hbm, "frac": round(byts / ms / 1e6 / hbm, 4)})
    del s, y, d
for B, L, kind in ((256, 200, "lambdarank"), (4096, 200, "lambdarank"), (4096, 40, "lambdarank"), (4096, 200, "pairdebias")):
    s = torch.randn(B, L, device="cuda")
    y = torch.randint(0, 5, (B, L), device="cuda").float() if kind == "lambdarank" else (torch.rand(B, L, device="cuda") < 0.2).float()
    tp = torch.ones(L, device="cuda"); tm = torch.ones(L, device="cuda")
    d = torch.empty(B, L, device="cuda")
    o = torch.zeros(2 * L + 2, device="cuda")
    fn = (lambda: eng.lambdarank(s, y, 1.0, tp, tm, d, o)) if kind == "lambdarank" else (lambda: eng.pairdebias(s, y, tp, tm, d, o))
    ms = timeit(fn)
    pairs = B * L * (L - 1) // 2
    out.append({"kernel": "K3 " + kind, "B": B, "L": L, "ms": round(ms, 4), "unordered_pairs": pairs,
                "Gpairs_per_s": round(pairs / ms / 1e6, 2)})
for r in out:
    print(json.dumps(r))
