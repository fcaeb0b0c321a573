"""GPU timing aid: back-to-back K1 forward (training / inference) and backward launches, CUDA events.
Usage: python tools/time_k1.py F L B h1,h2,h3 [modes]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultra_pytorch_b200 import _capi
from ultra_pytorch_b200.engine import RankerEngine

F, L, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hidden = [int(h) for h in sys.argv[4].split(",")]
modes = [int(m) for m in (sys.argv[5] if len(sys.argv) > 5 else "15,63").split(",")]
M = L * B
rs = np.random.RandomState(0)
eng = RankerEngine(F, hidden)
eng.params.copy_(torch.as_tensor(rs.uniform(-0.1, 0.1, size=eng.P) + 0.5, dtype=torch.float32, device="cuda"))
feats = torch.as_tensor(rs.uniform(-1, 1, size=(M + 1, F)), dtype=torch.float32, device="cuda")
docid = torch.as_tensor(rs.randint(0, M + 1, size=M), dtype=torch.int32, device="cuda")
dsc = torch.as_tensor(rs.randn(B, L), dtype=torch.float32, device="cuda")
flops_f = 2.0 * M * sum(k * n for k, n in zip([F] + hidden, hidden + [1]))
flops_t = 3 * flops_f - 2.0 * M * F * hidden[0]


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


for mode in modes:
    _capi.lib.ub200_set_tc_mode(mode)
    t_inf = timed(lambda: eng.forward(feats, docid, L, B, training=False))
    t_trn = timed(lambda: eng.forward(feats, docid, L, B, training=True))
    t_bwd = timed(lambda: eng.backward(feats, docid, L, B, dsc))
    print("mode %2d  M=%d  fwd(inference) %.1f us  fwd(training) %.1f us  bwd %.1f us   K1 train %.1f TFLOP/s" % (
        mode, M, t_inf, t_trn, t_bwd, flops_t / (t_trn + t_bwd) * 1e-6), flush=True)
