"""GPU debugging aid (not a pytest): A/B of the K1 engines (ub200_set_tc_mode masks) against the CUDA-core fp32 path,
layer by layer, with CUDA-event timings.  Usage: python tools/debug_f16.py F L B h1,h2,h3 [modes]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultra_pytorch_b200 import _capi  # noqa: E402
from ultra_pytorch_b200.engine import RankerEngine  # noqa: E402


def align(x, a=256):
    return (x + a - 1) // a * a


def main():
    F, L, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    hidden = [int(h) for h in sys.argv[4].split(",")]
    modes = [0] + [int(m) for m in (sys.argv[5] if len(sys.argv) > 5 else "15,31,63").split(",")]
    M = L * B
    rs = np.random.RandomState(0)
    eng = RankerEngine(F, hidden)
    flat = []
    k = F
    for n in hidden + [1]:
        flat += [1.0 + 0.2 * rs.randn(k), 0.2 * rs.randn(k), rs.uniform(-1, 1, size=n * k) / np.sqrt(k),
                 rs.uniform(-1, 1, size=n) / np.sqrt(k)]
        k = n
    eng.params.copy_(torch.as_tensor(np.concatenate(flat), dtype=torch.float32, device="cuda"))
    feats = torch.as_tensor(rs.uniform(-1, 1, size=(M + 1, F)), dtype=torch.float32, device="cuda")
    feats[M] = 0
    docid = torch.as_tensor(rs.randint(0, M + 1, size=M), dtype=torch.int32, device="cuda")
    dsc = torch.as_tensor(rs.randn(B, L), dtype=torch.float32, device="cuda")
    res = {}
    for mode in modes:
        _capi.lib.ub200_set_tc_mode(mode)
        scores = eng.forward(feats, docid, L, B, training=True).clone()
        torch.cuda.synchronize()
        ws = eng._mlp_ws(L, B, True)
        off = 0
        nl = len(hidden) + 1
        for j in range(nl):
            off += align(8 * M)
        ys = []
        for j in range(nl - 1):
            nb = 4 * M * hidden[j]
            ys.append(ws[off:off + nb].view(torch.float32).view(M, hidden[j]).clone())
            off += align(nb)
        grads = eng.backward(feats, docid, L, B, dsc).clone()
        torch.cuda.synchronize()
        # timing
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        reps = 20
        for _ in range(3):
            eng.forward(feats, docid, L, B, training=True)
            eng.backward(feats, docid, L, B, dsc)
        for _ in range(reps):
            ev[0].record()
            eng.forward(feats, docid, L, B, training=True)
            ev[1].record()
            eng.backward(feats, docid, L, B, dsc)
            ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1])
            tb += ev[1].elapsed_time(ev[2])
        grads = eng.backward(feats, docid, L, B, dsc).clone()      # a later step: operand images (mask 128) are in use
        torch.cuda.synchronize()
        res[mode] = (scores, ys, grads)
        print("mode %d done: fwd %.1f us  bwd %.1f us (eager launches, incl. prep)" % (mode, 1e3 * tf / reps, 1e3 * tb / reps),
              flush=True)
    s0, y0, g0 = res[0]

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().mean().clamp_min(1e-30))

    for mode in modes[1:]:
        s1, y1, g1 = res[mode]
        print("==== tc mask %d vs 0 ====" % mode)
        for j, (a, b) in enumerate(zip(y1, y0)):
            bad = (a - b).abs() > 1e-4 * b.abs().mean()
            print("layer %d Y: max|d|/mean|ref| = %.3e  nan=%d  bad=%d/%d" % (j, rel(a, b), int(torch.isnan(a).sum()),
                                                                        int(bad.sum()), a.numel()))
        print("scores: %.3e nan=%d" % (rel(s1, s0), int(torch.isnan(s1).sum())))
        for (name, off_, shape) in eng.layer_slices():
            n = int(np.prod(shape))
            a, b = g1[off_:off_ + n], g0[off_:off_ + n]
            scale = max(float(b.abs().mean()), 0.1 * float(g0.abs().max()))
            print("grad %-22s max|d|/scale = %.3e nan=%d" % (name, float((a - b).abs().max()) / scale,
                                                          int(torch.isnan(a).sum())))


if __name__ == "__main__":
    main()
