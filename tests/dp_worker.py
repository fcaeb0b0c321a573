"""Worker for tests/test_gpu_dp.py (launched by torch.distributed.run, one rank per GPU): data-parallel IPW / LambdaRank
steps; checks (1) replicas stay bitwise identical, (2) the DP step equals a single-GPU step on the merged batch."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultra_pytorch_b200 import synth  # noqa: E402
import ultra_pytorch_b200.learning_algorithm as la  # noqa: E402


def merged_feed(feeds, L):
    feats = np.concatenate([f["letor_features"] for f in feeds], axis=0)
    out = {"letor_features": feats}
    n_total = feats.shape[0]
    for l in range(L):
        d, y, base = [], [], 0
        for f in feeds:
            n = f["letor_features"].shape[0]
            di = f["docid_input%d" % l].astype(np.int64)
            d.append(np.where(di == n, n_total, di + base))
            y.append(f["label%d" % l])
            base += n
        out["docid_input%d" % l] = np.concatenate(d).astype(np.float32)
        out["label%d" % l] = np.concatenate(y).astype(np.float32)
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    la.B200Algorithm.VERBOSE = False
    F, L, B, hidden = 136, 40, 64, [256, 128, 64]
    ok = True
    # DLA's update is lr * sign(g) (fresh Adagrad every step, dla.py:153-154): a gradient entry near zero may change
    # sign between the sharded and the merged evaluation, so for DLA only the replica equality is asserted
    for algo, wl, vs_single in (("IPWrank", "c2_ipw_mslr10k", True), ("LambdaRank", "c4_lambdarank_mslr30k", True),
                                ("DLA", "c3_dla_yahoo", False)):
        settings = synth.exp_settings(wl)
        settings.update({"ranking_model_hparams": "hidden_layer_sizes=%s" % hidden, "selection_bias_cutoff": L,
                         "max_candidate_num": L})
        ds = types.SimpleNamespace(feature_size=F)
        torch.manual_seed(0)
        model = getattr(la, algo)(ds, settings)
        init = {k: v.clone() for k, v in model.model.state_dict().items()}
        dp_losses = []
        for step in range(4):                      # step 2+ runs through the CUDA-graph DP path
            feed = synth.make_feed(100 * step + rank, F, L, B, "click")
            dp_losses.append(model.train(feed)[0])
        flat = model.engine.params.clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        # single-GPU reference on the merged batches (rank 0 only; same seeds)
        err = 0.0
        if rank == 0 and vs_single:
            torch.manual_seed(0)
            # world_size() == 1 makes the reference model a plain single-GPU one (no symmetric-memory rendezvous, which
            # is collective, and no exchange in its steps)
            la.B200Algorithm.world_size = staticmethod(lambda: 1)
            ref = getattr(la, algo)(ds, settings)
            ref.model.load_state_dict(init)
            ref_losses = []
            for step in range(4):
                feeds = [synth.make_feed(100 * step + r, F, L, B, "click") for r in range(world)]
                ref_losses.append(ref.train(merged_feed(feeds, L))[0])
            # data-parallel train() returns the exact global loss one step late (base_algorithm.py: LAG_LOSS_DP); the
            # first call returns its own
            lag = 1 if la.B200Algorithm.LAG_LOSS_DP else 0
            want = [ref_losses[max(0, k - lag)] for k in range(4)]
            loss_err = max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(dp_losses, want))
            print("DP %s: losses %s vs single-GPU (lag %d) %s -> rel err %.2e" % (algo, dp_losses, lag, want, loss_err),
                  flush=True)
            if loss_err > 1e-4:
                err = max(err, 1.0)
            a, b = flat, ref.engine.params
            # exclude the entries whose gradient is mathematically zero under shift-invariant losses (last LayerNorm
            # bias, last linear bias): Adagrad turns their rounding noise into +-lr moves (tests/test_oracle_vs_golden.py)
            keep = torch.ones_like(a, dtype=torch.bool)
            nl = len(hidden)
            for name, off, shape in model.engine.layer_slices():
                if name in ("layer_norm%d.bias" % nl, "linear%d.bias" % nl):
                    keep[off:off + int(np.prod(shape))] = False
            err = max(err, float((a - b)[keep].abs().max() / b[keep].abs().mean()))
        la.B200Algorithm.world_size = staticmethod(
            lambda: dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1)
        t = torch.tensor([1.0 if same else 0.0, err], device="cuda")
        dist.broadcast(t, 0)
        if rank == 0:
            print("DP %s: replicas bitwise equal=%s, max|dp - single|/mean|param| = %.3e" % (algo, same, err),
                  flush=True)
        ok = ok and same and float(t[1]) < 2e-4
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
