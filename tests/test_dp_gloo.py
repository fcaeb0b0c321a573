"""World-size-2 `gloo` test of the data-parallel algebra: every rank computes UN-normalised
partials on its shard of the lists, ONE all-reduce over [grads | normalisers] follows, the division happens after
it - and the result equals the unsharded computation.  Runs on the CPU with the oracle as the per-rank compute."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    from oracle import ultra_oracle as uo
    rs = np.random.RandomState(5)
    F, hidden, L, B = 9, [8, 6], 5, 12
    params = {}
    for j, (k, n) in enumerate(uo.layer_sizes(F, hidden)):
        params["sequential.layer_norm%d.weight" % j] = 1.0 + 0.1 * rs.randn(k)
        params["sequential.layer_norm%d.bias" % j] = 0.1 * rs.randn(k)
        params["sequential.linear%d.weight" % j] = rs.randn(n, k) / np.sqrt(k)
        params["sequential.linear%d.bias" % j] = 0.1 * rs.randn(n)
    feats = rs.uniform(-1, 1, size=(L * B, F))
    docids = np.arange(L * B).reshape(B, L)
    clicks = (rs.rand(B, L) < 0.3).astype(np.float64)
    clicks[:, 0] = 1.0
    table = np.linspace(1, 5, 5)
    return uo, F, hidden, L, B, params, feats, docids, clicks, table


def _partials(uo, params, hidden, feats, docids, clicks, table):
    dt = np.float64
    n_layers = len(hidden) + 1
    s, cache = uo.ranking_scores(feats, np.ascontiguousarray(docids.T), params, n_layers, dt)
    pw = uo.ipw_weights(clicks, table, dt)
    loss, grad, num, den = uo.softmax_loss(s, clicks, pw, dt)
    g = uo.dnn_backward(uo.scores_grad_to_rows(grad * den), cache, params, n_layers, dt)    # un-normalised
    flat = np.concatenate([g[n].reshape(-1) for n in uo.param_names(n_layers)] + [np.array([num, den])])
    return flat


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uo, F, hidden, L, B, params, feats, docids, clicks, table = _problem()
    sl = slice(rank * (B // world), (rank + 1) * (B // world))
    buf = torch.from_numpy(_partials(uo, params, hidden, feats, docids[sl], clicks[sl], table))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)           # the ONE collective of a step
    out = buf.numpy()
    q.put((rank, out[:-2] / out[-1], out[-2] / out[-1]))
    dist.destroy_process_group()


def test_sharded_partials_allreduce_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    uo, F, hidden, L, B, params, feats, docids, clicks, table = _problem()
    full = _partials(uo, params, hidden, feats, docids, clicks, table)
    ref_g, ref_loss = full[:-2] / full[-1], full[-2] / full[-1]
    for rank, g, loss in results:
        assert np.allclose(g, ref_g, rtol=1e-10, atol=1e-12)
        assert abs(loss - ref_loss) < 1e-12
    assert np.array_equal(results[0][1], results[1][1])   # replicas stay bitwise equal
