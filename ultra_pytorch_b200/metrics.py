"""Evaluation metrics of `validation()`.

NDCG / ERR / MRR - what the reference's experiment settings ask for - are computed per ranked list ON THE DEVICE
(csrc/metrics.cu: PAD masking, stable ranking, the sequential DCG / ERR chains with the reference's torch-CPU
accumulation rules), so a validation batch returns B x (2 n + 1) floats instead of its B x L scores.  This module holds
the host side: the discount table (computed with the reference's own torch expression, so the kernel multiplies by
bit-identical values) and the batch means, reduced with the torch calls the reference ends with
(ultra/utils/metrics.py:297, :336, :494) on tensors of the same shape and memory layout.

Any other metric key (arp, precision, map, ...) is delegated to the reference's own module when the plugin runs inside
the reference (`ultra.utils.metrics` is importable under main.py); standing alone they are not available.
"""
import sys

import torch

from ._capi import check, int_array, lib

MAX_LABEL = None   # ERR normaliser; the reference keeps it in RankingMetricKey.MAX_LABEL (set by the data loader)
DEVICE_METRICS = ("ndcg", "err", "mrr")


def get_max_label():
    ref = sys.modules.get("ultra.utils.metrics")
    if ref is not None and getattr(ref.RankingMetricKey, "MAX_LABEL", None) is not None:
        return float(ref.RankingMetricKey.MAX_LABEL)
    if MAX_LABEL is None:
        raise ValueError("metrics.MAX_LABEL is not set (the reference sets it while loading the data set)")
    return float(MAX_LABEL)


_discounts = {}


def discount_table(L, device):
    """1 / log2(rank + 2) exactly as ultra/utils/metrics.py:212 forms it (torch CPU float32), uploaded once per L."""
    t = _discounts.get((L, str(device)))
    if t is None:
        t = (torch.tensor(1) / torch.log2(torch.arange(L, dtype=torch.float) + 2.0)).to(device)
        _discounts[(L, str(device))] = t
    return t


def per_list_metrics(scores, labels, docid, n_docs, topn, stream=None):
    """scores / labels [B, L] f32 cuda, docid [L, B] i32 cuda (PAD id == n_docs) or None ->
    (per_list [B, 2 n + 1] cuda: ndcg@topn | err@topn | mrr, flag [1] int32 cuda: labels the kernel cannot take)."""
    B, L = scores.shape
    n = len(topn)
    out = torch.empty(B, 2 * n + 1, dtype=torch.float32, device=scores.device)
    flag = torch.zeros(1, dtype=torch.int32, device=scores.device)
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    check(lib.ub200_rank_metrics(scores.data_ptr(), labels.data_ptr(), 0 if docid is None else docid.data_ptr(),
                                 int(n_docs), B, L, discount_table(L, scores.device).data_ptr(), int_array(topn), n,
                                 get_max_label(), out.data_ptr(), flag.data_ptr(), st), "ub200_rank_metrics")
    return out, flag


def batch_means(per_list, L, topn):
    """Host side: per_list [B, 2 n + 1] (CPU) -> {metric: [n values]} with the reference's closing reductions."""
    B, n = per_list.shape[0], len(topn)
    ndcg = torch.mean(per_list[:, :n].contiguous(), dim=0)                                 # metrics.py:494
    err = torch.mean(torch.stack([per_list[:, n + q:n + q + 1].contiguous() for q in range(n)], dim=0), dim=1)  # :336
    # metrics.py:297: mean over a [B, L] tensor of the per-list value times the example weights, which inherit the
    # TRANSPOSED memory layout of the labels (base_algorithm.py:181-182) - torch reduces in memory order
    w = torch.ones(L, B, dtype=torch.float32).t()
    mrr = torch.mean(per_list[:, 2 * n:2 * n + 1].contiguous() * torch.ones_like(w) * w).repeat(n)
    return {"ndcg": ndcg, "err": err.view(-1), "mrr": mrr}


def reference_metric_fn(metric_key, topn):
    """The reference's own implementation, for keys the device path does not cover (only inside the reference)."""
    ref = sys.modules.get("ultra.utils.metrics") or sys.modules.get("ultra.utils")
    if ref is None or not hasattr(ref, "make_ranking_metric_fn"):
        raise NotImplementedError("metric '%s' is computed by the reference's ultra.utils.metrics, which is not "
                                  "importable here; the device path covers %s" % (metric_key, DEVICE_METRICS))
    return ref.make_ranking_metric_fn(metric_key, topn)
