"""Evaluation metrics used by `validation()` - an op-for-op restatement of the reference's torch code
(ultra/utils/metrics.py) so that, given identical scores, every metric is bit-identical to the reference's.

Restated: _safe_div :156-170, _per_example_weights_to_per_list_weights :173-188, _discounted_cumulative_gain
:191-221, _prepare_and_validate_params :224-265, mean_reciprocal_rank :268-298, expected_reciprocal_rank :300-336,
average_relevance_position :338-372, precision :375-405, mean_average_precision :408-455,
normalized_discounted_cumulative_gain :456-495.  ('dcg' and 'ordered_pair_accuracy' are not restated: the reference's
`discounted_cumulative_gain` calls its helper with the wrong arguments, metrics.py:520-521, and cannot run.)

The tensors stay on whatever device they are given on (validation() evaluates them on the host copy of the scores,
matching the reference's CPU path); the reference hard-codes a module-level `device`.
"""
import sys

import numpy as np
import torch

MAX_LABEL = None   # ERR normaliser; the reference keeps it in RankingMetricKey.MAX_LABEL (set by the data loader)


def get_max_label():
    ref = sys.modules.get("ultra.utils.metrics")
    if ref is not None and getattr(ref.RankingMetricKey, "MAX_LABEL", None) is not None:
        return float(ref.RankingMetricKey.MAX_LABEL)
    if MAX_LABEL is None:
        raise ValueError("metrics.MAX_LABEL is not set (the reference sets it while loading the data set)")
    return float(MAX_LABEL)


def _safe_div(numerator, denominator):
    return torch.where(torch.eq(denominator, 0), torch.zeros_like(numerator), torch.div(numerator, denominator))


def _per_example_weights_to_per_list_weights(weights, relevance):
    return _safe_div(torch.sum(weights * relevance, 1, keepdim=True), torch.sum(relevance, 1, keepdim=True))


def _discounted_cumulative_gain(prediction, labels, weights=None, topn=None):
    dev = labels.device
    list_size = labels.shape[1]
    _, indices = prediction.sort(descending=True, dim=-1)
    sorted_labels = torch.gather(labels, dim=1, index=indices)
    sorted_weights = torch.gather(weights, dim=1, index=indices)
    discounts = (torch.tensor(1) / torch.log2(torch.arange(list_size, dtype=torch.float) + 2.0)).to(device=dev)
    gains = sorted_weights * torch.pow(torch.tensor(2.0, device=dev), sorted_labels.to(torch.float32)) - 1.0
    discounted_gains = (gains * discounts)[:, :np.max(topn)]
    cum_dcg = torch.cumsum(discounted_gains, dim=1)
    topn_tensor = torch.tensor(topn, dtype=torch.long) - torch.tensor(1)
    return cum_dcg[:, topn_tensor.to(dev)]


def _prepare_and_validate_params(labels, predictions, weights=None, topn=None):
    weights = 1.0 if weights is None else weights
    example_weights = torch.ones_like(labels) * weights
    assert predictions.shape == example_weights.shape
    assert predictions.shape == labels.shape
    assert predictions.dim() == 2
    list_size = predictions.shape[1]
    if topn is None:
        topn = [list_size]
    topn = [min(n, list_size) for n in topn]
    is_label_valid = labels >= 0.
    labels = torch.where(is_label_valid, labels, torch.zeros_like(labels))
    predictions = torch.where(
        is_label_valid, predictions,
        -1e-6 * torch.ones_like(predictions) + torch.min(input=predictions, dim=1, keepdim=True).values)
    return labels, predictions, example_weights, topn


def mean_reciprocal_rank(labels, predictions, weights=None, topn=None):
    list_size = predictions.size()[-1]
    labels, predictions, weights, topn = _prepare_and_validate_params(labels, predictions, weights, topn)
    _, indices = predictions.sort(descending=True, dim=-1)
    sorted_labels = torch.gather(labels, dim=1, index=indices)
    relevance = torch.ge(sorted_labels, 1.0).type(torch.float32)
    reciprocal_rank = 1.0 / torch.arange(start=1, end=list_size + 1, device=labels.device, dtype=torch.float32)
    mrr = torch.max(relevance * reciprocal_rank, dim=1, keepdim=True).values
    return torch.mean(mrr * torch.ones_like(weights) * weights).repeat(len(topn))


def expected_reciprocal_rank(labels, predictions, weights=None, topn=None):
    dev = labels.device
    labels, predictions, weights, topn = _prepare_and_validate_params(labels, predictions, weights, topn)
    _, indices = predictions.sort(descending=True, dim=-1)
    sorted_labels = torch.gather(labels, dim=1, index=indices)
    sorted_weights = torch.gather(weights, dim=1, index=indices)
    list_size = sorted_labels.size()[-1]
    pow = torch.as_tensor(2.0, device=dev)
    relevance = (torch.pow(pow, sorted_labels) - 1) / torch.pow(pow, torch.as_tensor(get_max_label(), device=dev))
    non_rel = torch.cumprod(1.0 - relevance, dim=1) / (1.0 - relevance)
    reciprocal_rank = 1.0 / torch.arange(start=1, end=list_size + 1, device=dev, dtype=torch.float32)
    mask = [torch.ge(reciprocal_rank, 1.0 / n).type(torch.float32) for n in topn]
    reciprocal_rank_topn = [reciprocal_rank * top_n_mask for top_n_mask in mask]
    err = [torch.sum(relevance * non_rel * rr * sorted_weights, dim=1, keepdim=True) for rr in reciprocal_rank_topn]
    err = torch.stack(err, dim=0)
    return torch.mean(err, dim=1)


def average_relevance_position(labels, predictions, weights=None, topn=None):
    list_size = predictions.size()[1]
    labels, predictions, weights, topn = _prepare_and_validate_params(labels, predictions, weights, topn)
    _, indices = predictions.sort(descending=True, dim=-1)
    sorted_labels = torch.gather(labels, dim=1, index=indices)
    sorted_weights = torch.gather(weights, dim=1, index=indices)
    position = torch.arange(1, list_size + 1, dtype=torch.float, device=labels.device)
    weighted_labels = sorted_labels * sorted_weights
    per_list_weights = torch.sum(weighted_labels, dim=1, keepdim=True)
    per_list_arp = _safe_div(torch.sum(position * weighted_labels, dim=1, keepdim=True), per_list_weights)
    return torch.mean(per_list_arp).repeat(len(topn))


def precision(labels, predictions, weights=None, topn=None):
    labels, predictions, weights, topn = _prepare_and_validate_params(labels, predictions, weights, topn)
    _, indices = predictions.sort(descending=True, dim=-1)
    sorted_labels = torch.gather(labels, dim=1, index=indices)
    sorted_weights = torch.gather(weights, dim=1, index=indices)
    relevance = torch.ge(sorted_labels, 1.0).to(dtype=torch.float)
    per_list_precision = _safe_div(torch.sum(relevance * sorted_weights, 1, keepdim=True),
                                   torch.sum(torch.ones_like(relevance) * sorted_weights, 1, keepdim=True))
    per_list_weights = _per_example_weights_to_per_list_weights(weights, torch.ge(labels, 1.0).to(dtype=torch.float))
    return torch.mean(per_list_precision * per_list_weights)


def mean_average_precision(labels, predictions, weights=None, topn=None):
    labels, predictions, weights, topn = _prepare_and_validate_params(labels, predictions, weights, topn)
    _, indices = predictions.sort(descending=True, dim=-1)
    sorted_labels = torch.gather(labels, dim=1, index=indices)
    sorted_weights = torch.gather(weights, dim=1, index=indices)
    sorted_relevance = torch.ge(sorted_labels, 1.0).to(dtype=torch.float32)
    per_list_relevant_counts = torch.cumsum(sorted_relevance, dim=1)
    per_list_cutoffs = torch.cumsum(torch.ones_like(sorted_relevance), dim=1)
    per_list_precisions = torch.nan_to_num(torch.div(per_list_relevant_counts, per_list_cutoffs))
    total_precision = torch.sum(input=per_list_precisions * sorted_weights * sorted_relevance, dim=1, keepdim=True)
    total_relevance = torch.sum(input=sorted_weights * sorted_relevance, dim=1, keepdim=True)
    per_list_map = torch.nan_to_num(torch.div(total_precision, total_relevance))
    per_list_weights = _per_example_weights_to_per_list_weights(
        weights, torch.ge(labels, 1.0).to(dtype=torch.float32))
    return torch.mean(per_list_map * per_list_weights).repeat(len(topn))


def normalized_discounted_cumulative_gain(labels, predictions, weights=None, topn=None):
    had_weights = weights is not None
    labels, predictions, weights, topn = _prepare_and_validate_params(labels, predictions, weights, topn)
    dcg = _discounted_cumulative_gain(predictions, labels, weights, topn)
    ideal_dcg = _discounted_cumulative_gain(labels, labels, weights, topn)
    per_list_ndcg = _safe_div(dcg, ideal_dcg)
    if had_weights:
        per_list_weights = _per_example_weights_to_per_list_weights(
            weights=weights, relevance=torch.pow(torch.tensor(2.0), labels.to(torch.float)) - 1.0)
        return torch.mean(per_list_ndcg * per_list_weights)
    return torch.mean(per_list_ndcg, dim=0)


_METRICS = {
    "mrr": mean_reciprocal_rank,
    "err": expected_reciprocal_rank,
    "arp": average_relevance_position,
    "ndcg": normalized_discounted_cumulative_gain,
    "precision": precision,
    "map": mean_average_precision,
}


def make_ranking_metric_fn(metric_key, topn=None, name=None):
    """Same factory signature as ultra.utils.make_ranking_metric_fn (metrics.py:62-153)."""
    assert metric_key in _METRICS, 'metric_key %s not supported.' % metric_key
    fn = _METRICS[metric_key]

    def metric_fn(labels, predictions, weights):
        return fn(labels, predictions, weights=weights, topn=topn)
    return metric_fn
