"""Seeded synthetic MSLR-WEB30K-shaped workloads: features ~ U(-1,1) (the reference pipelines
normalise features to [-1,1]), labels ~ UniformInt{0..4}, clicks from a position-biased click model with the
reference's default parameters (ultra/utils/click_models.py:68-110, example/ClickModel/pbm_0.1_1.0_4_1.0.json).
The emitted dict has exactly the format ClickSimulationFeed.get_batch produces (click_simulation_feed.py:141-156):
"letor_features" f64 [n_docs, F], "docid_input{l}" f32 [B], "label{l}" f32 [B] (PAD id == n_docs)."""
import os

import numpy as np

EXAM_PROB = np.array([0.68, 0.61, 0.48, 0.34, 0.28, 0.20, 0.11, 0.10, 0.08, 0.06])
CLICK_PROB = np.array([0.1, 0.16, 0.28, 0.52, 1.0])
IPW_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ipw_pbm_analytic.json")
# the PBM click model of the reference's examples (example/ClickModel/pbm_0.1_1.0_4_1.0.json: same numbers)
PBM_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pbm_0.1_1.0_4_1.0.json")

WORKLOADS = {
    # BASELINE.json configs[1..4] (B = the reference's default --batch_size, main.py:42)
    "c2_ipw_mslr10k": dict(algo="IPWrank", F=136, L=40, B=256, hidden=[256, 128, 64], labels="click"),
    "c3_dla_yahoo": dict(algo="DLA", F=700, L=20, B=256, hidden=[512, 256, 128], labels="click"),
    "c4_lambdarank_mslr30k": dict(algo="LambdaRank", F=136, L=200, B=256, hidden=[512, 256, 128], labels="click"),
    "c4_pairdebias_mslr30k": dict(algo="PairDebias", F=136, L=200, B=256, hidden=[512, 256, 128], labels="click"),
    "c5_dla_istella": dict(algo="DLA", F=220, L=100, B=256, hidden=[512, 256, 128], labels="click"),
    "c1_na_toy": dict(algo="NavieAlgorithm", F=136, L=9, B=256, hidden=[512, 256, 128], labels="graded"),
}


def pbm_clicks(rs, labels):
    B, L = labels.shape
    exam = EXAM_PROB[np.minimum(np.arange(L), len(EXAM_PROB) - 1)]
    p = exam[None, :] * CLICK_PROB[np.minimum(labels.astype(np.int64), len(CLICK_PROB) - 1)]
    return (rs.rand(B, L) < p).astype(np.float32)


def make_feed(seed, F, L, B, labels="click", pad_tail=0):
    """One input_feed.  Lists without any click are re-sampled (ClickSimulationFeed(check_validation=True),
    click_simulation_feed.py:89-93)."""
    rs = np.random.RandomState(seed)
    n_docs = B * L
    feats = rs.uniform(-1.0, 1.0, size=(n_docs, F)).astype(np.float32).astype(np.float64)
    rel = rs.randint(0, 5, size=(B, L))
    docids = np.arange(n_docs, dtype=np.int64).reshape(B, L)
    if pad_tail:
        docids[:, L - pad_tail:] = n_docs
        rel[:, L - pad_tail:] = 0
    if labels == "click":
        y = pbm_clicks(rs, rel)
        empty = y.sum(axis=1) == 0
        while empty.any():
            y[empty] = pbm_clicks(rs, rel[empty])
            empty = y.sum(axis=1) == 0
    else:
        y = rel.astype(np.float32)
    feed = {"letor_features": feats}
    for l in range(L):
        feed["docid_input%d" % l] = docids[:, l].astype(np.float32)
        feed["label%d" % l] = y[:, l].astype(np.float32)
    return feed


def exp_settings(workload):
    w = WORKLOADS[workload] if isinstance(workload, str) else workload
    hp = "propensity_estimator_json=%s" % IPW_JSON if w["algo"] == "IPWrank" else ""
    return {
        "learning_algorithm": "ultra_pytorch_b200.learning_algorithm.%s" % w["algo"],
        "learning_algorithm_hparams": hp,
        "ranking_model": "ultra_pytorch_b200.ranking_model.DNN",
        "ranking_model_hparams": "hidden_layer_sizes=%s" % str(w["hidden"]),
        "selection_bias_cutoff": w["L"],
        "max_candidate_num": w["L"],
        "metrics": ["ndcg", "err"],
        "metrics_topn": [1, 3, 5, 10],
    }


def train_flops_per_query(F, L, hidden):
    """L * (3 * 2 * sum_j K_j N_j - 2 * F * N_0), each fp32 MAC counted as 2 FLOPs once."""
    ks = [F] + list(hidden)
    ns = list(hidden) + [1]
    macs = sum(k * n for k, n in zip(ks, ns))
    return L * (6 * macs - 2 * F * ns[0])


class SyntheticDataset(object):
    """In-memory stand-in for ultra.utils.data_utils.Raw_data after pad() (data_utils.py:476-498): `features` (one row
    per document + the trailing zero row pad() appends), `initial_list[q]` (row ids, -1 = PAD), `labels[q]`."""

    def __init__(self, n_queries, L, F, seed=0, max_label=4):
        rs = np.random.RandomState(seed)
        self.feature_size, self.rank_list_size = F, L
        feats = rs.uniform(-1.0, 1.0, size=(n_queries * L + 1, F)).astype(np.float32).astype(np.float64)
        feats[-1] = 0.0
        self.features = feats                                   # an ndarray works wherever the list of lists does
        self.initial_list = np.arange(n_queries * L, dtype=np.int64).reshape(n_queries, L).tolist()
        self.labels = rs.randint(0, max_label + 1, size=(n_queries, L)).astype(float).tolist()


def synthetic_dataset(n_queries, L, F, seed=0):
    return SyntheticDataset(n_queries, L, F, seed)
