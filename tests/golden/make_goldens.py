"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, CPU path).

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists):
    python oracle/install_ref.py && python tests/golden/make_goldens.py
The .npz files are committed; /root/reference is never read by the tests at run time.

Each golden holds, for one (algorithm, shape) case driven through the reference's own
`BaseAlgorithm.train()/validation()` (ipw_rank.py:102-211, dla.py:179-285, pairwise_debias.py:106-203,
lambda_rank.py:96-245, navie_algorithm.py:76-149):
  - the hand-built input_feeds (same dict format ClickSimulationFeed.get_batch emits,
    click_simulation_feed.py:141-156), the initial state_dict,
  - validation() scores + metric values for the initial parameters,
  - per training step: loss, pre-clip grads (captured by wrapping clip_grad_norm_, arithmetic untouched),
    post-step parameters, t_plus/t_minus and DLA propensity-net parameters.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

ALGOS = {
    "na": "ultra.learning_algorithm.NavieAlgorithm",
    "ipw": "ultra.learning_algorithm.IPWrank",
    "dla": "ultra.learning_algorithm.DLA",
    "pairdebias": "ultra.learning_algorithm.PairDebias",
    "lambdarank": "ultra.learning_algorithm.LambdaRank",
    "prsrank": "ultra.learning_algorithm.PRSrank",
    "regem": "ultra.learning_algorithm.RegressionEM",
}

# name, algo, F, L_train (selection_bias_cutoff), L_max (max_candidate_num), B, hidden, label kind, n_steps
CASES = [
    ("na_small", "na", 10, 6, 6, 8, [16, 8], "graded", 3),
    ("ipw_small", "ipw", 10, 6, 8, 8, [16, 8], "click", 3),
    ("ipw_l45", "ipw", 12, 45, 45, 5, [24], "click", 2),          # list longer than the 40-entry IPW table
    ("dla_small", "dla", 10, 6, 8, 8, [16, 8], "click", 3),
    ("pairdebias_small", "pairdebias", 10, 6, 6, 8, [16, 8], "click", 3),
    ("lambdarank_small", "lambdarank", 10, 6, 6, 8, [16, 8], "graded", 3),
    ("lambdarank_click", "lambdarank", 10, 7, 9, 6, [16, 8], "click", 2),
    ("ipw_c2like", "ipw", 136, 40, 40, 16, [256, 128, 64], "click", 1),
    ("dla_wide", "dla", 72, 20, 20, 12, [128, 64, 32], "click", 1),
    # hidden == [] selects the reference's Linear ranker (ultra/ranking_model/Linear.py: LayerNorm -> Linear(F, 1))
    ("ipw_linear", "ipw", 10, 6, 8, 8, [], "click", 3),
    ("lambdarank_linear", "lambdarank", 14, 7, 7, 6, [], "graded", 2),
    ("prsrank_small", "prsrank", 10, 6, 8, 8, [16, 8], "graded", 3),
    ("prsrank_click", "prsrank", 10, 7, 7, 6, [16, 8], "click", 2),
    # BASELINE.json shapes (configs 3, 5, 4 and the config-2 net under the pairwise loss): every hidden layer runs on the
    # tensor cores, so reference-generated numbers go THROUGH the tcgen05 kernels for DLA / LambdaRank / PairDebias too.
    # Stored compactly (trailing True): features as float32 (they are float32-exact), no post-step parameters.
    # RegressionEM samples pseudo-labels: the uniform draws of get_bernoulli_sample are recorded with the golden
    ("regem_small", "regem", 10, 6, 8, 8, [16, 8], "click", 3),
    ("regem_c2net", "regem", 136, 40, 40, 16, [256, 128, 64], "click", 1, True),
    ("dla_c3like", "dla", 700, 20, 20, 16, [512, 256, 128], "click", 1, True),
    ("dla_c5like", "dla", 220, 100, 100, 8, [512, 256, 128], "click", 1, True),
    ("lambdarank_c4like", "lambdarank", 136, 200, 200, 8, [512, 256, 128], "graded", 1, True),
    ("pairdebias_c2net", "pairdebias", 136, 40, 40, 16, [256, 128, 64], "click", 1, True),
    # a non-default algorithm hparam (trailing string): L2 regularisation of the ranker parameters
    ("ipw_l2", "ipw", 10, 6, 8, 8, [16, 8], "click", 3, False, "l2_loss=0.01"),
    ("dla_l2", "dla", 10, 6, 8, 8, [16, 8], "click", 2, False, "l2_loss=0.01"),
    ("pairdebias_l2", "pairdebias", 10, 6, 6, 8, [16, 8], "click", 2, False, "l2_loss=0.01"),
    # the other hidden-layer activations of the reference ranker (second trailing string: ranking_model hparams)
    ("ipw_relu", "ipw", 10, 6, 8, 8, [16, 8], "click", 2, False, "", "activation_func=relu"),
    # (activation_func=selu cannot be generated: the reference's selu is a plain function and nn.Sequential.add_module
    # raises TypeError for it, DNN.py:52-54)
    ("dla_tanh", "dla", 10, 6, 8, 8, [16, 8], "click", 2, False, "", "activation_func=tanh"),
    ("lambdarank_sigmoid", "lambdarank", 10, 6, 6, 8, [16, 8], "graded", 2, False, "", "activation_func=sigmoid"),
]


def make_feed(rs, model, B, L_feed, F, kind, pad_frac=0.15):
    """Builds a feed dict exactly shaped like ClickSimulationFeed.get_batch's (click_simulation_feed.py:141-156)."""
    docids = np.zeros((B, L_feed), dtype=np.int64)
    feats = []
    for b in range(B):
        n_real = L_feed if rs.rand() > 0.5 else max(1, int(round(L_feed * (1.0 - pad_frac * rs.rand() * 3))))
        n_real = min(max(n_real, 1), L_feed)
        for l in range(L_feed):
            if l < n_real:
                docids[b, l] = len(feats)
                feats.append(rs.uniform(-1, 1, size=F).astype(np.float32).astype(np.float64))
            else:
                docids[b, l] = -1
    n_docs = len(feats)
    docids[docids < 0] = n_docs
    if kind == "click":
        exam = np.array([0.68, 0.61, 0.48, 0.34, 0.28, 0.2, 0.11, 0.1, 0.08, 0.06])
        p = 0.6 * exam[np.minimum(np.arange(L_feed), 9)] + 0.05
        labels = (rs.rand(B, L_feed) < p[None, :]).astype(np.float32)
        for b in range(B):  # ClickSimulationFeed drops lists without clicks (check_validation=True)
            if labels[b].sum() == 0:
                labels[b, rs.randint(0, L_feed)] = 1.0
        labels[0, 1:] = 0.0  # single click on the first doc
        labels[0, 0] = 1.0
    else:
        labels = rs.randint(0, 5, size=(B, L_feed)).astype(np.float32)
        labels[docids == n_docs] = 0.0
    feed = {model.letor_features_name: np.array(feats)}
    for l in range(L_feed):
        feed[model.docid_inputs_name[l]] = docids[:, l].astype(np.float32)
        feed[model.labels_name[l]] = labels[:, l].astype(np.float32)
    return feed, docids, labels


def state_to_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def run_case(ultra, name, algo, F, L_train, L_max, B, hidden, kind, n_steps, compact=False, hparams="",
             model_hparams=""):
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    ds = ultra.utils.data_utils.Raw_data()
    ds.feature_size = F
    ds.rank_list_size = L_max
    ultra.utils.metrics.RankingMetricKey.MAX_LABEL = 4.0
    exp_settings = {
        "learning_algorithm": ALGOS[algo],
        "learning_algorithm_hparams": hparams,
        "ranking_model": "ultra.ranking_model.DNN" if hidden else "ultra.ranking_model.Linear",
        "ranking_model_hparams": ",".join(filter(None, [("hidden_layer_sizes=%s" % str(hidden)) if hidden else "",
                                                        model_hparams])),
        "selection_bias_cutoff": L_train,
        "max_candidate_num": L_max,
        "metrics": ["ndcg", "err", "mrr"],
        "metrics_topn": [1, 3, 5, 10],
    }
    cls = ultra.utils.find_class(ALGOS[algo])
    model = cls(ds, exp_settings)
    drawn = []
    if algo == "regem":
        # record the uniform draws of the pseudo-label sampler (regression_EM.py:20-34; same expression, CPU branch)
        mod = sys.modules["ultra.learning_algorithm.regression_EM"]

        def recording_sampler(probs):
            u = torch.rand(probs.shape)
            drawn.append(u)
            return torch.ceil(probs - u)
        mod.get_bernoulli_sample = recording_sampler
    # give LayerNorm affine params non-trivial values so that their gradients are exercised
    with torch.no_grad():
        for n, p in model.model.named_parameters():
            if "layer_norm" in n:
                p.add_(0.2 * torch.randn_like(p))
    out = {}
    out["meta_F"] = np.int64(F)
    out["meta_L_train"] = np.int64(L_train)
    out["meta_L_max"] = np.int64(L_max)
    out["meta_B"] = np.int64(B)
    out["meta_hidden"] = np.asarray(hidden, dtype=np.int64)
    out["meta_n_steps"] = np.int64(n_steps)
    out["meta_algo"] = np.asarray(algo)
    out["meta_hparams"] = np.asarray(hparams)
    out["meta_model_hparams"] = np.asarray(model_hparams)
    out.update(state_to_np(model.model.state_dict(), "init/"))
    if algo == "dla":
        out.update(state_to_np(model.propensity_model.state_dict(), "init_prop/"))
    if algo in ("ipw", "prsrank"):
        out["ipw_table"] = np.asarray(model.propensity_estimator.IPW_list, dtype=np.float64)

    # validation on the initial parameters (forward only, L = max_candidate_num)
    vfeed, vdoc, vlab = make_feed(rs, model, B, L_max, F, "graded")
    feat_dtype = np.float32 if compact else np.float64          # tests/helpers.py: load_golden casts back to float64
    out["valid/features"] = vfeed[model.letor_features_name].astype(feat_dtype)
    out["valid/docids"] = vdoc
    out["valid/labels"] = vlab
    _, scores, summary = model.validation(dict(vfeed))
    out["valid/scores"] = scores.detach().cpu().numpy().copy()
    for k, v in summary.items():
        out["valid/metric/" + k] = np.float64(v)

    grads_seen = []
    orig_clip = torch.nn.utils.clip_grad_norm_

    def recording_clip(parameters, max_norm, *a, **kw):
        parameters = list(parameters)
        grads_seen.append([p.grad.detach().clone() if p.grad is not None else None for p in parameters])
        return orig_clip(parameters, max_norm, *a, **kw)

    torch.nn.utils.clip_grad_norm_ = recording_clip
    try:
        for step in range(n_steps):
            feed, doc, lab = make_feed(rs, model, B, L_max, F, kind)
            pre = "step%d/" % step
            out[pre + "features"] = feed[model.letor_features_name].astype(feat_dtype)
            out[pre + "docids"] = doc
            out[pre + "labels"] = lab
            grads_seen.clear()
            drawn.clear()
            loss, _, _ = model.train(dict(feed))
            out[pre + "loss"] = np.float64(loss)
            if algo == "regem":
                out[pre + "uniform"] = drawn[0].numpy().copy()
                out[pre + "propensity"] = model.propensity.detach().numpy().copy()
            names = [n for n, _ in model.model.named_parameters()]
            if algo == "dla":
                # separate_gradient_update clips the propensity net first, then the ranker (dla.py:161-163)
                pnames = [n for n, _ in model.propensity_model.named_parameters()]
                for n, g in zip(pnames, grads_seen[0]):
                    out[pre + "grad_prop/" + n] = g.numpy().copy()
                for n, g in zip(names, grads_seen[1]):
                    out[pre + "grad/" + n] = g.numpy().copy()
                if not compact:
                    out.update(state_to_np(model.propensity_model.state_dict(), pre + "param_prop/"))
            else:
                for n, g in zip(names, grads_seen[0]):
                    out[pre + "grad/" + n] = g.numpy().copy()
            if not compact:
                out.update(state_to_np(model.model.state_dict(), pre + "param/"))
            if algo in ("pairdebias", "lambdarank"):
                out[pre + "t_plus"] = model.t_plus.detach().numpy().copy()
                out[pre + "t_minus"] = model.t_minus.detach().numpy().copy()
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d KB) losses=%s" % (path, os.path.getsize(path) // 1024,
                                          [float(out["step%d/loss" % s]) for s in range(n_steps)]))


def main():
    ultra = ref_shim.load()
    only = sys.argv[1:]
    with ref_shim.ref_cwd():
        for case in CASES:
            if only and case[0] not in only:
                continue
            run_case(ultra, *case)


if __name__ == "__main__":
    main()
