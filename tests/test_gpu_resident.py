"""GPU test of the device-resident data set path (input_layer/resident.py): a feed that emits the data set's whole
feature matrix + global doc ids trains bit-identically to the per-batch-copy feed, while a step moves only ids and
labels to the device."""
import json
import os
import random
import types

import numpy as np
import pytest
import torch

from tests.test_click_feed import PBM, FakeData

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", ["NavieAlgorithm", "LambdaRank"])
def test_resident_dataset_trains_bit_identically(algo, tmp_path):
    import ultra_pytorch_b200.learning_algorithm as la
    from ultra_pytorch_b200.input_layer import ClickSimulationFeed
    la.B200Algorithm.VERBOSE = False
    L, F, B = 12, 136, 48
    ds = FakeData(200, L, F)
    p = os.path.join(str(tmp_path), "pbm.json")
    with open(p, "w") as f:
        json.dump(PBM, f)
    settings = {"learning_algorithm_hparams": "", "ranking_model": "ultra_pytorch_b200.ranking_model.DNN",
                "ranking_model_hparams": "hidden_layer_sizes=[64, 32]", "selection_bias_cutoff": L,
                "max_candidate_num": L, "metrics": ["ndcg"], "metrics_topn": [1, 3]}
    runs = []
    for hp in ("", "resident_features=True"):
        torch.manual_seed(0)
        random.seed(0)
        model = getattr(la, algo)(types.SimpleNamespace(feature_size=F), settings)
        feed = ClickSimulationFeed(model, B, "click_model_json=%s,oracle_mode=True,%s" % (p, hp))
        losses = []
        for step in range(5):                        # steps 3+ replay the captured CUDA graph
            f, _ = feed.get_next_batch(step * B % 150, ds)
            loss, _, _ = model.train(f)
            losses.append(loss)
        _, scores, summary = model.validation(feed.get_next_batch(7, ds)[0])
        runs.append((losses, model.engine.params.clone(), scores.clone(), dict(summary), model.last_h2d_bytes))
    (la_, pa, sa, ma, ha), (lb, pb, sb, mb, hb) = runs
    assert la_ == lb
    assert torch.equal(pa, pb) and torch.equal(sa, sb) and ma == mb
    assert hb == 8 * L * B and ha > 20 * hb          # ids + labels only vs ids + labels + feature rows


def _feed_arrays(tmp_path, nq, L, F, B, hp=""):
    import ultra_pytorch_b200.learning_algorithm as la
    from ultra_pytorch_b200.input_layer import ClickSimulationFeed
    la.B200Algorithm.VERBOSE = False
    ds = FakeData(nq, L, F)
    p = os.path.join(str(tmp_path), "pbm.json")
    with open(p, "w") as f:
        json.dump(PBM, f)
    settings = {"learning_algorithm_hparams": "", "ranking_model": "ultra_pytorch_b200.ranking_model.DNN",
                "ranking_model_hparams": "hidden_layer_sizes=[32, 16]", "selection_bias_cutoff": L,
                "max_candidate_num": L, "metrics": ["ndcg"], "metrics_topn": [1, 3]}
    torch.manual_seed(0)
    random.seed(0)
    model = la.NavieAlgorithm(types.SimpleNamespace(feature_size=F), settings)
    feed = ClickSimulationFeed(model, B, "click_model_json=%s,device_batches=True,%s" % (p, hp))
    return ds, model, feed, settings


@pytest.mark.parametrize("check_validation", [False, True])
def test_device_click_batches_have_the_reference_distribution(check_validation, tmp_path):
    """csrc/sampling.cu click_batch_kernel: queries uniform (conditioned on >= 1 click under check_validation), clicks
    ~ Bernoulli(exam_prob[l] * click_prob[label]) (PAD positions count as label 0, as in the reference), doc ids and
    PADs copied from the initial list."""
    nq, L, F, B = 24, 8, 12, 262144
    ds, model, feed, _ = _feed_arrays(tmp_path, nq, L, F, B)
    f, info = feed.get_batch(ds, check_validation=check_validation)
    init, rel, feats = feed._arrays(ds)
    q = np.asarray(info["rank_list_idxs"])
    docid = np.stack([f["docid_input%d" % l] for l in range(L)], axis=1).astype(np.int64)     # materialises
    clicks = np.stack([f["label%d" % l] for l in range(L)], axis=1)
    n_rows = feats.shape[0]
    assert f["letor_features"].shape == (n_rows, F) and docid.shape == (B, L)
    assert np.array_equal(docid, np.where(init[q] >= 0, init[q], n_rows))
    exam = np.asarray(PBM["exam_prob"])[np.minimum(np.arange(L), 9)]
    p = exam[None, :] * np.asarray(PBM["click_prob"])[rel.astype(int)]                         # [nq, L]
    p_any = 1.0 - np.prod(1.0 - p, axis=1)
    want_q = p_any / p_any.sum() if check_validation else np.full(nq, 1.0 / nq)
    got_q = np.bincount(q, minlength=nq) / float(B)
    assert np.all(np.abs(got_q - want_q) <= 5 * np.sqrt(want_q * (1 - want_q) / B) + 1e-6)
    if check_validation:
        assert np.all(clicks.sum(axis=1) > 0)
    else:
        for qi in range(nq):
            sel = q == qi
            got = clicks[sel].mean(axis=0)
            assert np.all(np.abs(got - p[qi]) <= 5 * np.sqrt(p[qi] * (1 - p[qi]) / sel.sum()) + 1e-6), qi


@pytest.mark.parametrize("name", ["cascade_model", "user_browsing_model"])
def test_device_sequential_click_models_match_the_host_samplers(name, tmp_path):
    """click_batch_kernel with click_model = 1 / 2 (lane 0 walks the list): per (query, position) click rates and the rate
    of a click given the previous click position against the array-form host samplers of the same models (which
    tests/test_click_feed.py checks against the reference's own samplers), 5 sigma"""
    import ultra_pytorch_b200.learning_algorithm as la
    from ultra_pytorch_b200.input_layer import ClickSimulationFeed
    from ultra_pytorch_b200.input_layer.click_simulation_feed import load_click_model
    la.B200Algorithm.VERBOSE = False
    nq, L, F, B = 6, 12, 8, 262144
    ds = FakeData(nq, L, F, ragged=False)
    if name == "cascade_model":
        desc = {"model_name": name, "eta": 1.0, "click_prob": [0.1, 0.16, 0.28, 0.52, 1.0], "exam_prob": [1.0] * 10}
    else:
        desc = {"model_name": name, "eta": 1.0, "click_prob": [0.1, 0.16, 0.28, 0.52, 1.0],
                "exam_prob": [[pow(x, 1.0) for x in row] for row in
                              __import__("ultra_pytorch_b200.input_layer.click_simulation_feed", fromlist=["x"])
                              ._UserBrowsingModel.ORIGINAL_RD_EXAM_TABLE]}
    p = os.path.join(str(tmp_path), "cm.json")
    with open(p, "w") as f:
        json.dump(desc, f)
    settings = {"learning_algorithm_hparams": "", "ranking_model": "ultra_pytorch_b200.ranking_model.DNN",
                "ranking_model_hparams": "hidden_layer_sizes=[32, 16]", "selection_bias_cutoff": L,
                "max_candidate_num": L, "metrics": ["ndcg"], "metrics_topn": [1, 3]}
    torch.manual_seed(0)
    random.seed(0)
    model = la.NavieAlgorithm(types.SimpleNamespace(feature_size=F), settings)
    feed = ClickSimulationFeed(model, B, "click_model_json=%s,device_batches=True" % p)
    f, info = feed.get_batch(ds, check_validation=False)
    q = np.asarray(info["rank_list_idxs"])
    clicks = np.stack([f["label%d" % l] for l in range(L)], axis=1).astype(np.float64)
    _, rel, _ = feed._arrays(ds)
    host = load_click_model(desc).sample(rel[q], np.random.default_rng(5))

    def prev_click(c):
        idx = np.where(c > 0, np.arange(L)[None, :], -1)
        run = np.maximum.accumulate(idx, axis=1)
        return np.concatenate([np.full((c.shape[0], 1), -1), run[:, :-1]], axis=1)
    pd_, ph = prev_click(clicks), prev_click(host)
    for qi in range(nq):
        sel = q == qi
        a, b = clicks[sel].mean(0), host[sel].mean(0)
        se = np.sqrt((a * (1 - a) + b * (1 - b)) / sel.sum()) + 1e-9
        assert (np.abs(a - b) <= 5 * se + 1e-3).all(), (name, qi, a, b)
    for r in range(1, L):
        for last in (-1, r - 1, r - 3):
            so, sh = pd_[:, r] == last, ph[:, r] == last
            if so.sum() < 2000 or sh.sum() < 2000:
                continue
            # pooled over the queries: same query mix on both sides (q is shared)
            a, b = clicks[so, r].mean(), host[sh, r].mean()
            se = np.sqrt(a * (1 - a) / so.sum() + b * (1 - b) / sh.sum()) + 1e-9
            assert abs(a - b) <= 5 * se + 3e-3, (name, r, last, a, b)
    if name == "cascade_model":
        assert (clicks.sum(axis=1) <= 1).all()


def test_training_from_device_batches_equals_training_from_their_host_copy(tmp_path):
    import ultra_pytorch_b200.learning_algorithm as la
    nq, L, F, B = 120, 10, 136, 64
    ds, model_a, feed, settings = _feed_arrays(tmp_path, nq, L, F, B)
    torch.manual_seed(0)
    model_b = la.NavieAlgorithm(types.SimpleNamespace(feature_size=F), settings)
    assert torch.equal(model_a.engine.params, model_b.engine.params)
    for step in range(6):
        f, _ = feed.get_batch(ds, check_validation=True)
        host = {k: f[k] for k in f.keys()}                      # plain dict, numpy arrays (+ the resident view)
        la_, _, _ = model_a.train(f)
        assert model_a.last_h2d_bytes == 0
        lb, _, _ = model_b.train(host)
        assert la_ == lb
    assert torch.equal(model_a.engine.params, model_b.engine.params)
