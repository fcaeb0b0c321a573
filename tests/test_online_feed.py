"""Tests of the StochasticOnlineSimulationFeed drop-in.

CPU: the host side (batch assembly over max_candidate_num, re-ordering, click simulation, feed format) with the
device sampler replaced by a numpy Gumbel-top-k stand-in.  GPU: the Plackett-Luce sampling kernel itself
(permutation frequencies against the closed-form Plackett-Luce probabilities, the large-tau limit, PAD handling,
reproducibility) and the feed driven by a real B200 learning algorithm."""
import itertools
import json
import os
import random
import types

import numpy as np
import pytest
import torch

from tests.test_click_feed import PBM, FakeData


def _model(L_train, L_max, F, scores_fn=None):
    m = types.SimpleNamespace(rank_list_size=L_train, max_candidate_num=L_max, feature_size=F,
                              letor_features_name="letor_features", hparams=types.SimpleNamespace(),
                              docid_inputs_name=["docid_input%d" % i for i in range(L_max)],
                              labels_name=["label%d" % i for i in range(L_max)])
    return m


def _feed_cls_with_host_sampler():
    from ultra_pytorch_b200.input_layer import StochasticOnlineSimulationFeed

    class HostSampled(StochasticOnlineSimulationFeed):
        """test stand-in for the device sampler: scores = -position (so a huge tau keeps the order), numpy Gumbel"""
        score_of = None

        def _sample_permutations(self, input_feed):
            L = self.max_candidate_num
            docid = np.stack([input_feed[self.model.docid_inputs_name[l]] for l in range(L)], axis=1)
            n_docs = len(input_feed[self.model.letor_features_name])
            valid = docid < n_docs
            scores = self.score_of(input_feed, docid)
            key = self.hparams.tau * scores + self.rng.gumbel(size=scores.shape)
            key = np.where(valid, key, -np.inf)
            perm = np.argsort(-key, axis=1, kind="stable")
            return perm
    return HostSampled


def _make(tmp_path, L_train, L_max, F, B, hp=""):
    p = os.path.join(str(tmp_path), "pbm.json")
    with open(p, "w") as f:
        json.dump(PBM, f)
    cls = _feed_cls_with_host_sampler()
    return cls(_model(L_train, L_max, F), B, ("click_model_json=%s," % p) + hp)


def test_online_feed_format_reordering_and_pads(tmp_path):
    L_train, L_max, F, B = 4, 7, 5, 32
    ds = FakeData(40, L_max, F)
    random.seed(1)
    feed = _make(tmp_path, L_train, L_max, F, B, "oracle_mode=True,tau=1.0")
    feed.score_of = lambda f, docid: np.where(docid < len(f["letor_features"]), docid % 3, 0).astype(np.float64)
    f, info = feed.get_batch(ds, check_validation=True)
    feats = f["letor_features"]
    n_docs = feats.shape[0]
    docid = np.stack([f["docid_input%d" % l] for l in range(L_max)], axis=1)
    label = np.stack([f["label%d" % l] for l in range(L_max)], axis=1)
    b = docid.shape[0]
    assert docid.dtype == np.float32 and label.dtype == np.float32 and feats.dtype == np.float64
    assert 0 < b <= B and len(info["rank_list_idxs"]) == B
    base = 0
    idx = [i for i in info["rank_list_idxs"] if sum(ds.labels[i]) != 0]      # check_validation drops all-zero lists
    assert len(idx) == b
    for r, q in enumerate(idx):
        n = sum(1 for x in ds.initial_list[q] if x >= 0)
        # the real documents of the list are a permutation of base .. base+n-1, pads (== n_docs) stay behind them
        assert sorted(docid[r, :n].astype(int).tolist()) == list(range(base, base + n))
        assert np.all(docid[r, n:] == n_docs)
        # features were appended list by list in the original order
        assert np.array_equal(feats[base:base + n], np.asarray(ds.features)[ds.initial_list[q][:n]])
        # oracle mode: the label of a shown document is its true label for the first rank_list_size ranks, 0 behind
        true = {base + x: ds.labels[q][x] for x in range(n)}
        for j in range(L_max):
            want = true[int(docid[r, j])] if (j < n and j < L_train) else 0.0
            assert label[r, j] == want
        base += n
    assert base == n_docs


def test_online_feed_large_tau_sorts_by_score_and_clicks_follow_the_click_model(tmp_path):
    L_train, L_max, F, B = 5, 5, 3, 4000
    ds = FakeData(30, L_max, F, ragged=False)
    random.seed(2)
    feed = _make(tmp_path, L_train, L_max, F, B, "tau=1000.0")
    feed.score_of = lambda f, docid: -docid.astype(np.float64)              # original order is the best order
    f, info = feed.get_batch(ds, check_validation=False)
    docid = np.stack([f["docid_input%d" % l] for l in range(L_max)], axis=1)
    assert np.all(np.diff(docid, axis=1) > 0)                               # order kept
    clicks = np.stack([f["label%d" % l] for l in range(L_max)], axis=1)
    true = np.asarray([ds.labels[i] for i in info["rank_list_idxs"]])
    p = np.asarray(PBM["exam_prob"])[:L_max][None, :] * np.asarray(PBM["click_prob"])[true.astype(int)]
    # click frequencies per (position, label) agree with exam_prob * click_prob within 4 sigma
    for l in range(L_max):
        for y in range(5):
            sel = true[:, l] == y
            if sel.sum() < 200:
                continue
            want = p[sel, l][0]
            got = clicks[sel, l].mean()
            assert abs(got - want) <= 4 * np.sqrt(want * (1 - want) / sel.sum()) + 1e-9, (l, y, got, want)


def test_online_feed_check_validation_redraws_clickless_lists(tmp_path):
    L_train, L_max, F, B = 3, 6, 3, 2000
    ds = FakeData(50, L_max, F)
    random.seed(3)
    feed = _make(tmp_path, L_train, L_max, F, B)
    feed.score_of = lambda f, docid: np.zeros(docid.shape)
    f, _ = feed.get_batch(ds, check_validation=True)
    clicks = np.stack([f["label%d" % l] for l in range(L_max)], axis=1)
    assert np.all(clicks[:, L_train:] == 0)
    # with up to 100 re-draws a list without clicks is essentially impossible unless its shown labels cannot be clicked
    assert (clicks.sum(axis=1) > 0).mean() > 0.999


def test_online_feed_with_a_cascade_click_model(tmp_path):
    """sequential click models (click_models.py:112-236) in the online feed: with the cascade model every validated
    list carries exactly one click, always on a real position of the prefix"""
    L_train, L_max, F, B = 4, 6, 3, 500
    ds = FakeData(50, L_max, F)
    random.seed(5)
    p = os.path.join(str(tmp_path), "cascade.json")
    with open(p, "w") as f:
        json.dump({"model_name": "cascade_model", "eta": 1.0, "click_prob": [0.1, 0.16, 0.28, 0.52, 1.0],
                   "exam_prob": [1.0] * 10}, f)
    feed = _feed_cls_with_host_sampler()(_model(L_train, L_max, F), B, "click_model_json=%s" % p)
    feed.score_of = lambda f_, docid: np.zeros(docid.shape)
    f, _ = feed.get_batch(ds, check_validation=True)
    clicks = np.stack([f["label%d" % l] for l in range(L_max)], axis=1)
    docid = np.stack([f["docid_input%d" % l] for l in range(L_max)], axis=1)
    n_docs = len(f["letor_features"])
    assert np.all(clicks[:, L_train:] == 0)
    assert np.all(clicks[docid >= n_docs] == 0)
    assert (clicks.sum(axis=1) == 1).mean() > 0.999


# ------------------------------------------------------------------------------------------------------------------
# GPU: the sampling kernel and the feed on a real B200 algorithm
# ------------------------------------------------------------------------------------------------------------------
def _pl_probability(scores, perm, tau):
    w = np.exp(tau * (np.asarray(scores, dtype=np.float64) - max(scores)))
    p, rest = 1.0, list(range(len(scores)))
    for i in perm:
        p *= w[i] / sum(w[j] for j in rest)
        rest.remove(i)
    return p


@pytest.mark.gpu
def test_pl_sample_matches_plackett_luce_distribution():
    from ultra_pytorch_b200.engine import RankerEngine
    eng = RankerEngine(4, [])
    scores = [0.3, 1.5, -0.4, 0.9]
    tau, B, L = 1.3, 400000, 4
    s = torch.tensor(scores, device="cuda").repeat(B, 1).contiguous()
    perm = eng.pl_sample(s, None, 0, tau, seed=1234, offset=7).cpu().numpy()
    assert np.array_equal(np.sort(perm, axis=1), np.tile(np.arange(L), (B, 1)))
    code = (perm * (L ** np.arange(L))[None, :]).sum(axis=1)
    for pm in itertools.permutations(range(L)):
        c = sum(v * L ** i for i, v in enumerate(pm))
        want = _pl_probability(scores, pm, tau)
        got = float((code == c).mean())
        assert abs(got - want) <= 5 * np.sqrt(want * (1 - want) / B) + 1e-6, (pm, got, want)
    # same (seed, offset) -> same permutations; another offset -> different ones
    again = eng.pl_sample(s, None, 0, tau, seed=1234, offset=7).cpu().numpy()
    other = eng.pl_sample(s, None, 0, tau, seed=1234, offset=8).cpu().numpy()
    assert np.array_equal(perm, again) and not np.array_equal(perm, other)


@pytest.mark.gpu
def test_pl_sample_large_tau_is_a_descending_sort_and_pads_stay():
    from ultra_pytorch_b200.engine import RankerEngine
    eng = RankerEngine(4, [])
    rs = np.random.RandomState(0)
    B, L, n_docs = 64, 300, 5000
    # well separated scores (spacing 0.5): with tau = 1e3 the Gumbel noise (|g| < ~20) cannot swap neighbours
    s = np.stack([0.5 * rs.permutation(L) - 40.0 for _ in range(B)]).astype(np.float32)
    lens = rs.randint(1, L + 1, size=B)
    lens[0], lens[1] = L, 1
    docid = np.full((L, B), n_docs, dtype=np.int32)
    for b in range(B):
        docid[:lens[b], b] = rs.randint(0, n_docs, size=lens[b])
    perm = eng.pl_sample(torch.from_numpy(s).cuda(), torch.from_numpy(docid).cuda(), n_docs, 1e3, seed=5).cpu().numpy()
    for b in range(B):
        n = lens[b]
        assert np.array_equal(perm[b, :n], np.argsort(-s[b, :n], kind="stable"))
        assert np.array_equal(perm[b, n:], np.arange(n, L))


@pytest.mark.gpu
def test_online_feed_with_b200_algorithm(tmp_path):
    import ultra_pytorch_b200.learning_algorithm as la
    from ultra_pytorch_b200.input_layer import StochasticOnlineSimulationFeed
    la.B200Algorithm.VERBOSE = False
    L_train, L_max, F, B = 5, 9, 12, 64
    ds = FakeData(80, L_max, F)
    p = os.path.join(str(tmp_path), "pbm.json")
    with open(p, "w") as f:
        json.dump(PBM, f)
    settings = {"learning_algorithm_hparams": "", "ranking_model": "ultra_pytorch_b200.ranking_model.DNN",
                "ranking_model_hparams": "hidden_layer_sizes=[16, 8]", "selection_bias_cutoff": L_train,
                "max_candidate_num": L_max, "metrics": ["ndcg"], "metrics_topn": [1, 3]}
    torch.manual_seed(0)
    random.seed(0)
    model = la.DLA(types.SimpleNamespace(feature_size=F), settings)
    feed = StochasticOnlineSimulationFeed(model, B, "click_model_json=%s,tau=1.0" % p)
    for step in range(3):
        f, info = feed.get_batch(ds, check_validation=True)
        n_docs = f["letor_features"].shape[0]
        docid = np.stack([f["docid_input%d" % l] for l in range(L_max)], axis=1)
        label = np.stack([f["label%d" % l] for l in range(L_max)], axis=1)
        real = docid < n_docs
        # every list keeps its documents (a permutation), pads stay at the end, clicks only in the first L_train ranks
        assert np.all(np.diff(real.astype(int), axis=1) <= 0)
        assert sorted(docid[real].astype(int).tolist()) == list(range(n_docs))
        assert np.all(label[:, L_train:] == 0) and np.all(label[~real] == 0)
        loss, _, _ = model.train(f)
        assert np.isfinite(loss)
