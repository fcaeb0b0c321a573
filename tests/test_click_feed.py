"""CPU tests of the vectorised ClickSimulationFeed drop-in against the reference feed:
bit-identical on the deterministic paths, equal in distribution on the sampled clicks."""
import json
import os
import random
import types

import numpy as np
import pytest

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PBM = {"click_prob": [0.1, 0.16, 0.28, 0.52, 1.0], "eta": 1.0,
       "exam_prob": [0.68, 0.61, 0.48, 0.34, 0.28, 0.2, 0.11, 0.1, 0.08, 0.06], "model_name": "position_biased_model"}


class FakeData(object):
    """Minimal stand-in for ultra.utils.data_utils.Raw_data after pad() (data_utils.py:476-498)."""

    def __init__(self, nq, L, F, seed=0, ragged=True):
        rs = np.random.RandomState(seed)
        self.feature_size, self.rank_list_size = F, L
        self.features, self.initial_list, self.labels = [], [], []
        doc = 0
        for _ in range(nq):
            n = int(rs.randint(max(1, L // 2), L + 1)) if ragged else L
            self.features.extend(rs.uniform(-1, 1, size=(n, F)).tolist())
            self.initial_list.append(list(range(doc, doc + n)) + [-1] * (L - n))
            self.labels.append(rs.randint(0, 5, size=n).astype(float).tolist())
            doc += n
        self.features.append([0.0] * F)


def _model(L, F):
    return types.SimpleNamespace(rank_list_size=L, feature_size=F, letor_features_name="letor_features",
                                 docid_inputs_name=["docid_input%d" % i for i in range(L)],
                                 labels_name=["label%d" % i for i in range(L)])


def _ours(tmp_path, L, F, B, hp=""):
    from ultra_pytorch_b200.input_layer import ClickSimulationFeed
    p = os.path.join(str(tmp_path), "pbm.json")
    with open(p, "w") as f:
        json.dump(PBM, f)
    return ClickSimulationFeed(_model(L, F), B, ("click_model_json=%s," % p) + hp), p


def test_feed_format_and_pad_convention(tmp_path):
    L, F, B = 7, 5, 16
    ds = FakeData(40, L, F)
    feed, _ = _ours(tmp_path, L, F, B)
    random.seed(0)
    f, info = feed.get_batch(ds, check_validation=True)
    feats = f["letor_features"]
    n_docs = feats.shape[0]
    assert feats.dtype == np.float64 and feats.shape[1] == F
    docid = np.stack([f["docid_input%d" % l] for l in range(L)], axis=1)
    clicks = np.stack([f["label%d" % l] for l in range(L)], axis=1)
    assert docid.dtype == np.float32 and clicks.dtype == np.float32 and docid.shape == (B, L)
    assert (clicks.sum(axis=1) > 0).all()                       # check_validation drops click-less lists
    real = docid != n_docs
    assert sorted(docid[real].astype(int).tolist()) == list(range(n_docs))   # every real doc referenced exactly once
    for b, q in enumerate(info["rank_list_idxs"]):             # features are the data set's rows, in list order
        ids = [d for d in ds.initial_list[q] if d >= 0]
        got = feats[docid[b][real[b]].astype(int)]
        assert np.array_equal(got, np.asarray(ds.features)[ids])
    assert len(info["input_list"]) == B


def test_click_rates_follow_the_position_biased_model(tmp_path):
    L, F, B = 12, 3, 512
    ds = FakeData(300, L, F, ragged=False)
    feed, _ = _ours(tmp_path, L, F, B)
    feed.rng = np.random.default_rng(1)
    init, labels, _ = feed._arrays(ds)
    tot, exp, n = np.zeros(L), np.zeros(L), 0
    for _ in range(40):
        f, info = feed.get_batch(ds, check_validation=False)
        clicks = np.stack([f["label%d" % l] for l in range(L)], axis=1)
        lab = labels[np.asarray(info["rank_list_idxs"])]
        exam = np.asarray(PBM["exam_prob"])[np.minimum(np.arange(L), 9)]
        exp += (exam[None, :] * np.asarray(PBM["click_prob"])[lab.astype(int)]).sum(axis=0)
        tot += clicks.sum(axis=0)
        n += B
    sigma = np.sqrt(np.maximum(exp, 1.0))
    assert (np.abs(tot - exp) < 5 * sigma).all(), (tot, exp)


@pytest.mark.parametrize("oracle_mode", [True, False])
def test_matches_reference_feed(tmp_path, oracle_mode):
    if not ref_shim.available():
        pytest.skip("oracle/_ref not installed")
    ultra = ref_shim.load()
    L, F, B = 9, 6, 8
    ds = FakeData(30, L, F, seed=3)
    ours, pbm_path = _ours(tmp_path, L, F, B, "oracle_mode=%s" % oracle_mode)
    ref = ultra.input_layer.ClickSimulationFeed(_model(L, F), B, "click_model_json=%s,oracle_mode=%s"
                                                % (pbm_path, oracle_mode))
    if oracle_mode:
        # deterministic: sequential batches are bit-identical, including the ragged last batch
        for index in (0, 8, 24):
            a, ia = ours.get_next_batch(index, ds, check_validation=False)
            b, ib = ref.get_next_batch(index, ds, check_validation=False)
            assert sorted(a.keys()) == sorted(b.keys())
            for k in a:
                assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
            assert len(ia["input_list"]) == len(ib["input_list"])
        a, _ = ours.get_data_by_index(ds, 5)
        b, _ = ref.get_data_by_index(ds, 5)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    else:
        # sampled clicks: same per-position click rate as the reference's own sampler (5 sigma)
        random.seed(7)
        ours.rng = np.random.default_rng(7)
        ta, tb, n = np.zeros(L), np.zeros(L), 0
        for _ in range(150):
            a, _ = ours.get_next_batch(0, ds, check_validation=False)
            b, _ = ref.get_next_batch(0, ds, check_validation=False)
            ta += np.stack([a["label%d" % l] for l in range(L)], axis=1).sum(axis=0)
            tb += np.stack([b["label%d" % l] for l in range(L)], axis=1).sum(axis=0)
            n += B
        sigma = np.sqrt(np.maximum(ta + tb, 1.0))
        assert (np.abs(ta - tb) < 5 * sigma).all(), (ta, tb)


def test_resident_feed_is_a_valid_reference_format_feed_with_global_ids(tmp_path):
    """resident_features=True (input_layer/resident.py): `letor_features` is a zero-copy view of the data set's whole
    matrix, doc ids are global row ids and PAD = number of rows - the gathered rows equal the per-batch copy of the
    default mode, list for list."""
    from ultra_pytorch_b200.input_layer.resident import ResidentFeatures
    L, F, B = 7, 5, 16
    ds = FakeData(40, L, F)
    plain, _ = _ours(tmp_path, L, F, B, "oracle_mode=True")
    res, _ = _ours(tmp_path, L, F, B, "oracle_mode=True,resident_features=True")
    fa, _ = plain.get_next_batch(3, ds)
    fb, _ = res.get_next_batch(3, ds)
    assert isinstance(fb["letor_features"], ResidentFeatures) and isinstance(fb["letor_features"], np.ndarray)
    assert fb["letor_features"].shape == (len(ds.features), F)
    assert np.shares_memory(fb["letor_features"], res._features)
    da = np.stack([fa["docid_input%d" % l] for l in range(L)], axis=1).astype(int)
    db = np.stack([fb["docid_input%d" % l] for l in range(L)], axis=1).astype(int)
    na, nb = fa["letor_features"].shape[0], fb["letor_features"].shape[0]
    assert np.array_equal(da == na, db == nb)                                   # same PAD pattern
    # what the reference's get_ranking_scores would gather (zero PAD row appended at index n, base_algorithm.py:148-152)
    ga = np.concatenate([fa["letor_features"], np.zeros((1, F))])[da]
    gb = np.concatenate([np.asarray(fb["letor_features"]), np.zeros((1, F))])[db]
    assert np.array_equal(ga, gb)
    for l in range(L):
        assert np.array_equal(fa["label%d" % l], fb["label%d" % l])
    # a second batch re-uses the same view object (the engine keys its device copy on it)
    fc, _ = res.get_next_batch(0, ds)
    assert fc["letor_features"] is fb["letor_features"]


def test_device_feed_behaves_like_the_reference_dict_on_the_host():
    """input_layer/resident.py DeviceFeed: consumers that are not B200 algorithms see a plain reference-format dict (the
    first access to an id / label key copies the device batch to the host once).  CPU tensors stand in for the device."""
    import torch
    from ultra_pytorch_b200.input_layer.click_simulation_feed import _LazyInfo
    from ultra_pytorch_b200.input_layer.resident import DeviceFeed, ResidentFeatures
    L, B, F = 3, 4, 5
    m = _model(L, F)
    feats = ResidentFeatures(np.arange(50.0).reshape(10, F))
    docid = torch.tensor([[0, 1, 2, 3], [4, 5, 10, 7], [8, 10, 10, 9]], dtype=torch.int32)       # [L, B], PAD id = 10
    labels = torch.tensor([[1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1]], dtype=torch.float32)     # [B, L]
    qidx = torch.tensor([3, 1, 2, 0], dtype=torch.int32)
    f = DeviceFeed(m, feats, docid, labels, qidx, 10)
    assert not f._materialised and f["letor_features"] is feats and not f._materialised        # no copy needed
    assert "docid_input1" in f and "label2" in f and len(f) == 1 + 2 * L
    assert f["docid_input1"].dtype == np.float32 and f["docid_input1"].tolist() == [4.0, 5.0, 10.0, 7.0]
    assert f._materialised and f["label0"].tolist() == [1.0, 0.0, 1.0, 0.0]
    assert sorted(f.keys()) == sorted(["letor_features"] + m.docid_inputs_name + m.labels_name)
    plain = dict(f.items())
    assert set(plain) == set(f.keys()) and plain["label2"].tolist() == [0.0, 0.0, 0.0, 1.0]
    info = _LazyInfo(f)
    assert info["rank_list_idxs"] == [3, 1, 2, 0]
    assert info["input_list"].tolist() == docid.numpy().T.tolist() and info["click_list"].shape == (B, L)
    assert info["letor_features"] is feats


@pytest.mark.parametrize("use_max", [True, False])
def test_direct_label_feed_is_bit_identical_to_the_reference(use_max):
    """input_layer/direct_label_feed.py vs ultra.input_layer.DirectLabelFeed: get_batch (same random.seed),
    get_next_batch and get_data_by_index, with and without check_validation."""
    if not ref_shim.available():
        pytest.skip("oracle/_ref not installed")
    ultra = ref_shim.load()
    from ultra_pytorch_b200.input_layer import DirectLabelFeed
    L_train, L_max, F, B = 5, 9, 6, 8
    ds = FakeData(30, L_max, F, seed=4)
    ds.labels[3] = [0.0] * len(ds.labels[3])            # a list without any relevant document
    m = _model(L_max, F)
    m.rank_list_size, m.max_candidate_num = L_train, L_max
    hp = "use_max_candidate_num=%s" % use_max
    ours, ref = DirectLabelFeed(m, B, hp), ultra.input_layer.DirectLabelFeed(m, B, hp)
    assert ours.rank_list_size == ref.rank_list_size

    def same(a, b):
        assert sorted(a.keys()) == sorted(b.keys())
        for k in a:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k

    for cv in (False, True):
        for seed in (0, 1, 2):
            random.seed(seed)
            a, ia = ours.get_batch(ds, check_validation=cv)
            random.seed(seed)
            b, ib = ref.get_batch(ds, check_validation=cv)
            same(a, b)
            assert ia["rank_list_idxs"] == ib["rank_list_idxs"]
        for index in (0, 3, 24):
            same(ours.get_next_batch(index, ds, check_validation=cv)[0], ref.get_next_batch(index, ds, check_validation=cv)[0])
    # the reference's get_data_by_index raises (it calls a method that does not exist, direct_label_feed.py:247); ours is
    # the one-list case of get_next_batch
    with pytest.raises(AttributeError):
        ref.get_data_by_index(ds, 5)
    one, _ = ours.get_data_by_index(ds, 5)
    ours_b1 = DirectLabelFeed(m, 1, hp)
    same(one, ours_b1.get_next_batch(5, ds)[0])
    with pytest.raises(NotImplementedError):
        ours.get_batch(ds, data_format="ULTRE")


@pytest.mark.parametrize("name", ["user_browsing_model", "cascade_model"])
def test_sequential_click_models_match_the_reference_in_distribution(name):
    """UserBrowsingModel / CascadeModel (click_models.py:112-236) in array form: per-position click rates AND the rate of
    a click at position r given the previous click position (the dependence these models add) against the reference's
    own sampleClicksForOneList, 5 sigma on 40 000 lists each."""
    ultra = ref_shim.load()
    from ultra_pytorch_b200.input_layer.click_simulation_feed import load_click_model
    cm_ref = {"user_browsing_model": ultra.utils.click_models.UserBrowsingModel,
              "cascade_model": ultra.utils.click_models.CascadeModel}[name](0.1, 1.0, 4, 1.0)
    desc = cm_ref.getModelJson()
    cm = load_click_model(json.loads(json.dumps(desc)))
    L, n = 12, 40000
    labels = np.array([3, 0, 4, 1, 2, 0, 0, 4, 1, 3, 2, 0], dtype=np.float64)
    ours = cm.sample(np.tile(labels, (n, 1)), np.random.default_rng(1))
    random.seed(2)
    ref = np.array([cm_ref.sampleClicksForOneList(list(labels))[0] for _ in range(n)], dtype=np.float64)

    def prev_click(c):
        idx = np.where(c > 0, np.arange(L)[None, :], -1)
        run = np.maximum.accumulate(idx, axis=1)
        return np.concatenate([np.full((c.shape[0], 1), -1), run[:, :-1]], axis=1)
    for a, b, what in ((ours.mean(0), ref.mean(0), "marginal"),):
        se = np.sqrt((a * (1 - a) + b * (1 - b)) / n) + 1e-9
        assert (np.abs(a - b) <= 5 * se + 1e-3).all(), (name, what, a, b)
    po, pr = prev_click(ours), prev_click(ref)
    for r in range(1, L):
        for last in (-1, r - 1, r - 2):
            so, sr = po[:, r] == last, pr[:, r] == last
            if so.sum() < 500 or sr.sum() < 500:
                continue
            a, b = ours[so, r].mean(), ref[sr, r].mean()
            se = np.sqrt(a * (1 - a) / so.sum() + b * (1 - b) / sr.sum()) + 1e-9
            assert abs(a - b) <= 5 * se + 2e-3, (name, r, last, a, b)
    # setExamProb(eta) rebuilds the same table as the reference
    cm.setExamProb(0.5)
    cm_ref.setExamProb(0.5)
    assert np.allclose(np.asarray(cm.exam_prob[-1]), np.asarray(cm_ref.exam_prob[-1]))


def test_feed_with_a_cascade_click_model(tmp_path):
    from ultra_pytorch_b200.input_layer import ClickSimulationFeed
    L, F, B = 7, 5, 64
    ds = FakeData(80, L, F)
    p = os.path.join(str(tmp_path), "cascade.json")
    with open(p, "w") as f:
        json.dump({"model_name": "cascade_model", "eta": 1.0, "click_prob": [0.1, 0.16, 0.28, 0.52, 1.0],
                   "exam_prob": [1.0] * 10}, f)
    feed = ClickSimulationFeed(_model(L, F), B, "click_model_json=%s" % p)
    random.seed(0)
    f, _ = feed.get_batch(ds, check_validation=True)
    clicks = np.stack([f["label%d" % l] for l in range(L)], axis=1)
    assert (clicks.sum(axis=1) == 1).all()          # cascade: exactly one click per (validated) list
