"""Vectorised drop-in for `ultra.input_layer.ClickSimulationFeed`
(reference: ultra/input_layer/click_simulation_feed.py:23-294, click model ultra/utils/click_models.py:68-110).

The reference assembles every batch with pure-Python loops (one `sampleClicksForOneList` call per query, one list
comprehension per (query, position), `np.array` of a list of feature lists): 67 ms per 256-query batch at config 2,
430 ms at L = 200 - more than the whole reference train() step and ~100x the B200 step.  This class produces the
SAME `input_feed` (keys, dtypes, shapes, PAD convention, click semantics) with numpy array operations on a cached
array view of the data set.  Random numbers come from a numpy Generator seeded from Python's `random` module at
construction (the reference draws from `random` directly), so runs stay reproducible under `random.seed(...)`,
but the click streams are equal in distribution, not draw for draw.  Deterministic paths (oracle_mode, sequential
`get_next_batch` with check_validation=False) are bit-identical to the reference (tests/test_click_feed.py).
"""
import json
import random

import numpy as np

from ..hparams import HParams
from .resident import DeviceFeed, ResidentFeatures

ORIGINAL_EXAM_PROB = [0.68, 0.61, 0.48, 0.34, 0.28, 0.20, 0.11, 0.10, 0.08, 0.06]   # click_models.py:76-77


def load_click_model(desc):
    """Array form of click_models.loadModelFromJson (click_models.py:7-16)."""
    name = desc.get('model_name', 'position_biased_model')
    if name == 'user_browsing_model':
        return _UserBrowsingModel(desc)
    if name == 'cascade_model':
        return _CascadeModel(desc)
    return _PositionBiasedModel(desc)


def _click_prob_of(labels, click_prob):
    """click_prob[int(label) if label > 0 else 0], last entry for labels beyond the table (click_models.py:98-104)"""
    rel = np.where(labels > 0, labels, 0).astype(np.int64)
    rel = np.where(rel < len(click_prob), rel, len(click_prob) - 1)
    return click_prob[rel]


class _SequentialClickModel(object):
    """Click models whose examination depends on earlier clicks of the same list: sampled position by position for all
    lists of the batch at once (the reference loops over lists AND positions, sampleClicksForOneList)."""
    position_independent = False

    def click_probability(self, labels):
        raise NotImplementedError("%s has no position-wise click probability (clicks depend on earlier clicks); the "
                                  "device click simulator and the online feed implement the position-biased model only"
                                  % type(self).__name__)


class _UserBrowsingModel(_SequentialClickModel):
    """Array form of UserBrowsingModel (click_models.py:112-185): P(exam) = table[rank][rank - last_click_rank - 1]."""
    DEVICE_CODE = 2                              # ub200_click_batch_model

    def device_exam_table(self):
        return np.ascontiguousarray(self._table)
    ORIGINAL_RD_EXAM_TABLE = [
        [1.0],
        [0.98, 1.0],
        [1.0, 0.62, 0.95],
        [1.0, 0.77, 0.42, 0.82],
        [1.0, 0.92, 0.55, 0.31, 0.69],
        [1.0, 0.96, 0.63, 0.4, 0.22, 0.54],
        [1.0, 0.99, 0.73, 0.46, 0.29, 0.17, 0.47],
        [1.0, 1.0, 0.89, 0.52, 0.35, 0.24, 0.14, 0.43],
        [1.0, 1.0, 0.95, 0.68, 0.4, 0.29, 0.19, 0.12, 0.41],
        [1.0, 1.0, 1.0, 0.96, 0.52, 0.36, 0.27, 0.18, 0.12, 0.43]]         # click_models.py:121-132

    def __init__(self, desc):
        self.eta = desc['eta']
        self.click_prob = np.asarray(desc['click_prob'], dtype=np.float64)
        self._set_table(desc['exam_prob'])

    def _set_table(self, rows):
        n = len(rows)
        self.exam_prob = [list(r) for r in rows]
        self._table = np.zeros((n, n), dtype=np.float64)
        for i, r in enumerate(rows):
            self._table[i, :len(r)] = r

    def setExamProb(self, eta):
        self.eta = eta
        self._set_table([[pow(x, eta) for x in row] for row in self.ORIGINAL_RD_EXAM_TABLE])

    def exam_probability(self, rank, last_click_rank):
        """getExamProb (click_models.py:174-185) for an array of last-click ranks"""
        n = self._table.shape[0]
        distance = rank - last_click_rank
        if rank < n:
            return self._table[rank, distance - 1]
        last = self._table[n - 1]
        idx = np.where(distance < n - 1, distance - 1, n - 2)
        return np.where(distance > rank, last[n - 1], last[idx])

    def sample(self, labels, rng):
        n, L = labels.shape
        clicks = np.zeros((n, L), dtype=np.float64)
        last = np.full(n, -1, dtype=np.int64)
        cp = _click_prob_of(labels, self.click_prob)
        u = rng.random((n, L))
        for r in range(L):
            c = u[:, r] < self.exam_probability(r, last) * cp[:, r]
            clicks[:, r] = c
            last = np.where(c, r, last)
        return clicks


class _CascadeModel(_SequentialClickModel):
    """Array form of CascadeModel (click_models.py:188-236): the user clicks at most once - every position is sampled
    with P = exam_prob[rank] * click_prob[label], positions after the first click report no click."""
    DEVICE_CODE = 1

    def device_exam_table(self):
        return np.asarray(self.exam_prob, dtype=np.float64)

    def __init__(self, desc):
        self.eta = desc['eta']
        self.click_prob = np.asarray(desc['click_prob'], dtype=np.float64)
        self.exam_prob = np.asarray(desc['exam_prob'], dtype=np.float64)

    def setExamProb(self, eta):
        self.eta = eta
        self.exam_prob = np.ones(10, dtype=np.float64)                     # click_models.py:195-197

    def sample(self, labels, rng):
        L = labels.shape[1]
        exam = self.exam_prob[np.minimum(np.arange(L), len(self.exam_prob) - 1)]
        c = rng.random(labels.shape) < exam[None, :] * _click_prob_of(labels, self.click_prob)
        before = np.cumsum(c, axis=1) - c                                   # clicks strictly before each position
        return (c & (before == 0)).astype(np.float64)


class _PositionBiasedModel(object):
    """Array form of PositionBiasedModel (click_models.py:68-110)."""
    position_independent = True
    DEVICE_CODE = 0

    def device_exam_table(self):
        return self.exam_prob

    def __init__(self, desc):
        self.eta = desc['eta']
        self.click_prob = np.asarray(desc['click_prob'], dtype=np.float64)
        self.exam_prob = np.asarray(desc['exam_prob'], dtype=np.float64)

    def setExamProb(self, eta):
        self.eta = eta
        self.exam_prob = np.power(np.asarray(ORIGINAL_EXAM_PROB, dtype=np.float64), eta)

    def click_probability(self, labels):
        """labels [n, L] (relevance, pads = 0) -> P(click) [n, L] = exam_prob[min(rank, last)] * click_prob[label]."""
        L = labels.shape[1]
        exam = self.exam_prob[np.minimum(np.arange(L), len(self.exam_prob) - 1)]
        rel = np.where(labels > 0, labels, 0).astype(np.int64)
        rel = np.where(rel < len(self.click_prob), rel, len(self.click_prob) - 1)
        return exam[None, :] * self.click_prob[rel]


class _LazyInfo(dict):
    """info_map of a device batch (click_simulation_feed.py:158-163): values are fetched from the device on first use."""

    def __init__(self, feed):
        dict.__init__(self)
        self._feed = feed

    def __missing__(self, key):
        f = self._feed
        if key == 'rank_list_idxs':
            v = f.query_idx.cpu().numpy().tolist()
        elif key == 'input_list':
            v = f.docid.cpu().numpy().T.astype(np.int64)
        elif key == 'click_list':
            v = f.labels.cpu().numpy().astype(np.float64)
        elif key == 'letor_features':
            v = dict.__getitem__(f, f.model.letor_features_name)
        else:
            raise KeyError(key)
        self[key] = v
        return v


class ClickSimulationFeed(object):
    MAX_SAMPLE_ROUND_NUM = 100

    @staticmethod
    def preprocess_data(data_set, hparam_str, exp_settings):
        return

    def __init__(self, model, batch_size, hparam_str):
        self.hparams = HParams(
            click_model_json='./example/ClickModel/pbm_0.1_1.0_4_1.0.json',    # click_simulation_feed.py:40-51
            oracle_mode=False,
            dynamic_bias_eta_change=0.0,
            dynamic_bias_step_interval=1000,
            resident_features=False,       # B200 extension: emit global doc ids + the whole matrix (resident.py)
            device_batches=False,          # B200 extension: sample queries + clicks on the GPU (implies resident)
        )
        print('Create simluated clicks feed')
        print(hparam_str)
        self.hparams.parse(hparam_str)
        self.click_model = None
        if not self.hparams.oracle_mode:
            with open(self.hparams.click_model_json) as fin:
                self.click_model = load_click_model(json.load(fin))
        self.start_index = 0
        self.count = 1
        self.rank_list_size = model.rank_list_size
        self.feature_size = model.feature_size
        self.batch_size = batch_size
        self.model = model
        self.global_batch_count = 0
        self.rng = np.random.default_rng(random.getrandbits(63))
        self._cache_key = None

    # ---- array view of the data set (built once per data set) ------------------------------------------
    def _arrays(self, data_set):
        key = (id(data_set), len(data_set.initial_list), len(data_set.features), self.rank_list_size)
        if self._cache_key != key:
            L = self.rank_list_size
            nq = len(data_set.initial_list)
            init = np.full((nq, L), -1, dtype=np.int64)
            labels = np.zeros((nq, L), dtype=np.float64)
            for i in range(nq):
                row = data_set.initial_list[i]
                n = min(len(row), L)
                init[i, :n] = row[:n]
                lab = data_set.labels[i]
                m = min(len(lab), n)
                labels[i, :m] = lab[:m]
            labels[init < 0] = 0.0                                         # click_simulation_feed.py:75-77
            self._init = init
            self._labels = labels
            self._features = np.asarray(data_set.features, dtype=np.float64)
            self._cache_key = key
        return self._init, self._labels, self._features

    def _simulate(self, labels):
        if self.hparams.oracle_mode:
            return labels.copy()
        if not self.click_model.position_independent:
            return self.click_model.sample(labels, self.rng)
        p = self.click_model.click_probability(labels)
        return (self.rng.random(labels.shape) < p).astype(np.float64)

    def _simulate_queries(self, idx):
        """_simulate(self._labels[idx]) with the click probabilities of the whole data set cached per (data set, eta):
        P(click) depends only on (query, position), so a batch is one row gather + one uniform draw."""
        if self.hparams.oracle_mode:
            return self._labels[idx].copy()
        if not self.click_model.position_independent:
            return self.click_model.sample(self._labels[idx], self.rng)
        key = (id(self._labels), float(self.click_model.eta), self.click_model.exam_prob.tobytes())
        if getattr(self, "_pclick_key", None) != key:
            self._pclick = self.click_model.click_probability(self._labels)
            self._pclick_key = key
        p = self._pclick[idx]
        return (self.rng.random(p.shape) < p).astype(np.float64)

    def _assemble(self, idx, clicks):
        """idx [b] query indices, clicks [b, L] -> (input_feed, info_map) in the reference's format."""
        init, _, features = self._init, self._labels, self._features
        L = self.rank_list_size
        rows = init[idx]                                                   # [b, L]
        real = rows >= 0
        if getattr(self.hparams, "resident_features", False):
            # global row ids into the data set's whole (device-resident) feature matrix; PAD id = number of rows
            if getattr(self, "_resident_src", None) is not features:
                self._resident_src = features
                self._resident_view = ResidentFeatures(features)
            letor_features = self._resident_view
            n_docs = features.shape[0]
            docid = np.where(real, rows, n_docs)
        else:
            n_real = real.sum(axis=1)
            base = np.concatenate([[0], np.cumsum(n_real)[:-1]])
            n_docs = int(n_real.sum())
            letor_features = features[rows[real]]                          # real docs, list by list, in order
            docid = np.where(real, base[:, None] + np.arange(L)[None, :], n_docs)   # base + x ; PAD id = n_docs
        input_feed = {self.model.letor_features_name: letor_features}
        docid_f = np.ascontiguousarray(docid.T.astype(np.float32))         # [L, b]
        label_f = np.ascontiguousarray(clicks.T.astype(np.float32))
        for l in range(L):
            input_feed[self.model.docid_inputs_name[l]] = docid_f[l]
            input_feed[self.model.labels_name[l]] = label_f[l]
        return input_feed, docid, n_docs

    def _check_list_size(self, data_set):
        if len(data_set.initial_list[0]) < self.rank_list_size:
            raise ValueError("Input ranklist length must be no less than the required list size,"
                             " %d != %d." % (len(data_set.initial_list[0]), self.rank_list_size))

    # ---- reference API ----------------------------------------------------------------------------------------
    # ---- N1 on the device: query sampling + click simulation + batch assembly in one kernel ------------------------
    def _device_batch(self, data_set, check_validation):
        import torch
        if not self.hparams.oracle_mode and not self.click_model.position_independent and self.rank_list_size > 256:
            raise NotImplementedError("device_batches=True samples %s for lists of up to 256 positions (drop "
                                      "device_batches: the host path has no limit, resident_features=True still applies)"
                                      % type(self.click_model).__name__)
        eng = getattr(self.model, "engine", None)
        if eng is None:
            raise TypeError("device_batches=True needs a B200 learning algorithm (the batch is assembled in its "
                            "device buffers)")
        init, labels, features = self._arrays(data_set)
        if getattr(self, "_resident_src", None) is not features:
            self._resident_src = features
            self._resident_view = ResidentFeatures(features)
        eng.ensure_resident(self._resident_view)
        dev = eng.device
        key = (id(init), id(labels))
        if getattr(self, "_dev_key", None) != key:
            self._dev_init = torch.from_numpy(init.astype(np.int32)).to(dev)
            self._dev_rel = torch.from_numpy(labels.astype(np.float32)).to(dev)
            L, B = self.rank_list_size, self.batch_size
            # a small ring of output buffers: a feed handed out earlier stays valid for a few more batches, and the
            # CUDA graphs of the training step (keyed on the buffer addresses) are reused
            self._dev_ring = [(torch.empty(L, B, dtype=torch.int32, device=dev),
                               torch.empty(B, L, dtype=torch.float32, device=dev),
                               torch.empty(B, dtype=torch.int32, device=dev)) for _ in range(4)]
            self._dev_key = key
            self._dev_cm_key = None
            self._dev_seed = random.getrandbits(63)
            self._dev_calls = 0
        oracle = bool(self.hparams.oracle_mode)
        if not oracle:
            exam = self.click_model.device_exam_table()
            cm_key = (float(self.click_model.eta), exam.tobytes())
            if self._dev_cm_key != cm_key:
                self._dev_exam = torch.from_numpy(exam.astype(np.float32)).to(dev)
                self._dev_cp = torch.from_numpy(self.click_model.click_prob.astype(np.float32)).to(dev)
                self._dev_cm_key = cm_key
        self._dev_calls += 1
        docid, lab, qidx = self._dev_ring[self._dev_calls % len(self._dev_ring)]
        eng.click_batch(self._dev_init, self._dev_rel, None if oracle else self._dev_exam,
                        None if oracle else self._dev_cp, oracle, bool(check_validation), 1000, features.shape[0],
                        self._dev_seed, self._dev_calls, docid, lab, qidx, click_model=self.click_model.DEVICE_CODE)
        return DeviceFeed(self.model, self._resident_view, docid, lab, qidx, features.shape[0])

    def get_batch(self, data_set, check_validation=False, data_format="ULTRA"):
        """Random batch for training (click_simulation_feed.py:101-174): draws queries uniformly with replacement and,
        with check_validation, keeps only lists with at least one click until batch_size lists are collected."""
        self._check_list_size(data_set)
        if getattr(self.hparams, "device_batches", False):
            feed = self._device_batch(data_set, check_validation)
            self.global_batch_count += 1
            if self.hparams.dynamic_bias_eta_change != 0 and not self.hparams.oracle_mode:
                if self.global_batch_count % self.hparams.dynamic_bias_step_interval == 0:
                    self.click_model.eta += self.hparams.dynamic_bias_eta_change
                    self.click_model.setExamProb(self.click_model.eta)
                    print('Dynamically change bias severity eta to %.3f' % self.click_model.eta)
            return feed, _LazyInfo(feed)
        init, labels, _ = self._arrays(data_set)
        length = init.shape[0]
        B = self.batch_size
        sel_idx, sel_clicks, have = [], [], 0
        rounds = 0
        while have < B:
            n_try = max(8, int((B - have) * 1.5) + 4)
            cand = (self.rng.random(n_try) * length).astype(np.int64)
            clicks = self._simulate_queries(cand)
            keep = clicks.sum(axis=1) > 0 if check_validation else np.ones(n_try, dtype=bool)
            cand, clicks = cand[keep][:B - have], clicks[keep][:B - have]
            sel_idx.append(cand)
            sel_clicks.append(clicks)
            have += len(cand)
            rounds += 1
            if rounds > 10000:
                raise RuntimeError("could not sample %d lists with clicks" % B)
        idx = np.concatenate(sel_idx)
        clicks = np.concatenate(sel_clicks, axis=0)
        input_feed, docid, _ = self._assemble(idx, clicks)
        info_map = {
            'rank_list_idxs': idx.tolist(),
            'input_list': docid,
            'click_list': clicks,
            'letor_features': input_feed[self.model.letor_features_name],
        }
        self.global_batch_count += 1
        if self.hparams.dynamic_bias_eta_change != 0 and not self.hparams.oracle_mode:
            if self.global_batch_count % self.hparams.dynamic_bias_step_interval == 0:
                self.click_model.eta += self.hparams.dynamic_bias_eta_change
                self.click_model.setExamProb(self.click_model.eta)
                print('Dynamically change bias severity eta to %.3f' % self.click_model.eta)
        return input_feed, info_map

    def _sequential(self, data_set, indices, check_validation):
        self._check_list_size(data_set)
        _, labels, _ = self._arrays(data_set)
        idx = np.asarray(indices, dtype=np.int64)
        clicks = self._simulate_queries(idx)
        if check_validation:
            keep = clicks.sum(axis=1) > 0
            idx, clicks = idx[keep], clicks[keep]
        input_feed, docid, _ = self._assemble(idx, clicks)
        return input_feed, {'input_list': docid, 'click_list': clicks}

    def get_next_batch(self, index, data_set, check_validation=False, data_format="ULTRA"):
        """Sequential batch starting at `index` (click_simulation_feed.py:176-241)."""
        n = min(self.batch_size, len(data_set.initial_list) - index)
        return self._sequential(data_set, range(index, index + n), check_validation)

    def get_data_by_index(self, data_set, index, check_validation=False):
        """click_simulation_feed.py:243-294."""
        return self._sequential(data_set, [index], check_validation)
