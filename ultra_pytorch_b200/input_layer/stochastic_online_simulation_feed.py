"""Drop-in for `ultra.input_layer.StochasticOnlineSimulationFeed`
(reference: ultra/input_layer/stochastic_online_simulation_feed.py:24-381).

The reference builds the batch with Python loops, scores it with `model.validation(input_feed, True)`, copies the
scores to the host and then, PER QUERY, draws a Plackett-Luce permutation with `np.random.choice`, re-orders the
documents and simulates clicks with the click model (:100-177).  Here the batch is assembled with array operations
(shared with the ClickSimulationFeed drop-in), the Plackett-Luce permutations of the whole batch are drawn on the GPU
from the score tensor the ranker just produced (csrc/sampling.cu: Gumbel-top-k with Philox noise, one CTA per list),
only the [B, L] permutation comes back to the host, and re-ordering + click simulation are array operations.

Same `input_feed` format, hparams (`click_model_json`, `oracle_mode`, `dynamic_bias_eta_change`,
`dynamic_bias_step_interval`, `tau`) and semantics as the reference; the random streams are equal in distribution,
not draw for draw (the reference uses the global `random` / `np.random` state).  Result interleaving
(`need_interleave`, DBGD family) is not part of this path and raises.
"""
import json
import random

import numpy as np

from ..hparams import HParams
from .click_simulation_feed import ClickSimulationFeed, load_click_model


class StochasticOnlineSimulationFeed(ClickSimulationFeed):
    def __init__(self, model, batch_size, hparam_str):
        self.hparams = HParams(
            click_model_json='./example/ClickModel/pbm_0.1_1.0_4_1.0.json',    # stochastic_online_simulation_feed.py:40-53
            oracle_mode=False,
            dynamic_bias_eta_change=0.0,
            dynamic_bias_step_interval=1000,
            tau=1.0,
            resident_features=False,       # B200 extension, see input_layer/resident.py
        )
        print('Create online stochastic simluation feed')
        print(hparam_str)
        self.hparams.parse(hparam_str)
        with open(self.hparams.click_model_json) as fin:
            self.click_model = load_click_model(json.load(fin))
        self.start_index = 0
        self.count = 1
        self.rank_list_size = model.rank_list_size
        self.max_candidate_num = model.max_candidate_num
        self.feature_size = model.feature_size
        self.batch_size = batch_size
        self.model = model
        self.global_batch_count = 0
        if getattr(getattr(model, "hparams", None), "need_interleave", False):
            raise NotImplementedError("result interleaving (need_interleave) is outside the B200 hot path")
        self.rng = np.random.default_rng(random.getrandbits(63))
        self._seed = random.getrandbits(63)
        self._draws = 0
        self._cache_key = None

    # ---- array view of the data set over max_candidate_num positions ------------------------------------------
    def _arrays(self, data_set):
        key = (id(data_set), len(data_set.initial_list), len(data_set.features), self.max_candidate_num)
        if self._cache_key != key:
            saved = self.rank_list_size
            self.rank_list_size = self.max_candidate_num            # the parent builds [nq, rank_list_size] arrays
            try:
                self._cache_key = None
                ClickSimulationFeed._arrays(self, data_set)
            finally:
                self.rank_list_size = saved
            self._cache_key = key
        return self._init, self._labels, self._features

    def _assemble_true(self, idx):
        """input_feed over max_candidate_num positions with the TRUE labels (prepare_true_labels_with_index, :80-98)."""
        saved = self.rank_list_size
        self.rank_list_size = self.max_candidate_num
        try:
            input_feed, docid, n_docs = self._assemble(idx, self._labels[idx])
        finally:
            self.rank_list_size = saved
        return input_feed, docid, n_docs

    # ---- the GPU part ------------------------------------------------------------------------------------
    def _sample_permutations(self, input_feed):
        """Scores the batch with the ranker and draws one Plackett-Luce permutation per list ON THE DEVICE.
        Returns perm [B, max_candidate_num] (host int64): perm[b][r] = position of the document shown at rank r."""
        scores = self.model.validation(input_feed, True)[1]           # [B, max_cand] cuda tensor (:113)
        st = getattr(self.model, "_last_validation_stage", None)
        if st is None or not hasattr(self.model, "engine"):
            raise TypeError("ultra_pytorch_b200.input_layer.StochasticOnlineSimulationFeed drives a B200 learning "
                            "algorithm (the Plackett-Luce sampling kernel reads the ranker's device buffers)")
        self._draws += 1
        perm = self.model.engine.pl_sample(scores, st.docid, st.n_docs, self.hparams.tau, self._seed, self._draws)
        return perm.cpu().numpy().astype(np.int64)

    def _simulate_prefix(self, labels, check_validation, valid):
        """Clicks on the first rank_list_size positions of the re-ranked lists (:150-161): PBM sample, re-drawn (up to
        MAX_SAMPLE_ROUND_NUM times) for lists without any click when check_validation is set.  `valid` masks the real
        positions: the reference samples only the list_len real documents (:150 truncates to the list), so a click drawn
        on a PAD slot (label 0 still has the noise-click probability) must neither survive nor stop the re-draw."""
        if self.hparams.oracle_mode:
            return labels * valid
        if self.click_model.position_independent:
            p = self.click_model.click_probability(labels) * valid
            draw = lambda rows: (self.rng.random((len(rows), labels.shape[1])) < p[rows]).astype(np.float64)
        else:
            # user-browsing / cascade model: sampled position by position; the real documents are a prefix of every
            # list, so a (discarded) click on a PAD slot cannot influence a real position behind it
            draw = lambda rows: self.click_model.sample(labels[rows], self.rng) * valid[rows]
        clicks = draw(np.arange(labels.shape[0]))
        if check_validation:
            for _ in range(self.MAX_SAMPLE_ROUND_NUM):
                redo = np.flatnonzero(clicks.sum(axis=1) == 0)
                if redo.size == 0:
                    break
                clicks[redo] = draw(redo)
        return clicks

    def simulate_clicks_online(self, input_feed, check_validation=False):
        """stochastic_online_simulation_feed.py:100-177 for the whole batch at once."""
        L, K = self.max_candidate_num, self.rank_list_size
        perm = self._sample_permutations(input_feed)                                          # [B, L]
        docid = np.stack([input_feed[self.model.docid_inputs_name[l]] for l in range(L)], axis=1)   # [B, L] f32
        label = np.stack([input_feed[self.model.labels_name[l]] for l in range(L)], axis=1)
        new_docid = np.take_along_axis(docid, perm, axis=1)            # positions behind list_len map to themselves
        new_label = np.take_along_axis(label, perm, axis=1)
        n_docs = len(input_feed[self.model.letor_features_name])
        list_len = (docid < n_docs).cumsum(axis=1).argmax(axis=1) + 1
        list_len = np.where((docid < n_docs).any(axis=1), list_len, 0)
        k = min(K, L)
        clicks = np.zeros_like(new_label)
        valid = (np.arange(L)[None, :] < list_len[:, None])
        clicks[:, :k] = self._simulate_prefix(new_label[:, :k].astype(np.float64), check_validation,
                                              valid[:, :k].astype(np.float64))
        clicks[~valid] = 0.0                                            # only real positions receive labels (:171-176)
        for l in range(L):
            input_feed[self.model.docid_inputs_name[l]] = np.ascontiguousarray(new_docid[:, l], dtype=np.float32)
            input_feed[self.model.labels_name[l]] = np.ascontiguousarray(clicks[:, l], dtype=np.float32)
        return input_feed

    # ---- reference API ----------------------------------------------------------------------------------------
    def _batch(self, data_set, idx, check_validation):
        self._check_list_size(data_set)
        _, labels, _ = self._arrays(data_set)
        idx = np.asarray(idx, dtype=np.int64)
        if check_validation:                                           # lists without relevant documents are skipped (:88-90)
            idx = idx[labels[idx].sum(axis=1) != 0]
        input_feed, docid, _ = self._assemble_true(idx)
        true_labels = labels[idx]
        input_feed = self.simulate_clicks_online(input_feed, check_validation)
        return input_feed, idx, docid, true_labels

    def get_batch(self, data_set, check_validation=False, data_format="ULTRA"):
        """stochastic_online_simulation_feed.py:179-250: batch_size random queries (with replacement; queries without
        relevant documents are dropped, not replaced, when check_validation is set)."""
        length = len(data_set.initial_list)
        cand = (self.rng.random(self.batch_size) * length).astype(np.int64)
        input_feed, idx, docid, true_labels = self._batch(data_set, cand, check_validation)
        info_map = {
            'rank_list_idxs': cand.tolist(),
            'input_list': docid,
            'click_list': true_labels,
            'letor_features': input_feed[self.model.letor_features_name],
        }
        self.global_batch_count += 1
        if self.hparams.dynamic_bias_eta_change != 0:
            if self.global_batch_count % self.hparams.dynamic_bias_step_interval == 0:
                self.click_model.eta += self.hparams.dynamic_bias_eta_change
                self.click_model.setExamProb(self.click_model.eta)
                print('Dynamically change bias severity eta to %.3f' % self.click_model.eta)
        return input_feed, info_map

    def get_next_batch(self, index, data_set, check_validation=False, data_format="ULTRA"):
        """:252-320 (the reference version raises AttributeError at `self.model.letor_features.name`)."""
        n = min(self.batch_size, len(data_set.initial_list) - index)
        input_feed, _, docid, true_labels = self._batch(data_set, range(index, index + n), check_validation)
        return input_feed, {'input_list': docid, 'click_list': true_labels}

    def get_data_by_index(self, data_set, index, check_validation=False):
        """:322-381."""
        input_feed, _, docid, true_labels = self._batch(data_set, [index], check_validation)
        return input_feed, {'input_list': docid, 'click_list': true_labels}
