"""Array-form drop-in for `ultra.input_layer.DirectLabelFeed` (reference: ultra/input_layer/direct_label_feed.py:18-284),
the feed of every validation / test sweep (main.py:128-136, 170-187, 252-266) and of config 1's training.

Same `input_feed` as the reference, built with array operations instead of Python loops; `get_batch` draws its queries
from Python's `random` in the reference's order, so with the same `random.seed` ALL three entry points are
bit-identical to the reference (tests/test_click_feed.py).  `resident_features=True` (input_layer/resident.py) emits
global doc ids + the data set's whole feature matrix, so a validation batch moves only ids and labels to the device.
ULTRE-format data sets (`features_dict`) are not supported."""
import random

import numpy as np

from ..hparams import HParams
from .click_simulation_feed import ClickSimulationFeed


class DirectLabelFeed(ClickSimulationFeed):
    def __init__(self, model, batch_size, hparam_str):
        self.hparams = HParams(
            use_max_candidate_num=True,        # direct_label_feed.py:36-40
            resident_features=False,           # B200 extension, see input_layer/resident.py
        )
        self.hparams.parse(hparam_str)
        self.hparams.oracle_mode = True        # the "clicks" of this feed are the true labels
        self.click_model = None
        self.start_index = 0
        self.count = 1
        self.rank_list_size = model.max_candidate_num if self.hparams.use_max_candidate_num else model.rank_list_size
        self.feature_size = model.feature_size
        self.batch_size = batch_size
        self.model = model
        self.global_batch_count = 0
        self.rng = None                        # nothing is sampled with numpy here
        self._cache_key = None
        print('Create direct label feed with list size %d with feature size %d' % (self.rank_list_size,
                                                                                  self.feature_size))

    @staticmethod
    def _only_ultra(data_format):
        if data_format == "ULTRE":
            raise NotImplementedError("the ULTRE data format (features_dict) is not supported by the B200 feeds")

    def get_batch(self, data_set, check_validation=False, data_format="ULTRA"):
        """direct_label_feed.py:94-158: batch_size queries drawn with replacement; lists without a relevant document
        are dropped (NOT replaced) when check_validation is set."""
        self._only_ultra(data_format)
        self._check_list_size(data_set)
        _, labels, _ = self._arrays(data_set)
        length = len(data_set.initial_list)
        drawn = [int(random.random() * length) for _ in range(self.batch_size)]
        idx = np.asarray(drawn, dtype=np.int64)
        if check_validation:
            idx = idx[labels[idx].sum(axis=1) != 0]
        input_feed, docid, _ = self._assemble(idx, labels[idx])
        info_map = {
            'rank_list_idxs': drawn,
            'input_list': docid,
            'click_list': labels[idx],
            'letor_features': input_feed[self.model.letor_features_name],
        }
        return input_feed, info_map

    def get_next_batch(self, index, data_set, check_validation=False, data_format="ULTRA"):
        """direct_label_feed.py:160-226."""
        self._only_ultra(data_format)
        return ClickSimulationFeed.get_next_batch(self, index, data_set, check_validation, data_format)
