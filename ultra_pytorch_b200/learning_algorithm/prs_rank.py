"""B200-native drop-in for `ultra.learning_algorithm.PRSrank` (reference: ultra/learning_algorithm/prs_rank.py:22-251):
a LambdaRank variant whose pair (r ranked above s) is weighted by delta-NDCG * ipw_r / ipw_s, with ipw taken from the
randomised-propensity table at the documents' DISPLAY positions (`getPropensityForOneList(use_non_clicked_data=True)`,
propensity_estimator.py:22-42), and whose loss is the weighted binary cross-entropy of sigmoid(sigma (s_r - s_s)).
Same pair kernel as LambdaRank / PairDebias (csrc/losses.cu: pairwise_kernel<2>): every unordered pair once, no [B, L, L]
temporaries (the reference materialises nine of them, prs_rank.py:126-146)."""
import torch

from .base_algorithm import B200Algorithm, HParams
from .ipw_rank import IPWrank


class PRSrank(B200Algorithm):
    def __init__(self, data_set, exp_settings):
        self.hparams = HParams(
            propensity_estimator_type='ultra.utils.propensity_estimator.RandomizedPropensityEstimator',
            propensity_estimator_json='./example/PropensityEstimator/randomized_pbm_0.1_1.0_4_1.0.json',
            learning_rate=0.05,                 # prs_rank.py:47
            max_gradient_norm=5.0,
            grad_strategy='ada',
            sigma=1.0,
        )
        print(exp_settings['learning_algorithm_hparams'])
        self.hparams.parse(exp_settings['learning_algorithm_hparams'])
        L = exp_settings['selection_bias_cutoff']
        self._init_common(data_set, exp_settings, extra_floats=2 * L + 2)
        self.model = self.create_model(self.feature_size)
        self.propensity_estimator = IPWrank._load_estimator(self)
        self._table = torch.as_tensor(self.propensity_estimator.IPW_list, dtype=torch.float32,
                                      device=self.engine.device).contiguous()
        self.learning_rate = float(self.hparams.learning_rate)
        self.sigma = self.hparams.sigma
        self._scal = torch.zeros(2, dtype=torch.float32, device=self.engine.device)

    def device_step(self, st):
        eng = self.engine
        L, B = st.L, st.B
        out = eng.extra[:2 * L + 2]
        if self._phase != "post":
            docid = st.docid.view(-1)
            scores = eng.forward(st.feats, docid, L, B, training=True)
            dscores = eng.dscores_buf(B, L)
            eng.prsrank(scores, st.labels, self.sigma, self._table, dscores, out)
            self._publish_early(out[2 * L:2 * L + 2])
            eng.backward(st.feats, docid, L, B, dscores)
        if self._phase == "pre":
            return None
        # gains are normalised by ONE batch-global IDCG (prs_rank.py:214-218, 228-231): applied here as 1 / idcg
        self._exchange_and_update(eng.state_sum, out[2 * L + 1:2 * L + 2], 1.0, self.learning_rate, self._opt_mode(),
                                  eng.norm)
        self._scal.copy_(out[2 * L:2 * L + 2])
        eng.join_publish()
        return self._scal

    def train(self, input_feed):
        """prs_rank.py:94-176."""
        if not self.model.training:
            self.model.train()
        st = self._stage(input_feed, self.rank_list_size)
        s = self._read_scalars(self.run_step(st))
        self.loss = float(s[0] / s[1])
        self._say(self.loss)
        self.global_step += 1
        return self.loss, None, self.train_summary
