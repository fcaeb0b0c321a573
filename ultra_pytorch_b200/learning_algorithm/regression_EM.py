"""B200-native drop-in for `ultra.learning_algorithm.RegressionEM`
(reference: ultra/learning_algorithm/regression_EM.py:37-190): the regression-based EM algorithm - examination
propensities per display position estimated online (E-step posteriors from the current ranker scores, M-step moving
average), the ranker trained with a pointwise sigmoid cross-entropy on Bernoulli pseudo-labels.

One step = K1 forward -> ONE kernel for E-step + sampling + loss + gradient + M-step sums (csrc/sampling.cu:
regression_em_kernel) -> K1 backward -> exchange (data parallel) + clip + Adagrad -> M-step update.  The reference's
`sigmoid_prob_b` is a constant zero that is never trained (regression_EM.py:99-101, not a parameter), so it does not
appear here.  The pseudo-labels are random: parity with the reference is exact given the same uniform draws (tested by
replaying the reference's draws) and distributional otherwise (Philox instead of torch.rand).
"""
import torch

from .base_algorithm import B200Algorithm, HParams


class RegressionEM(B200Algorithm):
    # the draw counter of the pseudo-label sampler is a kernel ARGUMENT: a captured CUDA graph would replay the same
    # pseudo-labels every step, so this algorithm launches its kernels eagerly
    USE_GRAPH = False

    def __init__(self, data_set, exp_settings):
        print('Build Regression-based EM algorithm.')
        self.hparams = HParams(
            EM_step_size=0.05,                  # regression_EM.py:62-69
            learning_rate=0.05,
            max_gradient_norm=5.0,
            l2_loss=0.0,
            grad_strategy='ada',
        )
        print(exp_settings['learning_algorithm_hparams'])
        self.hparams.parse(exp_settings['learning_algorithm_hparams'])
        L = exp_settings['selection_bias_cutoff']
        self._init_common(data_set, exp_settings, extra_floats=2 + L)
        self._check_l2()
        self.model = self.create_model(self.feature_size)
        self.learning_rate = float(self.hparams.learning_rate)
        dev = self.engine.device
        self.propensity = torch.ones(1, L, dtype=torch.float32, device=dev) * 0.9      # regression_EM.py:94-97
        self.sigmoid_prob_b = torch.zeros(1, dtype=torch.float32, device=dev)
        self._seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        self._draws = 0
        self.replay_uniforms = None          # tests: a [B, L] cuda tensor of uniform draws to use for the next step

    @property
    def propensity_weights(self):
        return 1.0 / self.propensity          # regression_EM.py:185

    def device_step(self, st):
        eng = self.engine
        L, B = st.L, st.B
        out = eng.extra[:2 + L]
        if self._phase != "post":
            docid = st.docid.view(-1)
            scores = eng.forward(st.feats, docid, L, B, training=True)
            dscores = eng.dscores_buf(B, L)
            self._draws += 1
            eng.regression_em(scores, st.labels, self.propensity.view(-1), self.replay_uniforms, self._seed,
                              self._draws, dscores, out)
            self._publish_early(out[:2])
            eng.backward(st.feats, docid, L, B, dscores)
        if self._phase == "pre":
            return None
        # mean over the B*L elements of the (global) batch: the normaliser out[1] is summed by the exchange
        self._exchange_and_update(eng.state_sum, out[1:2], 1.0, self.learning_rate, self._opt_mode(), eng.norm)
        eng.regem_update(self.propensity.view(-1), out, float(self.hparams.EM_step_size))
        eng.join_publish()
        return out

    def train(self, input_feed):
        """regression_EM.py:108-190."""
        if not self.model.training:
            self.model.train()
        st = self._stage(input_feed, self.rank_list_size)
        s = self._read_scalars(self.run_step(st))
        self.loss = float(s[0] / s[1]) + self._l2_loss_value()
        self.update_propensity_op = self.propensity
        self.global_step += 1
        if self.VERBOSE:
            print('Loss %f at global step %d' % (self.loss, self.global_step))
        return self.loss, None, self.train_summary
