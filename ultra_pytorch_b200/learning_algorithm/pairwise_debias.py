"""B200-native drop-in for `ultra.learning_algorithm.PairDebias`
(reference: ultra/learning_algorithm/pairwise_debias.py:34-203).  The reference's L*(L-1) Python loop of ~12 torch ops
(21.7 s/step at L=200 on CPU) is one CTA-per-list kernel."""
import torch

from .base_algorithm import B200Algorithm, HParams


class PairDebias(B200Algorithm):
    SAFE_DIV = 0

    def __init__(self, data_set, exp_settings):
        print('Build Pairwise Debiasing algorithm.')
        self.hparams = HParams(
            EM_step_size=0.05,                  # pairwise_debias.py:54
            learning_rate=0.005,
            max_gradient_norm=5.0,
            regulation_p=1,
            l2_loss=0.0,
            grad_strategy='ada',
        )
        print(exp_settings['learning_algorithm_hparams'])
        self.hparams.parse(exp_settings['learning_algorithm_hparams'])
        self._setup(data_set, exp_settings)

    def _setup(self, data_set, exp_settings):
        L = exp_settings['selection_bias_cutoff']
        self._init_common(data_set, exp_settings, extra_floats=2 * L + 2)
        self._check_l2()
        self.model = self.create_model(self.feature_size)
        self.learning_rate = float(self.hparams.learning_rate)
        dev = self.engine.device
        self.t_plus = torch.ones([1, self.rank_list_size], device=dev)     # pairwise_debias.py:93-97
        self.t_minus = torch.ones([1, self.rank_list_size], device=dev)
        self._scal = torch.zeros(2, dtype=torch.float32, device=dev)
        self._b_global = 1.0

    def _pair_kernel(self, scores, labels, dscores, out):
        self.engine.pairdebias(scores, labels, self.t_plus, self.t_minus, dscores, out)

    def device_step(self, st):
        eng = self.engine
        L, B = st.L, st.B
        out = eng.extra[:2 * L + 2]
        if self._phase != "post":
            docid = st.docid.view(-1)
            scores = eng.forward(st.feats, docid, L, B, training=True)
            dscores = eng.dscores_buf(B, L)
            self._pair_kernel(scores, st.labels, dscores, out)
            self._publish_early(out[2 * L:2 * L + 2])         # loss (+ idcg) are final here on a single GPU
            eng.backward(st.feats, docid, L, B, dscores)
        if self._phase == "pre":
            return None
        self._update(out, L, B)                          # exchange (data parallel) + clip + optimizer
        self._scal.copy_(out[2 * L:2 * L + 2])          # loss (+ idcg) before anything reuses the buffer
        eng.em_update(self.t_plus, self.t_minus, out, self.hparams.EM_step_size, self.hparams.regulation_p,
                      self.SAFE_DIV)
        eng.join_publish()
        return self._scal

    def _update(self, out, L, B):
        # the reference's loss carries a x batch_size factor ([B]*[B,1] broadcast, base_algorithm.py:246-247)
        eng = self.engine
        self._b_global = float(B * self.world_size())
        self._exchange_and_update(eng.state_sum, None, self._b_global, self.learning_rate, self._opt_mode(), eng.norm)

    def train(self, input_feed):
        """pairwise_debias.py:106-174."""
        if not self.model.training:
            self.model.train()
        st = self._stage(input_feed, self.rank_list_size)
        s = self._read_scalars(self.run_step(st))
        self.loss = float(s[0]) * self._b_global + self._l2_loss_value()
        if self.VERBOSE:
            print(" Loss %f at Global Step %d" % (self.loss, self.global_step))
        self.global_step += 1
        return self.loss, None, self.train_summary
