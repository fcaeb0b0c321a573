"""B200-native drop-in for `ultra.learning_algorithm.NavieAlgorithm`
(reference: ultra/learning_algorithm/navie_algorithm.py:24-149): listwise softmax loss on the raw labels."""
from .base_algorithm import B200Algorithm, HParams


class NavieAlgorithm(B200Algorithm):
    WEIGHT_MODE = 0

    def __init__(self, data_set, exp_settings):
        print('Build NavieAlgorithm')
        self.hparams = HParams(
            learning_rate=0.05,                 # navie_algorithm.py:33
            max_gradient_norm=5.0,
            loss_func='softmax_cross_entropy',
            l2_loss=0.0,
            grad_strategy='ada',
        )
        self.hparams.parse(exp_settings['learning_algorithm_hparams'])
        self._init_common(data_set, exp_settings, extra_floats=2)
        self._check_loss()
        self._check_l2()
        self.model = self.create_model(self.feature_size)
        self.learning_rate = float(self.hparams.learning_rate)
        self._table = None

    def _check_loss(self):
        if self.hparams.loss_func in ('sigmoid_loss', 'pairwise_loss'):
            # both raise inside the reference as well (base_algorithm.py:267, 306-307)
            raise NotImplementedError("loss_func=%s does not run in the reference either; only the softmax loss "
                                      "is implemented" % self.hparams.loss_func)

    def device_step(self, st):
        """All device work of one training step on an already-staged batch (no host sync).
        Returns the device tensor holding the loss scalars."""
        eng = self.engine
        L, B = st.L, st.B
        sums = eng.extra[:2]
        if self._phase != "post":
            docid = st.docid.view(-1)
            scores = eng.forward(st.feats, docid, L, B, training=True)
            dscores = eng.dscores_buf(B, L)
            eng.softmax_ce(scores, st.labels, self.WEIGHT_MODE, self._table, dscores, sums)
            self._publish_early(sums)
            eng.backward(st.feats, docid, L, B, dscores)
        if self._phase == "pre":
            return None
        self._exchange_and_update(eng.state_sum, sums[1:2], 1.0, self.learning_rate, self._opt_mode(), eng.norm)
        eng.join_publish()
        return sums

    def train(self, input_feed):
        """navie_algorithm.py:76-120 / ipw_rank.py:102-182."""
        self.global_step += 1
        if not self.model.training:
            self.model.train()
        st = self._stage(input_feed, self.rank_list_size)
        s = self._read_scalars(self.run_step(st))
        self.loss = float(s[0] / s[1]) + self._l2_loss_value()
        self._say(self.loss)
        return self.loss, None, self.train_summary
