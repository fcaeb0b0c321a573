"""B200-native drop-in for `ultra.learning_algorithm.LambdaRank`
(reference: ultra/learning_algorithm/lambda_rank.py:22-291): pairwise-debiased LambdaRank with delta-NDCG weights.
The ~12 materialised [B, L, L] temporaries of the reference are never formed: one CTA per list evaluates the pair
tile from shared memory."""
from .base_algorithm import HParams
from .pairwise_debias import PairDebias


class LambdaRank(PairDebias):
    SAFE_DIV = 1

    def __init__(self, data_set, exp_settings):
        self.hparams = HParams(
            EM_step_size=0.05,                  # lambda_rank.py:43
            learning_rate=0.05,
            max_gradient_norm=5.0,
            grad_strategy='ada',
            regulation_p=1,
            sigma=1.0,
        )
        print(exp_settings['learning_algorithm_hparams'])
        self.hparams.parse(exp_settings['learning_algorithm_hparams'])
        self._setup(data_set, exp_settings)
        self.sigma = self.hparams.sigma

    def _pair_kernel(self, scores, labels, dscores, out):
        self.engine.lambdarank(scores, labels, self.sigma, self.t_plus, self.t_minus, dscores, out)

    def _update(self, out, L, B):
        # gains are normalised by ONE batch-global IDCG (lambda_rank.py:263-266, 277): applied here as 1/idcg
        eng = self.engine
        self._exchange_and_update(eng.state_sum, out[2 * L + 1:2 * L + 2], 1.0, self.learning_rate, self._opt_mode(),
                                  eng.norm)

    def train(self, input_feed):
        """lambda_rank.py:96-216."""
        self.rank_list_size = self.exp_settings['selection_bias_cutoff']
        self.global_step += 1
        if not self.model.training:
            self.model.train()
        st = self._stage(input_feed, self.rank_list_size)
        s = self._read_scalars(self.run_step(st))
        self.loss = float(s[0] / s[1])
        self._say(self.loss)
        return self.loss, None, self.train_summary
