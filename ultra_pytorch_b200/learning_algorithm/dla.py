"""B200-native drop-in for `ultra.learning_algorithm.DLA` (reference: ultra/learning_algorithm/dla.py:21-330).

One kernel evaluates both listwise losses of the dual learning algorithm (ranker loss weighted by the propensity
net, examination loss weighted by the ranker), the DenoisingNet forward/backward and both weight normalisations."""
import torch
import torch.nn as nn

from .base_algorithm import B200Algorithm, HParams


class DenoisingNet(nn.Module):
    """dla.py:24-48.  ELU(Linear(one_hot(position))) == ELU(W[0, position] + b): the kernel evaluates that closed
    form; this module only owns the parameters (views of the flat buffer [W(L) | b])."""

    def __init__(self, input_vec_size, device):
        super(DenoisingNet, self).__init__()
        self.linear_layer = nn.Linear(input_vec_size, 1)
        self.elu_layer = nn.ELU()
        self.propensity_net = nn.Sequential(self.linear_layer, self.elu_layer)
        self.list_size = input_vec_size
        L = input_vec_size
        self.flat = torch.zeros(L + 1, dtype=torch.float32, device=device)
        with torch.no_grad():
            self.flat[:L].copy_(self.linear_layer.weight.detach().view(-1))
            self.flat[L:].copy_(self.linear_layer.bias.detach().view(-1))
        self.linear_layer.weight.data = self.flat[:L].view(1, L)
        self.linear_layer.bias.data = self.flat[L:].view(1)

    def forward(self, input_list):
        """Propensity logits [B, L] for L inputs of shape [B] (value-independent, dla.py:32-46)."""
        B = input_list[0].shape[0]
        L = self.list_size
        prop = torch.nn.functional.elu(self.flat[:L] + self.flat[L])
        return prop.view(1, L).expand(B, L)


class DLA(B200Algorithm):
    def __init__(self, data_set, exp_settings):
        print('Build DLA')
        self.hparams = HParams(
            learning_rate=0.05,                 # dla.py:72
            max_gradient_norm=5.0,
            loss_func='softmax_loss',
            logits_to_prob='softmax',
            propensity_learning_rate=-1.0,
            ranker_loss_weight=1.0,
            l2_loss=0.0,
            max_propensity_weight=-1,
            constant_propensity_initialization=False,
            grad_strategy='ada',
        )
        print(exp_settings['learning_algorithm_hparams'])
        self.hparams.parse(exp_settings['learning_algorithm_hparams'])
        if self.hparams.loss_func in ('sigmoid_loss', 'pairwise_loss') or self.hparams.logits_to_prob != 'softmax':
            raise NotImplementedError("DLA on B200 implements loss_func='softmax_loss', logits_to_prob='softmax' "
                                      "(the reference defaults)")
        L = exp_settings['selection_bias_cutoff']
        self._init_common(data_set, exp_settings, extra_floats=4 + L + 1)
        self._check_l2()
        # same construction order as the reference (propensity net first, dla.py:102-104) -> same initial weights
        # for the same torch seed
        self.propensity_model = DenoisingNet(self.rank_list_size, torch.device('cuda', torch.cuda.current_device()))
        self.model = self.create_model(self.feature_size)
        self.broadcast_initial_state(self.propensity_model.flat)
        if self.hparams.propensity_learning_rate < 0:
            self.propensity_learning_rate = float(self.hparams.learning_rate)
        else:
            self.propensity_learning_rate = float(self.hparams.propensity_learning_rate)
        self.learning_rate = float(self.hparams.learning_rate)
        eng = self.engine
        self._sums = eng.extra[:4]
        self._dprop = eng.extra[4:4 + L + 1]
        self.propensity_model.linear_layer.weight.grad = self._dprop[:L].view(1, L)
        self.propensity_model.linear_layer.bias.grad = self._dprop[L:].view(1)
        self._norms = torch.zeros(2, dtype=torch.float32, device=eng.device)
        self._scal = torch.zeros(6, dtype=torch.float32, device=eng.device)
        self._norm = None
        self._norm_pending = False

    L2_EXHAUSTS_CLIP_PARAMS = False     # dla.py:161-163 clips fresh parameter iterators

    def device_step(self, st):
        eng = self.engine
        L, B = st.L, st.B
        flat = self.propensity_model.flat
        if self._phase != "post":
            docid = st.docid.view(-1)
            scores = eng.forward(st.feats, docid, L, B, training=True)
            dscores = eng.dscores_buf(B, L)
            eng.dla_loss(scores, st.labels, flat[:L], flat[L:], dscores, self._dprop, self._sums)
            self._publish_early(self._sums)                 # both losses are final here on a single GPU
            eng.backward(st.feats, docid, L, B, dscores)
        if self._phase == "pre":
            return None
        # fresh optimizers every step (dla.py:153-154): accumulator starts from zero -> mode 1; the two parameter
        # groups are clipped separately (dla.py:161-163).  The ranker's update goes first: in data-parallel mode it
        # carries the exchange of the whole flat buffer, DenoisingNet gradients and normalisers included.
        mode = self._opt_mode(fresh=True)
        mg = self.hparams.max_gradient_norm
        self._exchange_and_update(None, self._sums[1:2], float(self.hparams.ranker_loss_weight), self.learning_rate,
                                  mode, self._norms[1:2])
        eng.clip_update(flat, self._dprop, None, self._sums[3:4], 1.0, mg, self.propensity_learning_rate, mode,
                        self._norms[0:1])
        self._scal[:4].copy_(self._sums)
        self._scal[4:6].copy_(self._norms)
        eng.join_publish()
        return self._scal

    def _post_clip_norm(self, norms):
        mg = self.hparams.max_gradient_norm
        post = [float(n) * min(1.0, mg / (float(n) + 1e-6)) if mg > 0 else float(n) for n in norms]
        return (post[0] ** 2 + post[1] ** 2) ** 0.5               # dla.py:166-177

    @property
    def norm(self):
        """Norm of the clipped gradients of the last step (dla.py:166-177; the reference spends 18 host syncs per step
        on it and never reads it).  With the early loss read-back it is fetched on first use."""
        if self._norm_pending:
            self._norm = self._post_clip_norm(self._norms.cpu().numpy())
            self._norm_pending = False
        return self._norm

    def train(self, input_feed):
        """dla.py:179-266 + separate_gradient_update dla.py:141-177."""
        self.rank_list_size = self.exp_settings['selection_bias_cutoff']
        if not self.model.training:
            self.model.train()
        st = self._stage(input_feed, self.rank_list_size)
        s = self._read_scalars(self.run_step(st))
        self.rank_loss = float(s[0] / s[1]) + self._l2_loss_value()
        self.exam_loss = float(s[2] / s[3])
        self.loss = self.exam_loss + self.hparams.ranker_loss_weight * self.rank_loss
        if len(s) >= 6:
            self._norm = self._post_clip_norm(s[4:6])
            self._norm_pending = False
        else:                       # early read-back: the gradient norms arrive with the end of the step -> lazily
            self._norm_pending = True
        self._say(self.loss)
        self.global_step += 1
        return self.loss, None, self.train_summary
