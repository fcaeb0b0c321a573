"""B200-native drop-in for `ultra.ranking_model.DNN` (reference: ultra/ranking_model/DNN.py:11-88).

Same constructor `(hparams_str, feature_size)`, same hparams (`hidden_layer_sizes`, `activation_func`, `norm`),
same `build(input_list, noisy_params=None, noise_rate=0.05, **kwargs)` contract and the same `state_dict()` keys /
shapes (`sequential.layer_norm{j}.{weight,bias}`, `sequential.linear{j}.{weight,bias}`), so checkpoints interchange
with the reference (main.py:75-79, 206).  The module is an ordinary `nn.Module`; its parameters are VIEWS into one
flat CUDA buffer (layout of include/ultra_b200.h) that the sm_100a kernels read and update in place.
"""
import torch
import torch.nn as nn

from ..engine import RankerEngine
from ..hparams import HParams


# hparam activation_func -> (UB200_ACT_* code of include/ultra_b200.h, module class; the modules only keep the reference's
# Sequential layout - they hold no parameters and the plugin never calls them)
# 'selu' is not offered: the reference's selu is a plain function that nn.Sequential.add_module rejects with a TypeError
# (DNN.py:52-54), so no reference run with it exists (the kernels and the oracle implement it, code 2, for completeness).
ACTIVATIONS = {'elu': (0, nn.ELU), 'relu': (1, nn.ReLU), 'tanh': (3, nn.Tanh), 'sigmoid': (4, nn.Sigmoid)}


class DNN(nn.Module):
    def __init__(self, hparams_str, feature_size, extra_floats=0):
        super(DNN, self).__init__()
        self.hparams = HParams(
            hidden_layer_sizes=[512, 256, 128],   # DNN.py:27
            activation_func='elu',                # DNN.py:30
            norm="layer",                         # DNN.py:31
        )
        self.hparams.parse(hparams_str)
        # base_ranking_model.py:63-69: elu (default; tensor-core kernels), relu / selu / tanh / sigmoid (fp32 CUDA-core
        # kernels).  norm: 'layer' only - 'batch' builds BatchNorm2d over 2-D activations in the reference, which raises
        # in its forward pass, and any other string silently drops the normalisation layers.
        if self.hparams.activation_func not in ACTIVATIONS or self.hparams.norm != 'layer':
            raise NotImplementedError(
                "ultra_pytorch_b200.DNN implements activation_func in %s with norm='layer' (got %r, %r); there is no "
                "fallback path" % (sorted(ACTIVATIONS), self.hparams.activation_func, self.hparams.norm))
        self._setup(feature_size, self.hparams.hidden_layer_sizes, extra_floats, self.hparams.activation_func)

    def _setup(self, feature_size, hidden, extra_floats, activation='elu'):
        self.feature_size = int(feature_size)
        self.output_sizes = list(hidden) + [1]
        # Build the same module sequence as the reference ON THE CPU first: with the same torch seed the
        # initial weights are bit-identical to the reference's (DNN.py:43-55).
        self.sequential = nn.Sequential()
        k = self.feature_size
        for j, n in enumerate(self.output_sizes):
            self.sequential.add_module('layer_norm{}'.format(j), nn.LayerNorm(k))
            self.sequential.add_module('linear{}'.format(j), nn.Linear(k, n))
            if j != len(self.output_sizes) - 1:
                self.sequential.add_module('act{}'.format(j), ACTIVATIONS[activation][1]())
            k = n
        self.engine = RankerEngine(self.feature_size, list(hidden), extra_floats=extra_floats,
                                   activation=ACTIVATIONS[activation][0])
        self._bind()

    def _bind(self):
        """Re-home every parameter (and its .grad) as a view of the engine's flat buffers."""
        eng = self.engine
        named = dict(self.sequential.named_parameters())
        with torch.no_grad():
            for name, off, shape in eng.layer_slices():
                p = named[name]
                cnt = p.numel()
                assert tuple(p.shape) == tuple(shape), (name, p.shape, shape)
                eng.params[off:off + cnt].copy_(p.detach().reshape(-1).to(eng.device, torch.float32))
                p.data = eng.params[off:off + cnt].view(shape)
                p.grad = eng.grads[off:off + cnt].view(shape)

    def _apply(self, fn, *args, **kwargs):
        # .to(device) / .cuda() on an already-resident module must not detach the views from the flat buffer
        super(DNN, self)._apply(fn, *args, **kwargs)
        named = dict(self.sequential.named_parameters())
        ok = all(named[n].data_ptr() == self.engine.params.data_ptr() + 4 * off
                 for n, off, _ in self.engine.layer_slices())
        if not ok:
            self._bind()
        return self

    def build(self, input_list, noisy_params=None, noise_rate=0.05, **kwargs):
        """input_list: L tensors [B, F] -> tuple of L tensors [B, 1] (DNN.py:58-88)."""
        eng = self.engine
        x = torch.cat([torch.as_tensor(t) for t in input_list], dim=0)
        x = x.to(device=eng.device, dtype=torch.float32).contiguous()
        if noisy_params is not None:                      # DBGD-family perturbation, in place (DNN.py:79-86)
            with torch.no_grad():
                for name, parameter in self.sequential.named_parameters():
                    if name in noisy_params:
                        parameter += (noisy_params[name] * noise_rate).to(device=eng.device)
        M = x.shape[0]
        out = torch.empty(M, dtype=torch.float32, device=eng.device)
        eng.forward(x, None, M, 1, training=False, scores=out)   # B = 1: scores[l] for row l
        return torch.split(out.view(M, 1), input_list[0].shape[0], dim=0)

    def forward(self, x):
        return torch.cat(self.build([x]), dim=0)
