"""B200-native drop-in for `ultra.ranking_model.Linear` (reference: ultra/ranking_model/Linear.py:11-78).

`LayerNorm(F) -> Linear(F, 1)`: the DNN ranker without hidden layers, so it runs on the same kernels (the final-layer
row kernels of csrc/mlp.cu read the gathered feature rows directly).  Same constructor `(hparams_str, feature_size)`,
hparam `norm` ('layer'), `build()` contract and `state_dict()` keys (`sequential.layer_norm0.*`,
`sequential.linear0.*`) as the reference.
"""
import torch.nn as nn

from ..hparams import HParams
from .DNN import DNN


class Linear(DNN):
    def __init__(self, hparams_str, feature_size, extra_floats=0):
        nn.Module.__init__(self)
        self.hparams = HParams(norm="layer")              # Linear.py:26-28
        self.hparams.parse(hparams_str)
        if self.hparams.norm != 'layer':
            raise NotImplementedError("ultra_pytorch_b200.Linear implements norm='layer' (the reference default); "
                                      "got %r and there is no fallback path" % (self.hparams.norm,))
        self._setup(feature_size, [], extra_floats)
