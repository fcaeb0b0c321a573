"""ultra_pytorch_b200 - B200-native (sm_100a) training hot path for ULTRA (unbiased learning to rank).

A drop-in for ONE path of ULTR-Community/ULTRA_pytorch: the per-list DNN scorer forward/backward and the
propensity-weighted ranking losses, behind the reference's own plugin surface:

    "ranking_model":      "ultra_pytorch_b200.ranking_model.DNN"
    "learning_algorithm": "ultra_pytorch_b200.learning_algorithm.{NavieAlgorithm,IPWrank,DLA,PairDebias,LambdaRank}"

All compute runs in hand-written CUDA kernels reached through the C ABI in include/ultra_b200.h
(ultra_pytorch_b200/lib/libultra_b200.so).  There is no CPU or eager fallback.
"""
__version__ = "0.1.0"
