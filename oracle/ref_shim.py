"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Makes the UNMODIFIED reference copy under oracle/_ref importable on this image
(torch 2.11 / numpy 2.3 / no tensorflow).  Three non-arithmetic shims:

1. ultra/ranking_model/base_ranking_model.py:8 does `import tensorflow as tf` (unused): a stub
   module is placed in sys.modules while `ultra` is imported and removed afterwards.
2. ultra/learning_algorithm/base_algorithm.py:186 (and pairwise_debias.py:127) call
   torch.as_tensor(list-of-float32-ndarray, dtype=int64), which newer numpy/torch reject:
   convert through np.asarray(...).astype(int64) first.
3. ipw_rank.py:164 / navie_algorithm.py:118 call nn.utils.clip_grad_value_ on a tensor without
   .grad (a no-op on torch 1.9, raises on torch 2.11): make it a no-op again.

Usage:  import oracle.ref_shim as rs; ultra = rs.load()      (cwd is switched to oracle/_ref by
`rs.ref_cwd()` where the reference needs its cwd-relative JSON fixtures.)
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")

_loaded = None


def available():
    return os.path.isdir(os.path.join(REF_COPY, "ultra"))


def load():
    """Import and return the reference `ultra` package (CPU path)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref missing: run `python oracle/install_ref.py` where /root/reference exists")
    import torch.utils.tensorboard  # noqa: F401  must be loaded before the tf stub exists
    sys.dont_write_bytecode = True
    if REF_COPY not in sys.path:
        sys.path.insert(0, REF_COPY)
    had_tf = "tensorflow" in sys.modules
    if not had_tf:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")
    try:
        import ultra  # noqa
        import ultra.utils  # noqa
        import ultra.learning_algorithm  # noqa
        import ultra.ranking_model  # noqa
        import ultra.input_layer  # noqa
    finally:
        if not had_tf:
            del sys.modules["tensorflow"]

    _as_tensor = torch.as_tensor

    def as_tensor(data, dtype=None, device=None):
        if isinstance(data, list) and data and isinstance(data[0], np.ndarray):
            data = np.asarray(data)
            if dtype == torch.int64:
                data = data.astype(np.int64)
        return _as_tensor(data, dtype=dtype, device=device)

    torch.as_tensor = as_tensor

    _clip = torch.nn.utils.clip_grad_value_

    def clip_grad_value_(parameters, clip_value, foreach=None):
        if isinstance(parameters, torch.Tensor):
            parameters = [parameters]
        parameters = [p for p in parameters if p.grad is not None]
        if parameters:
            return _clip(parameters, clip_value, foreach=foreach)

    torch.nn.utils.clip_grad_value_ = clip_grad_value_
    _loaded = ultra
    return ultra


@contextlib.contextmanager
def ref_cwd():
    """The reference resolves ./example/... and creates ./runs/ relative to the cwd."""
    old = os.getcwd()
    os.chdir(REF_COPY)
    try:
        yield
    finally:
        os.chdir(old)


def synthetic_raw_data(ultra, n_queries, list_len, feature_size, seed=1, max_label=4, ragged=False):
    """Seeded in-memory Raw_data: features U(-1,1), labels UniformInt{0..max_label}."""
    rs = np.random.RandomState(seed)
    ds = ultra.utils.data_utils.Raw_data()
    ds.feature_size = feature_size
    ds.rank_list_size = list_len
    ds.features = []
    ds.dids = []
    ds.initial_list = []
    ds.labels = []
    ds.qids = []
    ds.initial_list_lengths = []
    doc = 0
    for q in range(n_queries):
        n = list_len if not ragged else int(rs.randint(max(1, list_len // 2), list_len + 1))
        feats = rs.uniform(-1.0, 1.0, size=(n, feature_size)).astype(np.float32)
        ds.features.extend(feats.astype(np.float64).tolist())
        ds.dids.extend(["d%d" % (doc + i) for i in range(n)])
        ds.initial_list.append(list(range(doc, doc + n)))
        ds.labels.append(rs.randint(0, max_label + 1, size=n).astype(float).tolist())
        ds.qids.append("q%d" % q)
        ds.initial_list_lengths.append(n)
        doc += n
    ultra.utils.metrics.RankingMetricKey.MAX_LABEL = float(max_label)
    ds.pad(list_len)
    return ds
