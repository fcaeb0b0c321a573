"""Shared test helpers.  The parity tolerance is stated ONCE here:

    |a - b| <= rtol * max(|ref|, mean|ref|)        rtol = 1e-5 for fp32 scores / gradients

i.e. np.allclose(rtol=1e-5, atol=1e-5*mean|ref|): scores and gradients cross zero, so a purely
element-wise relative error is unbounded even between fp32 and fp64 evaluations of the reference.
"""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in g:
        if k.endswith("/features"):          # compact goldens store the (float32-exact) feature rows as float32
            g[k] = g[k].astype(np.float64)
    return g


def sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def scaled_err(a, ref, floor=0.0):
    """max |a-ref| / max(|ref|, mean|ref|, floor) - the quantity bounded by RTOL.

    `floor` is only used for gradient tensors that are mathematically zero (e.g. the last bias under a
    softmax loss: sum_l dscore_bl == 0), where the reference itself only holds rounding noise; callers pass
    0.1 * max|g| over ALL parameter gradients of the step (the scale clip_grad_norm_ works at)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = np.maximum(np.maximum(np.abs(ref), np.mean(np.abs(ref))), floor)
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(a - ref) / scale))


def grad_floor(grads):
    return 0.1 * max(float(np.max(np.abs(v))) for v in grads.values() if np.size(v))


def assert_close(a, ref, rtol=RTOL, what="", floor=0.0):
    e = scaled_err(a, ref, floor)
    assert e <= rtol, "%s: scaled error %.3e > %.1e" % (what, e, rtol)


def golden_hparams(g):
    """non-default algorithm hparams a golden was generated with (meta_hparams, e.g. "l2_loss=0.01") as keyword
    arguments of oracle.OracleTrainer"""
    text = str(g["meta_hparams"]) if "meta_hparams" in g else ""
    out = {}
    for item in filter(None, text.split(",")):
        k, v = item.split("=")
        out[k] = float(v)
    for item in filter(None, (str(g["meta_model_hparams"]) if "meta_model_hparams" in g else "").split(",")):
        k, v = item.split("=")
        out[k] = v                              # activation_func=...
    return out
