"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.

A plain-numpy restatement (closed forms + analytic gradients, no autograd, no torch ops) of the
reference's training hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file; the product path (ultra_pytorch_b200/) never does and
fails loudly when its CUDA library is missing.

Parity status: PINNED against outputs of the reference itself - tests/test_oracle_vs_golden.py checks
every function below against tests/golden/*.npz, which tests/golden/make_goldens.py produced by running
the unmodified reference (oracle/_ref) through BaseAlgorithm.train()/validation().  The reference's own
tests hold no golden vectors for this path.

Each function cites the reference file:line it restates (paths relative to the reference root).
`dt` is the arithmetic dtype: np.float32 mirrors the reference's fp32 path, np.float64 gives a
high-precision yardstick.
"""
import numpy as np

LN_EPS = 1e-5          # nn.LayerNorm default, ultra/ranking_model/DNN.py:46-47
ADAGRAD_EPS = 1e-10    # torch.optim.Adagrad default, ultra/learning_algorithm/ipw_rank.py:96
CLIP_EPS = 1e-6        # torch.nn.utils.clip_grad_norm_, ultra/learning_algorithm/base_algorithm.py:224


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def param_names(n_layers):
    """state_dict order of ultra/ranking_model/DNN.py:43-55 (layer_norm{j} then linear{j})."""
    names = []
    for j in range(n_layers):
        names += ["sequential.layer_norm%d.weight" % j, "sequential.layer_norm%d.bias" % j,
                  "sequential.linear%d.weight" % j, "sequential.linear%d.bias" % j]
    return names


def layer_sizes(feature_size, hidden):
    """(K_j, N_j) per linear layer; the last layer maps to 1 (DNN.py:36)."""
    outs = list(hidden) + [1]
    ks = [feature_size] + list(hidden)
    return list(zip(ks, outs))


# ----------------------------------------------------------------------------------------------
# A2: host gather  (ultra/learning_algorithm/base_algorithm.py:134-154)
# ----------------------------------------------------------------------------------------------
def gather_rows(features, docids_lb, dt=np.float32):
    """features [n_docs,F] (f64 in the feed), docids_lb int [L,B] with n_docs == PAD.
    Returns x [L*B, F] in position-major row order (row = l*B + b), cast like DNN.py:72-73."""
    F = features.shape[1]
    padded = np.concatenate([np.asarray(features), np.zeros((1, F))], axis=0)
    L, B = docids_lb.shape
    return padded[docids_lb.reshape(-1)].astype(dt)


# ----------------------------------------------------------------------------------------------
# A3: DNN forward / backward  (ultra/ranking_model/DNN.py:43-55,77)
# ----------------------------------------------------------------------------------------------
def elu(z):
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))


SELU_ALPHA = 1.6732632423543772848170429916717          # ultra/ranking_model/base_ranking_model.py:27-29
SELU_SCALE = 1.0507009873554804934193349852946


def activation(z, act):
    """hidden-layer activation, hparam activation_func (base_ranking_model.py:63-69; DNN.py:38-39, 52-53)"""
    if act == "elu":
        return elu(z)
    if act == "relu":
        return np.maximum(z, 0)
    if act == "selu":
        return SELU_SCALE * np.where(z >= 0, z, SELU_ALPHA * elu(z))
    if act == "tanh":
        return np.tanh(z)
    if act == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    raise ValueError(act)


def activation_grad(z, y, act, dt):
    """d activation / d z (autograd of the torch modules: nn.ELU / ReLU / Tanh / Sigmoid and the reference's selu)"""
    if act == "elu":
        return np.where(z > 0, dt(1.0), y + dt(1.0))
    if act == "relu":
        return (z > 0).astype(dt)
    if act == "selu":
        return np.where(z >= 0, dt(SELU_SCALE), y + dt(SELU_SCALE * SELU_ALPHA))
    if act == "tanh":
        return dt(1.0) - y * y
    if act == "sigmoid":
        return y * (dt(1.0) - y)
    raise ValueError(act)


def dnn_forward(x, params, n_layers, dt=np.float32, act="elu"):
    """x [M,K0].  params: dict name -> ndarray.  Returns (scores [M], cache)."""
    cache = []
    h = x.astype(dt)
    for j in range(n_layers):
        g = params["sequential.layer_norm%d.weight" % j].astype(dt)
        b = params["sequential.layer_norm%d.bias" % j].astype(dt)
        W = params["sequential.linear%d.weight" % j].astype(dt)
        c = params["sequential.linear%d.bias" % j].astype(dt)
        mean = h.mean(axis=1, keepdims=True, dtype=dt)
        var = ((h - mean) ** 2).mean(axis=1, keepdims=True, dtype=dt)
        rstd = (1.0 / np.sqrt(var + dt(LN_EPS))).astype(dt)
        xhat = ((h - mean) * rstd).astype(dt)
        a = (xhat * g + b).astype(dt)
        z = (a @ W.T + c).astype(dt)
        last = j == n_layers - 1
        y = z if last else activation(z, act).astype(dt)
        cache.append((xhat, rstd, a, z, y))
        h = y
    return h[:, 0], cache


def dnn_backward(dscores, cache, params, n_layers, dt=np.float32, act="elu"):
    """dscores [M] -> dict name -> grad (autograd of DNN.py:77 restated)."""
    grads = {}
    dy = dscores.astype(dt)[:, None]
    for j in reversed(range(n_layers)):
        xhat, rstd, a, z, y = cache[j]
        g = params["sequential.layer_norm%d.weight" % j].astype(dt)
        W = params["sequential.linear%d.weight" % j].astype(dt)
        last = j == n_layers - 1
        dz = dy if last else (dy * activation_grad(z, y, act, dt)).astype(dt)
        grads["sequential.linear%d.weight" % j] = (dz.T @ a).astype(dt)
        grads["sequential.linear%d.bias" % j] = dz.sum(axis=0).astype(dt)
        da = (dz @ W).astype(dt)
        grads["sequential.layer_norm%d.weight" % j] = (da * xhat).sum(axis=0).astype(dt)
        grads["sequential.layer_norm%d.bias" % j] = da.sum(axis=0).astype(dt)
        dxhat = (da * g).astype(dt)
        m1 = dxhat.mean(axis=1, keepdims=True, dtype=dt)
        m2 = (dxhat * xhat).mean(axis=1, keepdims=True, dtype=dt)
        dy = (rstd * (dxhat - m1 - xhat * m2)).astype(dt)
    return grads


def ranking_scores(features, docids_lb, params, n_layers, dt=np.float32, act="elu"):
    """BaseAlgorithm.ranking_model (base_algorithm.py:118-132): returns scores [B,L] and the cache."""
    L, B = docids_lb.shape
    x = gather_rows(features, docids_lb, dt)
    s, cache = dnn_forward(x, params, n_layers, dt, act)
    return s.reshape(L, B).T.copy(), cache


def scores_grad_to_rows(dscores_bl):
    """[B,L] gradient -> position-major row vector [L*B] (inverse of the reshape above)."""
    return np.ascontiguousarray(dscores_bl.T).reshape(-1)


# ----------------------------------------------------------------------------------------------
# A4: listwise softmax loss  (base_algorithm.py:18-30, 309-330)
# ----------------------------------------------------------------------------------------------
def log_softmax(s, dt):
    m = s.max(axis=1, keepdims=True)
    e = np.exp(s - m)
    return (s - m - np.log(e.sum(axis=1, keepdims=True, dtype=dt))).astype(dt)


def softmax_loss(scores, labels, pw=None, dt=np.float32):
    """Returns (loss, dloss/dscores [B,L], num = sum_b l_b, den = sum w).
    loss = sum_b [ -sum_l (w/W_b) log_softmax(s)_l * W_b ] / sum w;  pads are NOT masked."""
    s = scores.astype(dt)
    y = labels.astype(dt)
    pw = np.ones_like(y) if pw is None else pw.astype(dt)
    w = ((y + dt(1e-7)) * pw).astype(dt)
    Wb = w.sum(axis=1, keepdims=True, dtype=dt)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.nan_to_num(w / Wb).astype(dt)
    lsm = log_softmax(s, dt)
    per_list = (-(d * lsm).sum(axis=1, dtype=dt) * Wb[:, 0]).astype(dt)
    den = w.sum(dtype=dt)
    num = per_list.sum(dtype=dt)
    loss = num / den
    sm = np.exp(lsm)
    dsum = d.sum(axis=1, keepdims=True, dtype=dt)       # 1 unless the list is empty (then 0)
    grad = ((sm * dsum - d) * Wb / den).astype(dt)
    return dt(loss), grad, dt(num), dt(den)


# ----------------------------------------------------------------------------------------------
# A5: IPW weights  (ipw_rank.py:116-128, utils/propensity_estimator.py:22-42)
# ----------------------------------------------------------------------------------------------
def ipw_weights(clicks, table, dt=np.float32):
    B, L = clicks.shape
    idx = np.minimum(np.arange(L), len(table) - 1)
    t = np.asarray(table, dtype=np.float64)[idx].astype(dt)     # torch.as_tensor(list of python floats) -> f32
    return np.where(clicks > 0, t[None, :], dt(0.0)).astype(dt)


# ----------------------------------------------------------------------------------------------
# A8-A10: DLA  (dla.py:24-48, 179-266, 287-306)
# ----------------------------------------------------------------------------------------------
def dla_losses(scores, clicks, prop_w, prop_b, ranker_loss_weight=1.0, dt=np.float32):
    """prop_w [L] (= DenoisingNet.linear_layer.weight[0]), prop_b scalar.
    Returns dict(loss, rank_loss, exam_loss, dscores [B,L], dprop_w [L], dprop_b)."""
    s = scores.astype(dt)
    B, L = s.shape
    pre = (prop_w.astype(dt) + dt(prop_b)).astype(dt)          # Linear(onehot_l) = W[0,l] + b   (dla.py:32-46)
    prop = elu(pre).astype(dt)                                   # [L], identical for every list
    prop_bl = np.broadcast_to(prop[None, :], (B, L)).astype(dt)
    sm_p = np.exp(log_softmax(prop_bl, dt))
    pw = (sm_p[:, :1] / sm_p).astype(dt)                         # get_normalized_weights (dla.py:287-301)
    rank_loss, dscores, _, _ = softmax_loss(s, clicks, pw, dt)
    sm_s = np.exp(log_softmax(s, dt))
    rw = (sm_s[:, :1] / sm_s).astype(dt)
    exam_loss, dprop_bl, _, _ = softmax_loss(prop_bl, clicks, rw, dt)
    dprop = dprop_bl.sum(axis=0, dtype=dt)
    dpre = (dprop * np.where(pre > 0, dt(1.0), prop + dt(1.0))).astype(dt)
    return dict(loss=dt(exam_loss + dt(ranker_loss_weight) * rank_loss), rank_loss=rank_loss, exam_loss=exam_loss,
                dscores=(dt(ranker_loss_weight) * dscores).astype(dt), dprop_w=dpre, dprop_b=dpre.sum(dtype=dt))


# ----------------------------------------------------------------------------------------------
# A11: PairDebias  (pairwise_debias.py:106-174, base_algorithm.py:228-248)
# ----------------------------------------------------------------------------------------------
def softplus(x):
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def pairdebias(scores, clicks, t_plus, t_minus, em_step=0.05, reg_p=1.0, dt=np.float32):
    """i, j are DISPLAY positions.  Includes the reference's x batch_size factor ([B]*[B,1] -> [B,B]
    broadcast at base_algorithm.py:246-247).  Returns dict(loss, dscores, t_plus, t_minus, T_plus, T_minus)."""
    s = scores.astype(dt)
    c = clicks.astype(dt)
    B, L = s.shape
    tp = t_plus.reshape(-1).astype(dt)
    tm = t_minus.reshape(-1).astype(dt)
    mask = np.minimum(dt(1.0), np.maximum(c[:, :, None] - c[:, None, :], dt(0.0)))      # [B,i,j]
    diff = (s[:, None, :] - s[:, :, None]).astype(dt)                                    # s_j - s_i
    pl = softplus(diff).astype(dt)
    P = (dt(B) * (mask * pl).sum(axis=0, dtype=dt)).astype(dt)                           # [i,j]
    np.fill_diagonal(P, 0.0)
    loss = (P / tp[:, None] / tm[None, :]).sum(dtype=dt)
    T_plus = (P / tm[None, :]).sum(axis=1, dtype=dt)
    T_minus = (P / tp[:, None]).sum(axis=0, dtype=dt)
    sg = sigmoid(diff).astype(dt)                                                         # sigma(s_j - s_i)
    coef = (dt(B) * mask * sg / (tp[None, :, None] * tm[None, None, :])).astype(dt)
    offdiag = dt(1.0) - np.eye(L, dtype=dt)
    coef = coef * offdiag[None]
    dscores = (-coef.sum(axis=2, dtype=dt) + coef.sum(axis=1, dtype=dt)).astype(dt)
    with np.errstate(divide="ignore", invalid="ignore"):
        new_tp = ((1 - em_step) * tp + em_step * np.power(T_plus / T_plus[0], 1.0 / (reg_p + 1))).astype(dt)
        new_tm = ((1 - em_step) * tm + em_step * np.power(T_minus / T_minus[0], 1.0 / (reg_p + 1))).astype(dt)
    return dict(loss=dt(loss), dscores=dscores, t_plus=new_tp, t_minus=new_tm, T_plus=T_plus, T_minus=T_minus)


# ----------------------------------------------------------------------------------------------
# A12: LambdaRank  (lambda_rank.py:96-140, 247-291; utils/metrics.py:156-170)
# ----------------------------------------------------------------------------------------------
def safe_div(n, d):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(d == 0, np.zeros_like(n), n / d)


def lambdarank(scores, labels, t_plus, t_minus, sigma=1.0, em_step=0.05, reg_p=1.0, dt=np.float32):
    """i, j are PREDICTED-RANK positions (stable descending sort).  Reproduces BCE-with-logits applied to the
    probability p_ij and the batch-global natural-log IDCG.  Returns dict like pairdebias + idcg."""
    s = scores.astype(dt)
    y = labels.astype(dt)
    B, L = s.shape
    tp = t_plus.reshape(-1).astype(dt)
    tm = t_minus.reshape(-1).astype(dt)
    order = np.argsort(-s, axis=1, kind="stable")
    ps = np.take_along_axis(s, order, axis=1)
    ys = np.take_along_axis(y, order, axis=1)
    S = np.clip(ys[:, :, None] - ys[:, None, :], -1.0, 1.0).astype(dt)
    Pbar = (dt(0.5) * (dt(1.0) + S)).astype(dt)
    sij = (ps[:, :, None] - ps[:, None, :]).astype(dt)
    with np.errstate(over="ignore"):
        p = (dt(1.0) / (np.exp(-dt(sigma) * sij) + dt(1.0))).astype(dt)
    ideal = -np.sort(-y, axis=1)
    pos = np.arange(1, L + 1, dtype=dt)
    idcg = ((np.power(dt(2.0), ideal) - dt(1.0)) / np.log(pos + dt(1.0))[None, :]).sum(dtype=dt)   # ONE scalar
    gains = ((np.power(dt(2.0), ys) - dt(1.0)) / idcg).astype(dt)
    disc = (dt(1.0) / np.log2(np.arange(L, dtype=dt) + dt(2.0))).astype(dt)
    delta = (np.abs(gains[:, :, None] - gains[:, None, :]) * np.abs(disc[None, :, None] - disc[None, None, :])).astype(dt)
    term = (delta * (np.maximum(p, 0) - p * Pbar + np.log1p(np.exp(-np.abs(p))))).astype(dt)
    pair = term.sum(axis=0, dtype=dt)                                                       # [i,j]
    T_plus = (pair / tm[None, :]).sum(axis=1, dtype=dt)
    T_minus = (pair.T / tp[None, :]).sum(axis=1, dtype=dt)
    loss = safe_div(pair, tp[:, None] * tm[None, :]).sum(dtype=dt)
    # gradient wrt the sorted scores, then scattered back through the permutation
    inv = safe_div(np.ones((L, L), dtype=dt), tp[:, None] * tm[None, :])
    dterm_dp = (delta * (sigmoid(p) - Pbar)).astype(dt)
    dp_ds = (dt(sigma) * p * (dt(1.0) - p)).astype(dt)
    A = (dterm_dp * dp_ds * inv[None]).astype(dt)                                           # d/d(ps_i - ps_j)
    dps = (A.sum(axis=2, dtype=dt) - A.sum(axis=1, dtype=dt)).astype(dt)
    dscores = np.zeros_like(s)
    np.put_along_axis(dscores, order, dps, axis=1)
    new_tp = ((1 - em_step) * tp + em_step * np.power(safe_div(T_plus, T_plus[0]), 1.0 / (reg_p + 1))).astype(dt)
    new_tm = ((1 - em_step) * tm + em_step * np.power(safe_div(T_minus, T_minus[0]), 1.0 / (reg_p + 1))).astype(dt)
    return dict(loss=dt(loss), dscores=dscores, t_plus=new_tp, t_minus=new_tm, T_plus=T_plus, T_minus=T_minus,
                idcg=dt(idcg))


def prsrank(scores, labels, ipw_table, sigma=1.0, dt=np.float32):
    """PRSrank (prs_rank.py:94-151).  ipw_l = IPW_list[min(l, len-1)] for EVERY display position
    (getPropensityForOneList(use_non_clicked_data=True), propensity_estimator.py:22-42), pw = _safe_div(1, ipw).
    Sorted by predicted score: prs_rs = ipw_r * pw_s for r < s (triu, diagonal=1); p_rs = 1 / (exp(-sigma (s_r - s_s)) + 1);
    loss = sum_{b, r<s} delta_NDCG_rs * prs_rs * BCE(p_rs, (1 + clamp(y_r - y_s)) / 2) with torch's BCE (both logs clamped
    at -100); delta-NDCG with ONE batch-global natural-log IDCG (prs_rank.py:214-218, 228-231).  Gradient through
    binary_cross_entropy's backward ((p - t) / max(p (1 - p), 1e-12)) and the sigmoid."""
    s = scores.astype(dt)
    y = labels.astype(dt)
    B, L = s.shape
    tab = np.asarray(ipw_table, dtype=dt)
    ipw_pos = tab[np.minimum(np.arange(L), len(tab) - 1)]
    order = np.argsort(-s, axis=1, kind="stable")
    ps = np.take_along_axis(s, order, axis=1)
    ys = np.take_along_axis(y, order, axis=1)
    ipw = ipw_pos[order]                                        # ipw of the display position of the doc at rank r
    pw = safe_div(np.ones_like(ipw), ipw)
    tri = np.triu(np.ones((L, L), dtype=dt), k=1)
    prs = (ipw[:, :, None] * pw[:, None, :] * tri[None]).astype(dt)
    S = np.clip(ys[:, :, None] - ys[:, None, :], -1.0, 1.0).astype(dt)
    T = (dt(0.5) * (dt(1.0) + S)).astype(dt)
    sij = (ps[:, :, None] - ps[:, None, :]).astype(dt)
    with np.errstate(over="ignore"):
        p = (dt(1.0) / (np.exp(-dt(sigma) * sij) + dt(1.0))).astype(dt)
    ideal = -np.sort(-y, axis=1)
    pos = np.arange(1, L + 1, dtype=dt)
    idcg = ((np.power(dt(2.0), ideal) - dt(1.0)) / np.log(pos + dt(1.0))[None, :]).sum(dtype=dt)   # ONE scalar
    gains = ((np.power(dt(2.0), ys) - dt(1.0)) / idcg).astype(dt)
    disc = (dt(1.0) / np.log2(np.arange(L, dtype=dt) + dt(2.0))).astype(dt)
    delta = (np.abs(gains[:, :, None] - gains[:, None, :]) * np.abs(disc[None, :, None] - disc[None, None, :])).astype(dt)
    w = (delta * prs).astype(dt)
    with np.errstate(divide="ignore"):
        lp = np.maximum(np.log(p), dt(-100.0))
        l1p = np.maximum(np.log(dt(1.0) - p), dt(-100.0))
    loss = (-w * (T * lp + (dt(1.0) - T) * l1p)).sum(dtype=dt)
    pq = (p * (dt(1.0) - p)).astype(dt)
    A = (w * dt(sigma) * (p - T) * (pq / np.maximum(pq, dt(1e-12)))).astype(dt)            # d loss / d (ps_r - ps_s)
    dps = (A.sum(axis=2, dtype=dt) - A.sum(axis=1, dtype=dt)).astype(dt)
    dscores = np.zeros_like(s)
    np.put_along_axis(dscores, order, dps, axis=1)
    return dict(loss=dt(loss), dscores=dscores, idcg=dt(idcg))


# ----------------------------------------------------------------------------------------------
# A7: clip_grad_norm_ + Adagrad / SGD  (base_algorithm.py:208-226, dla.py:141-166)
# ----------------------------------------------------------------------------------------------
def regression_em(scores, clicks, propensity, uniforms, em_step=0.05, dt=np.float32):
    """One RegressionEM step on given scores (ultra/learning_algorithm/regression_EM.py:122-183; `sigmoid_prob_b` is the
    constant 0): E-step posteriors, pseudo-labels ceil(p_r1 - u) (:20-34), mean BCE-with-logits and its gradient, M-step."""
    s = np.asarray(scores, dtype=dt)
    c = np.asarray(clicks, dtype=dt)
    e = np.asarray(propensity, dtype=dt).reshape(1, -1)
    gamma = sigmoid(s).astype(dt)
    den = 1 - e * gamma
    p_e1_r0 = e * (1 - gamma) / den
    p_e0_r1 = (1 - e) * gamma / den
    p_r1 = c + (1 - c) * p_e0_r1
    labels = np.ceil(p_r1 - np.asarray(uniforms, dtype=dt)).astype(dt)
    bce = np.maximum(s, 0) - s * labels + np.log1p(np.exp(-np.abs(s)))
    n = s.size
    new_prop = (1 - em_step) * e + em_step * np.mean(c + (1 - c) * p_e1_r0, axis=0, keepdims=True)
    return {"loss": float(bce.astype(np.float64).sum() / n), "dscores": ((gamma - labels) / n).astype(dt),
            "labels": labels, "propensity": new_prop.astype(dt)}


def rank_metrics_per_list(scores_bl, labels_bl, docids_bl, n_docs, topn, max_label):
    """Per-list NDCG@n / ERR@n / MRR exactly as the reference's torch-CPU code evaluates them for one list
    (ultra/learning_algorithm/base_algorithm.py:88-116 PAD masking; ultra/utils/metrics.py:224-265 label validation,
    :191-221 DCG, :456-495 NDCG, :300-336 ERR, :268-298 MRR), in float32 with torch's CPU rules: cumsum / cumprod
    accumulate in double and round every prefix to float, everything else is float32 op by op; the sort is the stable
    descending one.  Returns [B, 2 n + 1] float32: ndcg | err | mrr (the batch means are left to the caller)."""
    f32 = np.float32
    s = np.asarray(scores_bl, dtype=f32).copy()
    y = np.asarray(labels_bl, dtype=f32).copy()
    B, L = s.shape
    if docids_bl is not None:
        s[np.asarray(docids_bl) == n_docs] = f32(-100000.0)
    topn = [min(int(n), L) for n in topn]
    n = len(topn)
    # the discount table comes from torch itself (metrics.py:212): torch's vectorised log2 and numpy's differ in the
    # last bit for some ranks, and the point of this function is bit-fidelity to the reference's CPU arithmetic
    import torch
    disc = (torch.tensor(1) / torch.log2(torch.arange(L, dtype=torch.float) + 2.0)).numpy()
    out = np.zeros((B, 2 * n + 1), dtype=f32)
    for b in range(B):
        p, v = s[b].copy(), y[b].copy()
        bad = ~(v >= 0)
        mn = p.min()
        v[bad] = 0
        p[bad] = f32(f32(-1e-6) + mn)
        order = np.argsort(-p.astype(np.float64), kind="stable")
        ideal = np.argsort(-v.astype(np.float64), kind="stable")
        ys, yi = v[order], v[ideal]
        cd = ci = 0.0
        dcg, idcg = {}, {}
        for r in range(max(topn)):
            cd += float(f32(f32(f32(2.0) ** ys[r] - f32(1.0)) * disc[r]))
            ci += float(f32(f32(f32(2.0) ** yi[r] - f32(1.0)) * disc[r]))
            dcg[r], idcg[r] = f32(cd), f32(ci)
        for q, t in enumerate(topn):
            out[b, q] = f32(0) if idcg[t - 1] == 0 else f32(dcg[t - 1] / idcg[t - 1])
        cp = 1.0
        err = [f32(0)] * n
        mrr = f32(0)
        for r in range(L):
            rel = f32(f32(f32(2.0) ** ys[r] - f32(1.0)) / f32(2.0 ** max_label))
            om = f32(f32(1.0) - rel)
            cp *= float(om)
            with np.errstate(divide="ignore", invalid="ignore"):
                nonrel = f32(f32(cp) / om)
            rr = f32(f32(1.0) / f32(r + 1))
            for q, t in enumerate(topn):
                rrq = rr if r < t else f32(rr * f32(0))
                err[q] = f32(err[q] + f32(f32(f32(rel * nonrel) * rrq) * f32(1.0)))
            mrr = max(mrr, f32((f32(1.0) if ys[r] >= 1.0 else f32(0.0)) * rr))
        out[b, n:2 * n] = err
        out[b, 2 * n] = mrr
    return out


def clip_grad_norm(grads, names, max_norm, dt=np.float32):
    """Returns (total_norm, clipped grads).  torch: coef = max_norm/(norm+1e-6) clamped to 1, always applied."""
    total = np.sqrt(sum((np.linalg.norm(grads[n].astype(dt).reshape(-1)) ** 2 for n in names))).astype(dt)
    coef = min(float(max_norm) / (float(total) + CLIP_EPS), 1.0)
    return total, {n: (grads[n].astype(dt) * dt(coef)).astype(dt) for n in names}


def adagrad_step(params, grads, state_sum, names, lr, dt=np.float32):
    """In place.  torch.optim.Adagrad with lr_decay=0, weight_decay=0, initial_accumulator_value=0."""
    for n in names:
        g = grads[n].astype(dt)
        state_sum[n] = (state_sum[n] + g * g).astype(dt)
        params[n] = (params[n] - dt(lr) * g / (np.sqrt(state_sum[n]) + dt(ADAGRAD_EPS))).astype(dt)


def sgd_step(params, grads, names, lr, dt=np.float32):
    for n in names:
        params[n] = (params[n] - dt(lr) * grads[n].astype(dt)).astype(dt)


# ----------------------------------------------------------------------------------------------
# whole train steps (what BaseAlgorithm.train does, per algorithm)
# ----------------------------------------------------------------------------------------------
class OracleTrainer:
    """Replays `train(input_feed)` of NA / IPW / DLA / PairDebias / LambdaRank / PRSrank / RegressionEM on the CPU.

    State mirrors the reference objects: params (ranker state_dict), Adagrad accumulators (persistent for
    NA/IPW/PairDebias/LambdaRank, re-created every step for DLA, dla.py:153-154), t_plus/t_minus, the
    DenoisingNet parameters."""

    def __init__(self, algo, params, feature_size, hidden, L_train, ipw_table=None, prop_params=None,
                 learning_rate=None, max_gradient_norm=5.0, sigma=1.0, em_step=0.05, reg_p=1.0, dt=np.float32,
                 l2_loss=0.0, activation_func="elu"):
        self.algo = algo
        self.dt = dt
        self.hidden = list(hidden)
        self.n_layers = len(hidden) + 1
        self.names = param_names(self.n_layers)
        self.params = {n: np.array(params[n], dtype=dt) for n in self.names}
        self.state_sum = {n: np.zeros_like(self.params[n]) for n in self.names}
        self.F = feature_size
        self.L = L_train
        defaults = {"na": 0.05, "ipw": 0.05, "dla": 0.05, "pairdebias": 0.005, "lambdarank": 0.05, "prsrank": 0.05,
                    "regem": 0.05}
        self.lr = defaults[algo] if learning_rate is None else learning_rate
        self.max_norm = max_gradient_norm
        self.sigma, self.em_step, self.reg_p = sigma, em_step, reg_p
        self.l2 = dt(l2_loss)              # hparam l2_loss (NA / IPW / DLA / PairDebias / RegressionEM)
        self.act = activation_func         # ranking_model hparam activation_func
        self.ipw_table = ipw_table
        if algo == "dla":
            self.prop_w = np.array(prop_params["linear_layer.weight"], dtype=dt).reshape(-1)
            self.prop_b = dt(np.array(prop_params["linear_layer.bias"]).reshape(-1)[0])
        if algo in ("pairdebias", "lambdarank"):
            self.t_plus = np.ones(L_train, dtype=dt)
            self.t_minus = np.ones(L_train, dtype=dt)
        if algo == "regem":
            self.propensity = (np.ones(L_train, dtype=dt) * dt(0.9)).reshape(1, -1)     # regression_EM.py:94-97
            self.uniforms = None            # the caller provides the step's uniform draws [B, L_train]
        self.last = {}

    def scores(self, features, docids_bl):
        L = docids_bl.shape[1]
        s, cache = ranking_scores(features, np.ascontiguousarray(docids_bl.T), self.params, self.n_layers, self.dt,
                                  self.act)
        return s, cache

    def train(self, features, docids_bl, labels_bl):
        """docids_bl / labels_bl are [B, L_feed]; only the first L_train positions are used."""
        dt = self.dt
        d = docids_bl[:, :self.L]
        y = labels_bl[:, :self.L].astype(dt)
        s, cache = self.scores(features, d)
        extra = {}
        if self.algo == "na":
            loss, ds, _, _ = softmax_loss(s, y, None, dt)
        elif self.algo == "ipw":
            pw = ipw_weights(y, self.ipw_table, dt)
            loss, ds, _, _ = softmax_loss(s, y, pw, dt)
        elif self.algo == "dla":
            r = dla_losses(s, y, self.prop_w, self.prop_b, 1.0, dt)
            loss, ds = r["loss"], r["dscores"]
            extra = r
        elif self.algo == "pairdebias":
            r = pairdebias(s, y, self.t_plus, self.t_minus, self.em_step, self.reg_p, dt)
            loss, ds = r["loss"], r["dscores"]
            self.t_plus, self.t_minus = r["t_plus"], r["t_minus"]
        elif self.algo == "lambdarank":
            r = lambdarank(s, y, self.t_plus, self.t_minus, self.sigma, self.em_step, self.reg_p, dt)
            loss, ds = r["loss"], r["dscores"]
            self.t_plus, self.t_minus = r["t_plus"], r["t_minus"]
        elif self.algo == "prsrank":
            r = prsrank(s, y, self.ipw_table, self.sigma, dt)
            loss, ds = r["loss"], r["dscores"]
        elif self.algo == "regem":
            r = regression_em(s, y, self.propensity, self.uniforms, self.em_step, dt)
            loss, ds = r["loss"], r["dscores"]
            self.propensity = r["propensity"]
        else:
            raise ValueError(self.algo)
        grads = dnn_backward(scores_grad_to_rows(ds), cache, self.params, self.n_layers, dt, self.act)
        if self.l2 > 0 and self.algo in ("na", "ipw", "dla", "pairdebias", "regem"):
            # loss += l2 * sum(p ** 2) / 2 for every ranker parameter (ipw_rank.py:153-157, navie_algorithm.py:110-114,
            # pairwise_debias.py:167-169, regression_EM.py:167-169, base_algorithm.py:332-333; dla.py:146-150 adds it to
            # rank_loss, whose weight in the total loss is ranker_loss_weight = 1 here) -> gradient l2 * p
            for n in self.names:
                loss = loss + self.l2 * dt(np.sum(self.params[n].astype(dt) ** 2) / 2)
                grads[n] = grads[n] + self.l2 * self.params[n]
        self.last = dict(scores=s, dscores=ds, grads=grads, loss=loss)
        # with l2_loss > 0 the reference hands clip_grad_norm_ the generator its L2 loop has already exhausted
        # (ipw_rank.py:153-159, navie_algorithm.py:108-116, pairwise_debias.py:166-171, regression_EM.py:164-178):
        # nothing is clipped; dla.py:161-163 builds new iterators and clips
        no_clip = self.l2 > 0 and self.algo in ("na", "ipw", "pairdebias", "regem")
        norm, clipped = clip_grad_norm(grads, self.names, self.max_norm, dt)
        if no_clip:
            clipped = {n: grads[n].astype(dt) for n in self.names}
        if self.algo == "dla":
            pg = {"w": extra["dprop_w"], "b": np.asarray(extra["dprop_b"])}
            _, pc = clip_grad_norm(pg, ["w", "b"], self.max_norm, dt)
            self.last["grad_prop_w"], self.last["grad_prop_b"] = pg["w"], pg["b"]
            # fresh Adagrad every step (dla.py:153-154): accumulators start from zero
            fresh = {n: np.zeros_like(self.params[n]) for n in self.names}
            adagrad_step(self.params, clipped, fresh, self.names, self.lr, dt)
            pp = {"w": self.prop_w, "b": np.asarray(self.prop_b, dtype=dt)}
            adagrad_step(pp, pc, {"w": np.zeros_like(self.prop_w), "b": np.zeros((), dtype=dt)}, ["w", "b"], self.lr, dt)
            self.prop_w, self.prop_b = pp["w"], dt(pp["b"])
        else:
            adagrad_step(self.params, clipped, self.state_sum, self.names, self.lr, dt)
        return float(loss)
