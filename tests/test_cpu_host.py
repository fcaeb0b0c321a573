"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, host-side helpers
(hparams parser, metrics restatement vs the reference's goldens) behave like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import golden_names, load_golden, sub

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ultra_b200.h")).read()
    return sorted(set(re.findall(r"UB200_API[^;(]*?\b(ub200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ultra_pytorch_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    lib.ub200_abi_version.restype = ctypes.c_int
    assert lib.ub200_abi_version() == 1


def test_ctypes_table_matches_header():
    from ultra_pytorch_b200 import _capi
    assert sorted(_capi.SIGNATURES.keys()) == _header_symbols()
    # argument counts agree with the header declarations
    text = open(os.path.join(ROOT, "include", "ultra_b200.h")).read()
    for name, (_, args) in _capi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, re.S)
        decl = m.group(1).strip()
        n = 0 if decl in ("void", "") else len(decl.split(","))
        assert n == len(args), (name, n, len(args))


def test_param_count_and_workspace_sizes_without_gpu():
    from ultra_pytorch_b200 import _capi
    hid = _capi.int_array([256, 128, 64])
    assert _capi.lib.ub200_mlp_param_count(136, hid, 3) == 77457          # SURVEY.md section 8 (config 2)
    hid2 = _capi.int_array([512, 256, 128])
    assert _capi.lib.ub200_mlp_param_count(136, hid2, 3) == 236561
    assert _capi.lib.ub200_mlp_param_count(700, hid2, 3) == 526457
    assert _capi.lib.ub200_mlp_workspace_bytes(40, 256, 136, hid, 3, 1) > 0
    assert _capi.lib.ub200_mlp_param_count(0, hid, 3) == 0               # bad spec -> 0, no crash


def test_product_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ultra_pytorch_b200._capi import UltraB200Error
    from ultra_pytorch_b200.engine import RankerEngine
    with pytest.raises(UltraB200Error):
        RankerEngine(8, [4])


def test_hparams_parser():
    from ultra_pytorch_b200.hparams import HParams
    h = HParams(hidden_layer_sizes=[512, 256, 128], activation_func='elu', learning_rate=0.05, flag=False,
                regulation_p=1)
    h.parse("hidden_layer_sizes=[256, 128,64],learning_rate=0.1,unknown=3,flag=True,regulation_p=2")
    assert h.hidden_layer_sizes == [256, 128, 64] and h.learning_rate == 0.1 and h.flag is True
    assert h.regulation_p == 2 and not hasattr(h, "unknown")
    h.parse("")
    h.parse("activation_func=relu")
    assert h.activation_func == "relu" and h.hidden_layer_sizes == [256, 128, 64]


@pytest.mark.parametrize("name", golden_names())
def test_metric_batch_means_reproduce_reference_values(name):
    """Host side of the validation metrics (ultra_pytorch_b200/metrics.py: batch_means) fed with the oracle's per-list
    values computed from the REFERENCE's own scores: NDCG and MRR must reproduce the reference's numbers bit for bit,
    ERR to 1e-6 (torch.sum's vectorised order inside a list is not reproducible; the per-list chain is)."""
    from oracle import ultra_oracle as uo
    from ultra_pytorch_b200 import metrics as m
    g = load_golden(name)
    topn = [1, 3, 5, 10]
    L = g["valid/scores"].shape[1]
    per_list = uo.rank_metrics_per_list(g["valid/scores"], g["valid/labels"], g["valid/docids"],
                                        g["valid/features"].shape[0], topn, 4.0)
    vals = m.batch_means(torch.from_numpy(per_list), L, topn)
    for metric in ("ndcg", "err", "mrr"):
        for n, v in zip(topn, vals[metric]):
            ref = float(g["valid/metric/%s_%d" % (metric, n)])
            if metric == "err":
                assert abs(v.item() - ref) <= 1e-6 * max(1.0, abs(ref)), (metric, n, v.item(), ref)
            else:
                assert v.item() == ref, (metric, n, v.item(), ref)


def test_other_metric_keys_are_delegated_to_the_reference_module():
    from oracle import ref_shim
    from ultra_pytorch_b200 import metrics as m
    if not ref_shim.available():
        with pytest.raises(NotImplementedError):
            m.reference_metric_fn("map", [1, 3])
        pytest.skip("oracle/_ref not installed")
    if torch.cuda.is_available():
        pytest.skip("the reference's metrics module pins its tensors to cuda when a GPU is visible")
    ultra = ref_shim.load()
    ultra.utils.metrics.RankingMetricKey.MAX_LABEL = 4.0
    rs = np.random.RandomState(0)
    scores = torch.from_numpy(rs.randn(17, 12).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, 5, size=(17, 12)).astype(np.float32))
    for metric in ("arp", "precision", "map"):
        ours = m.reference_metric_fn(metric, [1, 3, 5, 10])(labels, scores, None)
        ref = ultra.utils.make_ranking_metric_fn(metric, [1, 3, 5, 10])(labels, scores, None)
        assert torch.equal(torch.as_tensor(ours), torch.as_tensor(ref)), metric


def test_host_packer_pool_is_exact_for_any_size_and_thread_count():
    """csrc/hostpack.cpp: the f64 -> f32 conversion on the persistent worker pool (AVX2 + streaming stores, scalar
    head/tail) equals numpy's cast bit for bit for ragged sizes, unaligned destinations, any thread count, and after
    the workers went to sleep; the packed feed has the documented layout."""
    import ctypes
    import time
    from ultra_pytorch_b200 import _capi
    lib = _capi.lib
    rs = np.random.RandomState(0)
    for trial in range(120):
        n = int(rs.randint(1, 300000)) if trial % 3 else int(rs.randint(1, 70))
        src = rs.randn(n) * 10.0 ** rs.randint(-3, 3)
        off = int(rs.randint(0, 8))
        buf = np.empty(n + 8, np.float32)
        dst = buf[off:off + n]
        assert lib.ub200_convert_f64_f32_host(src.ctypes.data, dst.ctypes.data, n, int(rs.randint(1, 9))) == 0
        assert np.array_equal(dst, src.astype(np.float32)), (trial, n)
    F, L, B = 24, 7, 33
    n = 200
    feats = rs.uniform(-1, 1, (n, F))
    d = [rs.randint(0, n + 1, B).astype(np.float32) for _ in range(L)]
    y = [rs.rand(B).astype(np.float32) for _ in range(L)]
    nb = lib.ub200_feed_bytes(n, F, L, B)
    out = np.full(nb, 255, np.uint8)
    PtrArr = ctypes.c_void_p * L
    time.sleep(0.05)                       # the pool's workers fall asleep after ~2 ms without work
    assert lib.ub200_pack_feed_host(feats.ctypes.data, n, F, PtrArr(*[x.ctypes.data for x in d]),
                                    PtrArr(*[x.ctypes.data for x in y]), L, B, out.ctypes.data, nb, 4) == 0
    off_f = (8 * L * B + 255) // 256 * 256
    hf = out[off_f:].view(np.float32).reshape(n + 1, F)
    assert np.array_equal(hf[:n], feats.astype(np.float32)) and not hf[n].any()
    assert np.array_equal(out[:4 * L * B].view(np.int32).reshape(L, B), np.stack(d).astype(np.int32))
    assert np.array_equal(out[4 * L * B:8 * L * B].view(np.float32).reshape(B, L), np.stack(y).T)


def test_column_ptrs_address_the_same_data_as_the_individual_arrays():
    """engine.column_ptrs: the stacked-block pointer array packs exactly like one pointer per feed array, for the
    homogeneous fast path and for the ragged / mixed-dtype fallback."""
    import ctypes
    from ultra_pytorch_b200 import _capi
    from ultra_pytorch_b200.engine import column_ptrs
    lib = _capi.lib
    rs = np.random.RandomState(1)
    L, B = 9, 37
    d = [rs.randint(0, 500, B).astype(np.float32) for _ in range(L)]
    y = [rs.rand(B).astype(np.float32) for _ in range(L)]
    ref = np.zeros(8 * L * B, np.uint8)
    P = ctypes.c_void_p * L
    assert lib.ub200_pack_ids_host(P(*[x.ctypes.data for x in d]), P(*[x.ctypes.data for x in y]), L, B, 500,
                                   ref.ctypes.data, ref.size) == 0
    for dd, yy in ((d, y), ([x.astype(np.float64) for x in d], [list(map(float, x)) for x in y])):
        out = np.full(8 * L * B, 7, np.uint8)
        dp, kd = column_ptrs(dd, B)
        lp, kl = column_ptrs(yy, B)
        assert lib.ub200_pack_ids_host(dp, lp, L, B, 500, out.ctypes.data, out.size) == 0
        assert np.array_equal(out, ref)
    # ids beyond the PAD row (or negative / NaN) are an error, like np.take's IndexError in the reference
    # (base_algorithm.py:150): the kernels gather rows unchecked
    bad = [x.copy() for x in d]
    bad[3][5] = 501.0
    dp, kd = column_ptrs(bad, B)
    lp, kl = column_ptrs(y, B)
    assert lib.ub200_pack_ids_host(dp, lp, L, B, 500, out.ctypes.data, out.size) == 5
    assert b"outside" in lib.ub200_last_error()
    bad[3][5] = -1.0
    dp, kd = column_ptrs(bad, B)
    assert lib.ub200_pack_ids_host(dp, lp, L, B, 500, out.ctypes.data, out.size) == 5
    with pytest.raises(Exception):
        column_ptrs([d[0], d[1][:5]], B)          # ragged arrays are not a valid feed


def test_staged_views_layout_and_cache():
    """engine.make_staged / StagedCache: the three views alias the staging buffer at the documented offsets, the same
    (buffer, L, B, n_docs) returns the same object, anything else a fresh one."""
    from ultra_pytorch_b200.engine import StagedCache, make_staged
    L, B, n_docs, F = 5, 7, 11, 3
    off_f = (8 * L * B + 255) // 256 * 256
    total = off_f + 4 * (n_docs + 1) * F
    buf = torch.zeros(total + 100, dtype=torch.uint8)
    raw = buf.numpy()
    raw[:4 * L * B].view(np.int32)[:] = np.arange(L * B)
    raw[4 * L * B:8 * L * B].view(np.float32)[:] = np.arange(L * B) + 0.5
    raw[off_f:total].view(np.float32)[:] = -np.arange((n_docs + 1) * F)
    st = make_staged(buf, L, B, n_docs, F)
    assert st.docid.dtype == torch.int32 and tuple(st.docid.shape) == (L, B) and st.docid[2, 3].item() == 2 * B + 3
    assert tuple(st.labels.shape) == (B, L) and st.labels[4, 1].item() == 4 * L + 1 + 0.5
    assert tuple(st.feats.shape) == (n_docs + 1, F) and st.feats[n_docs, F - 1].item() == -((n_docs + 1) * F - 1)
    assert (st.B, st.L, st.n_docs, st.h2d_bytes) == (B, L, n_docs, total)
    assert st.docid.data_ptr() == buf.data_ptr() and st.feats.data_ptr() == buf.data_ptr() + off_f
    cache = StagedCache(limit=3)
    a = cache.get(buf, L, B, n_docs, F)
    assert cache.get(buf, L, B, n_docs, F) is a
    b = cache.get(buf, L, B, n_docs - 1, F)
    assert b is not a and b.n_docs == n_docs - 1
    other = torch.zeros(total + 100, dtype=torch.uint8)
    c = cache.get(other, L, B, n_docs, F)
    assert c is not a and c.docid.data_ptr() == other.data_ptr()
    cache.get(buf, L, B, n_docs - 2, F)                 # 4th entry: the cache starts over
    assert len(cache.entries) == 1 and cache.get(buf, L, B, n_docs, F) is not a


def test_column_ptrs2_addresses_both_halves_of_one_block():
    """engine.column_ptrs2: docid and label pointer arrays into ONE stacked [2 L, B] copy of the feed's per-position arrays"""
    import ctypes
    from ultra_pytorch_b200 import engine as eng
    L, B = 5, 7
    d = [np.arange(B, dtype=np.float32) + 10 * l for l in range(L)]
    y = [np.arange(B, dtype=np.float32) + 100 + 10 * l for l in range(L)]
    dp, lp, keep = eng.column_ptrs2(d, y, B)
    for l in range(L):
        a = np.ctypeslib.as_array(ctypes.cast(dp[l], ctypes.POINTER(ctypes.c_float)), (B,))
        b = np.ctypeslib.as_array(ctypes.cast(lp[l], ctypes.POINTER(ctypes.c_float)), (B,))
        assert (a == d[l]).all() and (b == y[l]).all()
    with pytest.raises(ValueError):
        eng.column_ptrs2(d, y[:-1], B)
