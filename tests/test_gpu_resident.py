"""GPU test of the device-resident data set path (input_layer/resident.py): a feed that emits the data set's whole
feature matrix + global doc ids trains bit-identically to the per-batch-copy feed, while a step moves only ids and
labels to the device."""
import json
import os
import random
import types

import numpy as np
import pytest
import torch

from tests.test_click_feed import PBM, FakeData

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", ["NavieAlgorithm", "LambdaRank"])
def test_resident_dataset_trains_bit_identically(algo, tmp_path):
    import ultra_pytorch_b200.learning_algorithm as la
    from ultra_pytorch_b200.input_layer import ClickSimulationFeed
    la.B200Algorithm.VERBOSE = False
    L, F, B = 12, 136, 48
    ds = FakeData(200, L, F)
    p = os.path.join(str(tmp_path), "pbm.json")
    with open(p, "w") as f:
        json.dump(PBM, f)
    settings = {"learning_algorithm_hparams": "", "ranking_model": "ultra_pytorch_b200.ranking_model.DNN",
                "ranking_model_hparams": "hidden_layer_sizes=[64, 32]", "selection_bias_cutoff": L,
                "max_candidate_num": L, "metrics": ["ndcg"], "metrics_topn": [1, 3]}
    runs = []
    for hp in ("", "resident_features=True"):
        torch.manual_seed(0)
        random.seed(0)
        model = getattr(la, algo)(types.SimpleNamespace(feature_size=F), settings)
        feed = ClickSimulationFeed(model, B, "click_model_json=%s,oracle_mode=True,%s" % (p, hp))
        losses = []
        for step in range(5):                        # steps 3+ replay the captured CUDA graph
            f, _ = feed.get_next_batch(step * B % 150, ds)
            loss, _, _ = model.train(f)
            losses.append(loss)
        _, scores, summary = model.validation(feed.get_next_batch(7, ds)[0])
        runs.append((losses, model.engine.params.clone(), scores.clone(), dict(summary), model.last_h2d_bytes))
    (la_, pa, sa, ma, ha), (lb, pb, sb, mb, hb) = runs
    assert la_ == lb
    assert torch.equal(pa, pb) and torch.equal(sa, sb) and ma == mb
    assert hb == 8 * L * B and ha > 20 * hb          # ids + labels only vs ids + labels + feature rows
