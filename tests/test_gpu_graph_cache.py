"""GPU test of the CUDA-graph / buffer lifetime rule (engine.generation): batch sizes vary from step to step under
DirectLabelFeed (direct_label_feed.py:117-125 drops lists without relevant documents), so more (L, B) shapes appear
than the engine keeps workspaces for.  A graph captured while a workspace was alive must never be replayed after that
workspace has been released - training with graphs and a tiny shape cache must stay bit-identical to eager training."""
import types

import pytest
import torch

from ultra_pytorch_b200 import synth

pytestmark = pytest.mark.gpu


def test_variable_batch_sizes_with_graphs_match_eager():
    import ultra_pytorch_b200.learning_algorithm as la
    la.B200Algorithm.VERBOSE = False
    F, L = 136, 10
    settings = {"learning_algorithm_hparams": "", "ranking_model": "ultra_pytorch_b200.ranking_model.DNN",
                "ranking_model_hparams": "hidden_layer_sizes=[64, 64]", "selection_bias_cutoff": L,
                "max_candidate_num": L, "metrics": ["ndcg"], "metrics_topn": [1, 3]}
    sizes = [8, 9, 10, 11, 12, 13]                      # 6 shapes against a cache of 3
    losses = {}
    for mode in ("graph", "eager"):
        torch.manual_seed(0)
        model = la.NavieAlgorithm(types.SimpleNamespace(feature_size=F), settings)
        model.USE_GRAPH = mode == "graph"
        model.engine.MAX_SHAPES = 3
        out = []
        for rnd in range(5):                            # a shape is captured on its third visit
            for B in sizes:
                feed = synth.make_feed(100 * rnd + B, F, L, B, labels="graded")
                loss, _, _ = model.train(feed)
                out.append(loss)
        torch.cuda.synchronize()
        losses[mode] = (out, model.engine.params.clone())
        if mode == "graph":
            assert model.engine.generation > 0          # evictions happened, graphs were invalidated
    assert losses["graph"][0] == losses["eager"][0]
    assert torch.equal(losses["graph"][1], losses["eager"][1])
