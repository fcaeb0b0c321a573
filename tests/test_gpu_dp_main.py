"""Data parallel through the reference's UNMODIFIED main.py: `torchrun --nproc-per-node 2` over
oracle/_ref/main.py with the B200 plugin selected in the settings JSON.  The plugin bootstraps the process group, the
replicas stay equal (every rank dumps its parameters at exit through UB200_DP_DUMP) and exactly one checkpoint exists."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_through_unmodified_main_py(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if not ref_shim.available():
        pytest.skip("oracle/_ref not installed")
    settings = {
        "train_input_feed": "ultra.input_layer.ClickSimulationFeed", "train_input_hparams": "",
        "valid_input_feed": "ultra.input_layer.DirectLabelFeed", "valid_input_hparams": "",
        "test_input_feed": "ultra.input_layer.DirectLabelFeed", "test_input_hparams": "",
        "ranking_model": "ultra_pytorch_b200.ranking_model.DNN", "ranking_model_hparams": "hidden_layer_sizes=[64, 32]",
        "learning_algorithm": "ultra_pytorch_b200.learning_algorithm.IPWrank", "learning_algorithm_hparams": "",
        "metrics": ["err", "ndcg"], "metrics_topn": [1, 3, 5, 10], "objective_metric": "ndcg_10",
    }
    sfile = os.path.join(str(tmp_path), "settings.json")
    with open(sfile, "w") as f:
        json.dump(settings, f)
    model_dir = os.path.join(str(tmp_path), "model") + "/"
    os.makedirs(model_dir)
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env["UB200_DP_DUMP"] = os.path.join(str(tmp_path), "params_rank%d.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "oracle", "run_ref_main.py"), "--data_dir=./tests/data/",
           "--model_dir=" + model_dir, "--output_dir=" + os.path.join(str(tmp_path), "out") + "/",
           "--setting_file=" + sfile, "--batch_size=16", "--max_train_iteration=12", "--steps_per_checkpoint=6"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    ckpts = [n for n in os.listdir(model_dir) if n.endswith(".ckpt")]
    assert ckpts == ["ultra_pytorch_b200.learning_algorithm.IPWrank.ckpt"], os.listdir(model_dir)
    p0 = torch.load(os.path.join(str(tmp_path), "params_rank0.pt"))
    p1 = torch.load(os.path.join(str(tmp_path), "params_rank1.pt"))
    assert torch.equal(p0, p1)
    sd = torch.load(os.path.join(model_dir, ckpts[0]), map_location="cpu")
    assert all(torch.isfinite(v).all() for v in sd.values())
