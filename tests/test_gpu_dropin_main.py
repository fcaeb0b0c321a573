"""Drop-in test: the reference's UNMODIFIED main.py (oracle/_ref) trains, checkpoints, reloads and ranks with the
B200 plugin selected purely through the two class-path strings of the settings JSON."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SETTINGS = {
    "ipw": ("ultra_pytorch_b200.learning_algorithm.IPWrank", "ultra.input_layer.ClickSimulationFeed"),
    "dla": ("ultra_pytorch_b200.learning_algorithm.DLA", "ultra.input_layer.ClickSimulationFeed"),
    "lambdarank": ("ultra_pytorch_b200.learning_algorithm.LambdaRank", "ultra.input_layer.ClickSimulationFeed"),
    "na": ("ultra_pytorch_b200.learning_algorithm.NavieAlgorithm", "ultra.input_layer.DirectLabelFeed"),
    # the vectorised feed drop-in together with the B200 algorithm
    "pairdebias_b200feed": ("ultra_pytorch_b200.learning_algorithm.PairDebias",
                            "ultra_pytorch_b200.input_layer.ClickSimulationFeed"),
    # BASELINE config 5: online simulation + DLA, with the reference's own feed (it calls validation(feed, True) and
    # re-ranks on the host) and with the drop-in whose Plackett-Luce sampling runs on the GPU
    "dla_online_reffeed": ("ultra_pytorch_b200.learning_algorithm.DLA",
                           "ultra.input_layer.StochasticOnlineSimulationFeed"),
    "dla_online_b200feed": ("ultra_pytorch_b200.learning_algorithm.DLA",
                            "ultra_pytorch_b200.input_layer.StochasticOnlineSimulationFeed"),
    # every feed replaced by its drop-in, data sets resident in HBM (train: batches assembled on the device)
    "ipw_all_b200feeds": ("ultra_pytorch_b200.learning_algorithm.IPWrank",
                          "ultra_pytorch_b200.input_layer.ClickSimulationFeed"),
    # RegressionEM
    "regem": ("ultra_pytorch_b200.learning_algorithm.RegressionEM", "ultra.input_layer.ClickSimulationFeed"),
    # the Linear ranker
    "ipw_linear": ("ultra_pytorch_b200.learning_algorithm.IPWrank", "ultra.input_layer.ClickSimulationFeed"),
}


def _run(args, timeout=600):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    return subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_main.py")] + args, env=env,
                          capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("algo", sorted(SETTINGS))
def test_unmodified_main_py_drives_the_plugin(algo, tmp_path):
    if not ref_shim.available():
        pytest.skip("oracle/_ref not installed (python oracle/install_ref.py needs /root/reference)")
    cls, feed = SETTINGS[algo]
    all_ours = algo == "ipw_all_b200feeds"
    eval_feed = "ultra_pytorch_b200.input_layer.DirectLabelFeed" if all_ours else "ultra.input_layer.DirectLabelFeed"
    settings = {
        "train_input_feed": feed, "train_input_hparams": "device_batches=True" if all_ours else "",
        "valid_input_feed": eval_feed, "valid_input_hparams": "resident_features=True" if all_ours else "",
        "test_input_feed": eval_feed, "test_input_hparams": "resident_features=True" if all_ours else "",
        "ranking_model": "ultra_pytorch_b200.ranking_model.%s" % ("Linear" if algo.endswith("_linear") else "DNN"),
        "ranking_model_hparams": "" if algo.endswith("_linear") else "hidden_layer_sizes=[64, 32]",
        "learning_algorithm": cls, "learning_algorithm_hparams": "",
        "metrics": ["err", "ndcg"], "metrics_topn": [1, 3, 5, 10], "objective_metric": "ndcg_10",
    }
    sfile = os.path.join(str(tmp_path), "settings.json")
    with open(sfile, "w") as f:
        json.dump(settings, f)
    model_dir = os.path.join(str(tmp_path), "model") + "/"
    out_dir = os.path.join(str(tmp_path), "out") + "/"
    os.makedirs(model_dir)
    common = ["--data_dir=./tests/data/", "--model_dir=" + model_dir, "--output_dir=" + out_dir,
              "--setting_file=" + sfile, "--batch_size=16"]
    r = _run(common + ["--max_train_iteration=12", "--steps_per_checkpoint=6"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ndcg_10" in r.stdout
    ckpt = os.path.join(model_dir, "%s.ckpt" % cls)
    assert os.path.isfile(ckpt), os.listdir(model_dir)
    sd = torch.load(ckpt, map_location="cpu")
    assert list(sd.keys())[:4] == ["sequential.layer_norm0.weight", "sequential.layer_norm0.bias",
                                   "sequential.linear0.weight", "sequential.linear0.bias"]
    assert all(torch.isfinite(v).all() for v in sd.values())
    r = _run(common + ["--test_only=True"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Reading model parameters" in r.stdout
    assert any(n.endswith(".ranklist") for n in os.listdir(out_dir)), os.listdir(out_dir)
