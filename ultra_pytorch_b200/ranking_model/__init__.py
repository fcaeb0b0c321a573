"""Drop-in ranking models: `"ranking_model": "ultra_pytorch_b200.ranking_model.DNN"` in the settings JSON."""
from .DNN import DNN  # noqa: F401
from .Linear import Linear  # noqa: F401
