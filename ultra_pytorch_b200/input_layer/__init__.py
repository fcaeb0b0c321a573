"""Drop-in input feeds: `"train_input_feed": "ultra_pytorch_b200.input_layer.ClickSimulationFeed"`."""
from .click_simulation_feed import ClickSimulationFeed  # noqa: F401
from .stochastic_online_simulation_feed import StochasticOnlineSimulationFeed  # noqa: F401  
from .direct_label_feed import DirectLabelFeed  # noqa: F401
