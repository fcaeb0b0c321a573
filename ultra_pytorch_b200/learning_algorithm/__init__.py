"""Drop-in learning algorithms: put e.g. `"learning_algorithm": "ultra_pytorch_b200.learning_algorithm.IPWrank"`
in the settings JSON (resolved by ultra.utils.find_class, ultra/utils/sys_tools.py:7-21)."""
from .base_algorithm import B200Algorithm  # noqa: F401
from .navie_algorithm import NavieAlgorithm  # noqa: F401
from .ipw_rank import IPWrank  # noqa: F401
from .dla import DLA, DenoisingNet  # noqa: F401
from .pairwise_debias import PairDebias  # noqa: F401
from .lambda_rank import LambdaRank  # noqa: F401
from .prs_rank import PRSrank  # noqa: F401
from .regression_EM import RegressionEM  # noqa: F401
