"""Data-parallel launcher for the reference's unmodified main.py (one replica per GPU of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/launch_dp.py /path/to/ULTRA_pytorch/main.py --setting_file=... --data_dir=... --model_dir=... [...]

main.py itself needs no change: the B200 learning algorithm named in the settings JSON joins the NCCL group when it is
constructed (ultra_pytorch_b200/learning_algorithm/base_algorithm.py: _bootstrap_data_parallel) and rank 0 alone writes
the checkpoint.  This wrapper only makes the run tidy: it seeds the feed's random generators differently per rank (so
the ranks sample different queries even when the caller seeds them), silences the duplicate prints of ranks > 0 and runs
main.py with the reference directory as the working directory (its relative ./example paths)."""
import os
import random
import runpy
import sys

import numpy as np

if __name__ == "__main__":
    main_py = os.path.abspath(sys.argv[1])
    rank = int(os.environ.get("RANK", "0"))
    random.seed(1000003 * rank + int(os.environ.get("UB200_SEED", "0")))
    np.random.seed(7919 * rank + int(os.environ.get("UB200_SEED", "0")))
    if rank != 0 and os.environ.get("UB200_DP_VERBOSE", "0") != "1":
        sys.stdout = open(os.devnull, "w")
    os.chdir(os.path.dirname(main_py))
    sys.path.insert(0, os.path.dirname(main_py))
    sys.argv = [main_py] + sys.argv[2:]
    runpy.run_path(main_py, run_name="__main__")
