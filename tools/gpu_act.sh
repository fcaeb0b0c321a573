mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "other_activations or relu or tanh or sigmoid or mlp_forward_backward or ipw_small" 2>&1 | tail -15 | cut -c1-200
