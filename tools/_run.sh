mkdir -p gpurun_out/r02y
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -q -x --durations=6 2>&1 | tail -14
