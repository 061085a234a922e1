#!/bin/bash
mkdir -p gpurun_out
echo "== head + synthetic step tests"; timeout 900 python -m pytest tests/test_gpu_head.py "tests/test_gpu_eprl.py::test_synthetic_training_step_swaps_in_with_step_level_parity" -q -m gpu 2>&1 | tail -30
echo "== synthetic step b64"; timeout 600 python examples/edrl_step_synthetic.py --batch 64 --steps 10 2>&1 | tail -3
