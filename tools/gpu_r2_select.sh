#!/bin/bash
# round 2: the sift select (correctness, timing), then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "== sift tests"; timeout 900 python -m pytest tests/test_gpu_eprl.py -x -q -m gpu -k "topk or sift or select" 2>&1 | tail -8
echo "== timing: sift"; timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_sift.txt | tail -20
echo "== whole GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_eprl.py::test_sift_select_against_stable_sort 2>&1 | tail -25
