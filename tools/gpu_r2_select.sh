#!/bin/bash
# round 2, first GPU call: the sift select (correctness under both group widths, timing against the round-1 kernels),
# then the whole GPU suite and a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "== sift tests (default)"; timeout 900 python -m pytest tests/test_gpu_eprl.py -x -q -m gpu -k "topk or sift or select" 2>&1 | tail -15
echo "== sift tests (G=32)"; EDRL_TOPK_SIFT_G=32 timeout 900 python -m pytest tests/test_gpu_eprl.py -x -q -m gpu -k "topk or sift" 2>&1 | tail -8
echo "== timing: round-1 kernels"; EDRL_TOPK_SIFT=0 timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_r1kernels.txt | tail -20
echo "== timing: sift, warp per row"; EDRL_TOPK_SIFT_G=32 timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_sift_g32.txt | tail -20
echo "== timing: sift default"; timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_sift.txt | tail -20
