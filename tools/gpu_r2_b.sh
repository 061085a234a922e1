#!/bin/bash
mkdir -p gpurun_out
echo "== sift tests"; timeout 600 python -m pytest tests/test_gpu_eprl.py -x -q -m gpu -k "topk or sift or select" 2>&1 | tail -4
echo "== timing: sift"; timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_sift.txt | tail -20
