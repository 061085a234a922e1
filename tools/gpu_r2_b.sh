#!/bin/bash
mkdir -p gpurun_out
echo "== sift tests"; timeout 600 python -m pytest tests/test_gpu_eprl.py -x -q -m gpu -k "topk or sift or select" 2>&1 | tail -4
echo "== timing: sift"; timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_sift.txt | head -6
timeout 120 python tools/one_topk.py 262144 800 100 0 > gpurun_out/plain_sift.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_sift_kernel -s 2 -c 1 -f -o /tmp/prof_sift python tools/one_topk.py 262144 800 100 0 > gpurun_out/ncu_sift.log 2>&1
ncu -i /tmp/prof_sift.ncu-rep --page raw --csv > gpurun_out/r02_sift800_raw.csv 2>/dev/null
ncu -i /tmp/prof_sift.ncu-rep --page source --csv > gpurun_out/r02_sift800_source.csv 2>/dev/null
