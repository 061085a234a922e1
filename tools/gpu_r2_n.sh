#!/bin/bash
mkdir -p gpurun_out
echo "== eprl tests"; timeout 900 python -m pytest tests/test_gpu_eprl.py -q -m gpu -x 2>&1 | tail -3
echo "== timing"; timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_occ.txt | head -18
timeout 120 python tools/one_topk.py 262144 800 100 0 > gpurun_out/plain_sift.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_sift_kernel -s 2 -c 1 -f -o /tmp/prof_sift python tools/one_topk.py 262144 800 100 0 > gpurun_out/ncu_sift.log 2>&1
echo "capture rc=$?"
ncu -i /tmp/prof_sift.ncu-rep --page raw --csv > gpurun_out/r02d_sift800_raw.csv 2>/dev/null
timeout 120 python tools/one_topk.py 262144 800 100 1 > gpurun_out/plain_sift_s.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:topk_sift_kernel -s 2 -c 1 -f -o /tmp/prof_sift_s python tools/one_topk.py 262144 800 100 1 > gpurun_out/ncu_sift_s.log 2>&1
echo "capture sorted rc=$?"
ncu -i /tmp/prof_sift_s.ncu-rep --page raw --csv > gpurun_out/r02d_sift800_sorted_raw.csv 2>/dev/null
ls -la gpurun_out/r02d_*
