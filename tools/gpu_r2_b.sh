#!/bin/bash
mkdir -p gpurun_out
echo "== timing: sift"; timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_sift.txt | tail -20
timeout 120 python tools/one_topk.py 262144 800 100 0 > gpurun_out/plain_sift.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_sift_kernel -s 2 -c 1 -f -o /tmp/prof_sift python tools/one_topk.py 262144 800 100 0 > gpurun_out/ncu_sift.log 2>&1
ls -la /tmp/prof_sift.ncu-rep
ncu -i /tmp/prof_sift.ncu-rep --page raw --csv > gpurun_out/r02_sift800_raw.csv 2>/dev/null
ncu -i /tmp/prof_sift.ncu-rep --page source --csv > gpurun_out/r02_sift800_source.csv 2>/dev/null
[ $(stat -c %s /tmp/prof_sift.ncu-rep) -lt 30000000 ] && cp /tmp/prof_sift.ncu-rep gpurun_out/r02_prof_sift800.ncu-rep
echo "== reference driver + sharded tests"; timeout 600 python -m pytest tests/test_gpu_reference_driver.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -25
