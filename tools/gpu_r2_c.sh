#!/bin/bash
mkdir -p gpurun_out
echo "== timing: sift"; timeout 300 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk_sift.txt | head -6
echo "== bench ours"; ( time timeout 1200 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err ) 2>&1 | tail -4
tail -c 600 gpurun_out/bench_r02a.err
echo "== bench reference"; ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --ref-budget-s 100 > gpurun_out/bench_ref_r02a.json 2> gpurun_out/bench_ref_r02a.err ) 2>&1 | tail -4
cat gpurun_out/bench_ref_r02a.json | cut -c1-1500
