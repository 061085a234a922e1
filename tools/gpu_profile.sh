#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then the launch list of the same
# command, then --set full captures.  gpurun brings back at most 64 MiB per call, so the captures are split:
#   bash tools/gpu_profile.sh a   launch list + the TF32 training sweep of the bench command + the F16S sweep
#   bash tools/gpu_profile.sh b   the quad sweep (N=8192, d=1024) + the Essence-Point select kernels
# Summaries are copied into profiles/ by tools/summarize_profiles.py (run locally).
mkdir -p gpurun_out
PART=${1:-a}
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
if [ "$PART" = "a" ]; then
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  echo "launch list rc=$?"
  $CMD > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"mmd_sweep256_kernel" -s 3 -c 1 -o gpurun_out/prof_mmd_final -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo "full capture rc=$?"
  python tools/run_sweep.py f16s > gpurun_out/plain_sweep_f16s.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:mmd_sweep256 -s 2 -c 1 -o gpurun_out/prof_sweep_f16s -f python tools/run_sweep.py f16s > gpurun_out/ncu_sweep_f16s.log 2>&1
  echo "f16s capture rc=$?"
else
  python tools/run_sweep.py tf32 8192 1024 > gpurun_out/plain_sweep_quad.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:mmd_sweep_quad -s 2 -c 1 -o gpurun_out/prof_sweep_quad -f python tools/run_sweep.py tf32 8192 1024 > gpurun_out/ncu_sweep_quad.log 2>&1
  echo "quad capture rc=$?"
  python tools/run_topk.py > gpurun_out/plain_topk.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:topk_ -s 2 -c 2 -o gpurun_out/prof_topk_final -f python tools/run_topk.py > gpurun_out/ncu_topk.log 2>&1
  echo "topk capture rc=$?"
fi
