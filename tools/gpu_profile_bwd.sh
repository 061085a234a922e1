#!/bin/bash
# A/B of the backward variants + one ncu --set full capture of each (same command line, plain run first).
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline"
for v in pair stream legacy; do
  case $v in pair) E="";; stream) E="EDRL_MMD_BWD_STREAM=1";; legacy) E="EDRL_MMD_BWD_LEGACY=1";; esac
  env $E $CMD 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'step ms', round(d['ms_per_step'],3), 'bwd ms', round(d['roofline']['ms'],3), 'fwd ms', round(d['roofline_fwd']['ms'],3))"
done
$CMD > gpurun_out/plain_pair.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"mmd_bwd_pair" -s 3 -c 1 -o gpurun_out/prof_bwd_pair -f $CMD > gpurun_out/ncu_pair.log 2>&1
echo "ncu pair rc=$?"
EDRL_MMD_BWD_STREAM=1 $CMD > gpurun_out/plain_stream.log 2>&1 && EDRL_MMD_BWD_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:"mmd_bwd_pair" -s 3 -c 1 -o gpurun_out/prof_bwd_stream -f $CMD > gpurun_out/ncu_stream.log 2>&1
echo "ncu stream rc=$?"
