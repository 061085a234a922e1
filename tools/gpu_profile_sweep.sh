#!/bin/bash
# ncu --set full capture of the fused sweep in two precision modes (plain run first, B200_PROFILING.md recipe)
mkdir -p gpurun_out
for prec in tf32 f16s; do
  python tools/run_sweep.py $prec > gpurun_out/plain_sweep_$prec.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:mmd_sweep256 -s 2 -c 1 -o gpurun_out/prof_sweep_$prec -f python tools/run_sweep.py $prec > gpurun_out/ncu_sweep_$prec.log 2>&1
  echo "$prec capture rc=$?"
done
python tools/run_sweep.py f16s > gpurun_out/plain_sweep_ll.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_sweep_f16s.csv python tools/run_sweep.py f16s > /dev/null 2>&1
python tools/run_sweep.py tf32 > gpurun_out/plain_sweep_ll.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_sweep_tf32.csv python tools/run_sweep.py tf32 > /dev/null 2>&1
echo done
