#!/bin/bash
# round-2 final validation: GPU suite, smoke(), default bench line + reference arm, launch list, ncu captures
mkdir -p gpurun_out
T=${1:-r02c}
echo "== gpu suite"; ( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | tail -10
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; ( time timeout 1500 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err ) 2>&1 | tail -3; tail -c 400 gpurun_out/bench_$T.err
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err ) 2>&1 | tail -3
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_$T.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'], l['roofline']['part_b_select']['frac'], l['roofline']['part_b_select']['sorted']['frac'])
print(json.dumps(l.get('precision_modes'))[:1500])
print(json.dumps(l.get('scale_anchor')))
r=json.loads(open('gpurun_out/bench_ref_$T.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], r['steps_measured'], r['cpu_baseline']['kind'])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-drivers"
echo "== launch list"
timeout 300 $CMD > gpurun_out/plain_$T.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launch_$T.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 "$@" > gpurun_out/plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name capture rc=$?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${T}_${name}_raw.csv 2>/dev/null
}
cap prof_sweep_tf32 mmd_sweep256 2 python tools/run_sweep.py tf32 8192 512
cap prof_sweep_quad mmd_sweep_quad 2 python tools/run_sweep.py tf32 8192 1024
cap prof_sweep_3xtf32 mmd_sweep256 2 python tools/run_sweep.py 3xtf32 8192 512
ls -la gpurun_out/*_raw.csv | tail -5
