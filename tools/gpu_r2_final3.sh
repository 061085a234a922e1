#!/bin/bash
# final validation (round 2, last kernels): GPU suite, smoke(), default bench line, reference arm, 3xTF32 capture
mkdir -p gpurun_out
T=${1:-r02f}
echo "== gpu suite"; ( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | tail -10
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; ( time timeout 1500 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err ) 2>&1 | tail -3; tail -c 400 gpurun_out/bench_$T.err
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err ) 2>&1 | tail -3
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_$T.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['e2e']['ms_per_step_runs'], l['roofline']['frac'], l['roofline']['part_b_select']['frac'], l['roofline']['part_b_select']['sorted']['frac'])
print(json.dumps({k:(v.get('ms_per_step'), v.get('frac')) for k,v in l.get('precision_modes',{}).items() if isinstance(v,dict)}))
print(json.dumps(l.get('scale_anchor')))
r=json.loads(open('gpurun_out/bench_ref_$T.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], r['steps_measured'], r['cpu_baseline']['kind'])
PY
timeout 300 python tools/run_sweep.py 3xtf32 8192 512 > gpurun_out/plain_3x.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mmd_sweep256 -s 2 -c 1 -f -o /tmp/prof3x python tools/run_sweep.py 3xtf32 8192 512 > gpurun_out/ncu_3x.log 2>&1
echo "3x capture rc=$?"
ncu -i /tmp/prof3x.ncu-rep --page raw --csv > gpurun_out/${T}_prof_sweep_3xtf32_raw.csv 2>/dev/null
