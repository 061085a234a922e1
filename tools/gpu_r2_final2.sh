#!/bin/bash
# final validation without the ncu captures: GPU suite, smoke(), default bench line + reference arm
mkdir -p gpurun_out
T=${1:-r02d}
echo "== gpu suite"; ( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | tail -10
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; ( time timeout 1500 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err ) 2>&1 | tail -3; tail -c 400 gpurun_out/bench_$T.err
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err ) 2>&1 | tail -3
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_$T.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'], l['roofline']['part_b_select']['frac'], l['roofline']['part_b_select']['sorted']['frac'])
print(json.dumps({k:(v.get('ms_per_step'), v.get('frac')) for k,v in l.get('precision_modes',{}).items() if isinstance(v,dict)}))
print(json.dumps(l.get('scale_anchor')))
print(json.dumps(l.get('roofline_bwd_separate'))[:400])
print(json.dumps(l.get('essence_point'))[:1500])
r=json.loads(open('gpurun_out/bench_ref_$T.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], r['steps_measured'], r['cpu_baseline']['kind'])
PY
