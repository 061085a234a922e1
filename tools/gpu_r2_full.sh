#!/bin/bash
# full validation: the GPU suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
echo "== gpu suite"; ( time timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | tail -10
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; ( time timeout 1500 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err ) 2>&1 | tail -3; tail -c 400 gpurun_out/bench_r02b.err
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02b.json 2> gpurun_out/bench_ref_r02b.err ) 2>&1 | tail -3
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r02b.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'], l['roofline']['part_b_select']['frac'], l['roofline']['part_b_select']['sorted']['frac'])
print(json.dumps(l.get('precision_modes'))[:1500])
print(json.dumps(l.get('next_rows'))[:2500])
r=json.loads(open('gpurun_out/bench_ref_r02b.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], r['steps_measured'], r['cpu_baseline']['kind'])
PY
