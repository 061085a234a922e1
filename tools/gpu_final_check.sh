#!/bin/bash
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (short)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-drivers 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline']['part_b_select']['frac'], l['gpu_launches'], l['clocks'])"
