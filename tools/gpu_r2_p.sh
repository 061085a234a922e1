#!/bin/bash
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
EDRL_MMD_FUSED=0 timeout 600 python -m pytest tests/test_gpu_mmd.py -q -m gpu -k "kat or edge or variants or ragged or midsize or full_size_vs or beyond" 2>&1 | tail -3
timeout 600 python bench.py --no-extras --no-drivers --no-cpu-baseline --steps 10 --warmup 3 | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['frac'], json.dumps(l['roofline'].get('part_a',{}))[:600])
print(json.dumps({k:v for k,v in l.items() if 'separate' in k or 'bwd' in k})[:1200])
"
