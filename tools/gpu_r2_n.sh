#!/bin/bash
mkdir -p gpurun_out
for b in 0 1; do echo "== EDRL_TOPK_BALLOT=$b"; EDRL_TOPK_BALLOT=$b timeout 300 python tools/time_topk.py 2>&1 | head -12; done
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_eprl.py -q -m gpu -x 2>&1 | tail -3
