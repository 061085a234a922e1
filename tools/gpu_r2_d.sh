#!/bin/bash
mkdir -p gpurun_out
echo "== dilr + views tests"; timeout 600 python -m pytest tests/test_gpu_dilr.py tests/test_gpu_views.py -q -m gpu 2>&1 | tail -30
echo "== driver test (DILR swapped too)"; timeout 600 python -m pytest tests/test_gpu_reference_driver.py -x -q -m gpu 2>&1 | tail -8
