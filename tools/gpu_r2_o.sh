#!/bin/bash
EDRL_MMD_FUSED=0 timeout 600 python -m pytest tests/test_gpu_mmd.py -q -m gpu -k "kat or edge or variants or ragged or midsize or full_size_vs" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_mmd.py -q -m gpu -k "autograd_thread or hybrid or row_ranges or fused_pass" 2>&1 | tail -5
