#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_mmd.py -q -m gpu -x -k "hybrid" 2>&1 | tail -15
