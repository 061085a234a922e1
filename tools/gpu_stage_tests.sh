#!/bin/bash
# Run the GPU parity tests in separate processes (a faulting kernel poisons its CUDA context, so
# every stage gets its own) with a hard timeout per stage.  Output: gpurun_out/stage_*.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name timeout pytest-args...
  local name=$1 to=$2; shift 2
  timeout "$to" python -m pytest -m gpu -q -x --timeout=120 "$@" > "gpurun_out/stage_$name.log" 2>&1
  echo "stage $name exit=$? : $(tail -1 gpurun_out/stage_$name.log)"
}
run eprl      300 tests/test_gpu_eprl.py
run gram      180 tests/test_gpu_mmd.py -k centred_gram
run kmat      180 tests/test_gpu_mmd.py -k gaussian_kernel_matrix
run kat       300 tests/test_gpu_mmd.py -k kat_forward_backward
run variants  300 tests/test_gpu_mmd.py -k "variants or edge or errors"
run shards    180 tests/test_gpu_mmd.py -k row_ranges
run big       400 tests/test_gpu_mmd.py -k "full_size or midsize"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/stage_smoke.log 2>&1; echo "smoke exit=$?"
