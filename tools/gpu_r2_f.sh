#!/bin/bash
mkdir -p gpurun_out
echo "== sharded tests (2 GPUs)"; timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu 2>&1 | tail -8
echo "== bench 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; tail -c 300 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r02_2gpu.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','e2e','parity','strong_scaling_base','cpu_baseline','roofline'):
    print(k, json.dumps(l.get(k))[:500])
PY
