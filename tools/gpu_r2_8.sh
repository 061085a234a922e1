#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; tail -c 300 gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','parity','strong_scaling_base'):
    print(k, json.dumps(l.get(k))[:400])
print('e2e', l['e2e']['ms_per_step'])
PY
echo "== sharded tests on all GPUs"; timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu -k "quad or nccl" 2>&1 | tail -4
