#!/bin/bash
mkdir -p gpurun_out
for r in 0.6 0.75 0.88 1.0 1.2 1.5; do echo "== ratio $r"; EDRL_MMD_HYBRID_RATIO=$r timeout 400 python tools/time_shard.py 65536 1024 8; EDRL_MMD_HYBRID_RATIO=$r timeout 400 python tools/time_shard.py 16384 1024 1 noanchor; done
echo "== off"; EDRL_MMD_HYBRID=0 timeout 400 python tools/time_shard.py 65536 1024 8; EDRL_MMD_HYBRID=0 timeout 400 python tools/time_shard.py 16384 1024 1 noanchor
