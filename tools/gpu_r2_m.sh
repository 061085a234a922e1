#!/bin/bash
mkdir -p gpurun_out
for o in 0 1; do for h in 0 1; do echo "== order $o hybrid $h"; EDRL_MMD_QUAD_ORDER=$o EDRL_MMD_HYBRID=$h timeout 300 python tools/time_shard.py 65536 1024 8; EDRL_MMD_QUAD_ORDER=$o EDRL_MMD_HYBRID=$h timeout 300 python tools/time_shard.py 32768 1024 1 noanchor; done; done
