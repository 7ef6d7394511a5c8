#!/bin/bash
# sharded prove check at N = $NG GPUs
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29561 \
    scripts/dist_prove_check.py $SIZES 2>&1 | grep -E '^\{|Error|error|Traceback|assert' | tee gpurun_out/dist_prove_n$NG.json
