#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_ntt_gpu.py -x -q -m gpu 2>&1 | tail -4
for L in 24 26; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 scripts/dist_ntt_check.py $L 2>&1 | grep -E '^\{|Error|error|Traceback' | tee -a gpurun_out/dist_ntt_fused_n$NG.json
done
