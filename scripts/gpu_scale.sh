#!/bin/bash
# Multi-GPU scaling: bench.py (MSM weak scaling + sharded NTT, NCCL vs fused peer stores) and the sharded NTT check, N = $NG
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
echo "bench rc=$?"; grep -E "Error|error|Traceback" gpurun_out/bench_n$NG.err | head -5; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$NG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","unit","n_gpus","ms_per_step","result_verified")}, d["e2e"]["value"], d.get("ntt_sharded"))
PY
rm -f gpurun_out/dist_ntt_n$NG.json
for L in 24 26; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29552 scripts/dist_ntt_check.py $L 2>&1 | grep -E '^\{|Error|Traceback' | tee -a gpurun_out/dist_ntt_n$NG.json
done
