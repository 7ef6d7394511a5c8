#!/bin/bash
# bench.py exactly as the driver launches it for N = $NG
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
echo "bench rc=$?"; grep -E "Error|error|Traceback" gpurun_out/bench_n$NG.err | head -5; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$NG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","unit","n_gpus","ms_per_step","result_verified")}, d["e2e"]["value"])
print({k:d["ntt_sharded"][k] for k in ("ms","verified","roundtrip_ok","fused_equals_nccl")})
print(d.get("prove_sharded"))
PY
