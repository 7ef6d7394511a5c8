#!/bin/bash
# round 2, 8-GPU box: bench at N=4 and N=8 on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for N in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench n$N rc=$?"; grep -v OMP gpurun_out/r02_bench_n$N.err | tail -c 600
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print($N, d['value'], d['e2e']['value'], d['rounds_ms'], d['parity'], d['clocks'])
print(d['msm']['ms'], d['msm']['e2e']['ms'], d['ntt_sharded']['ms'])
k=d['prove_sharded_2e24']; print(k['value'], k['rounds_ms'], k.get('equals_single_gpu_proof'), k.get('single_gpu_ms'))
PY
done
