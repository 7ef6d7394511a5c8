#!/bin/bash
# round 2, 2-GPU call: multi-process tests + bench at N=2 on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02f_pytest_multi.txt 2>&1
echo "multi rc=$?"; tail -8 gpurun_out/r02f_pytest_multi.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 \
  > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
echo "bench n2 rc=$?"; grep -v OMP gpurun_out/r02f_bench_n2.err | tail -c 800
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['rounds_ms'], d['parity'], d['clocks'])
print(d['msm']['ms'], d['ntt_sharded']['ms'], d['prove_sharded_2e24']['value'], d['prove_sharded_2e24'].get('equals_single_gpu_proof'))
PY
