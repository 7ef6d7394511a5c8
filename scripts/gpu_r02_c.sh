#!/bin/bash
# round 2, 2-GPU call: the in-library communicator (tests), bench at N=2 with lib collectives vs torch callbacks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02c_pytest_multi.txt 2>&1
echo "multi rc=$?"; tail -15 gpurun_out/r02c_pytest_multi.txt
for coll in lib torch; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 \
  --skip-large --collectives $coll > gpurun_out/r02c_bench_n2_$coll.json 2> gpurun_out/r02c_bench_n2_$coll.err
echo "bench n2 $coll rc=$?"; tail -c 600 gpurun_out/r02c_bench_n2_$coll.err | grep -v OMP; python - <<PY
import json
d=json.loads(open('gpurun_out/r02c_bench_n2_$coll.json').read().strip().splitlines()[-1])
print('$coll', d['value'], d['e2e']['value'], d['rounds_ms'], d['parity'], d['roofline']['phases_ms_per_step'])
PY
done
