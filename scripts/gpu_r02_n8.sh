#!/bin/bash
# round 2, 8-GPU call: bench at N=8 (headline 2^20 sharded prove, strong-scaling MSM / NTT, 2^24-gate prove with single-GPU check)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 \
  > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "bench n8 rc=$?"; grep -v OMP gpurun_out/r02_bench_n8.err | tail -c 1500; head -c 7000 gpurun_out/r02_bench_n8.json
