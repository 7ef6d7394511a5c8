#!/bin/bash
# N=2 validation of the sharded MSM path (NCCL all-gather of partial results) + quick regression of the GPU tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_all.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 2 --warmup 3 --log-n 24 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; tail -5 gpurun_out/bench_n2.err; cut -c1-900 gpurun_out/bench_n2.json
