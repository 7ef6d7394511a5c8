#!/bin/bash
# round 2, 2-GPU call: multi-process tests, widget tests, bench at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02b_gpus.txt
timeout 900 python -m pytest tests/test_prover_gpu.py -m gpu -x -q -k "ecc or dev or duplicate" > gpurun_out/r02b_pytest_widgets.txt 2>&1
echo "widgets rc=$?"; tail -5 gpurun_out/r02b_pytest_widgets.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02b_pytest_multi.txt 2>&1
echo "multi rc=$?"; tail -12 gpurun_out/r02b_pytest_multi.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
  > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err
echo "bench n2 rc=$?"; tail -c 2500 gpurun_out/r02b_bench_n2.err; head -c 6000 gpurun_out/r02b_bench_n2.json
