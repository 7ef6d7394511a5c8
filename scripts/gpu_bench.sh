#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msm_gpu.py -x -q -m gpu -k g1_sum 2>&1 | tail -3
timeout 900 python bench.py --log-n 22 --ntt-log-n 22 --steps 3 --warmup 3 --cpu-sample-log 18 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"; tail -5 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
