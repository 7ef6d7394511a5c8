#!/bin/bash
# bench.py as the driver launches it at N = 8, then the sharded-prove check including 2^24 against the single-GPU proof
mkdir -p gpurun_out
NG=8 bash scripts/gpu_bench_n.sh
NG=8 SIZES="24" bash scripts/gpu_dist_prove.sh
