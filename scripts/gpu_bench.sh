#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
