#!/bin/bash
# final round-2 record on one GPU: whole GPU suite, smoke, the driver's bench command
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_tests.log
tail -4 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/final_bench_n1.json
