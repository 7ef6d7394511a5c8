#!/bin/bash
# Prover-layer GPU tests (KZG + prove + batched MSM), prove timing, and an ncu launch list of one prove
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_prover_gpu.py tests/test_msm_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_prover.txt
timeout 600 python scripts/prove_bench.py 16 18 20 22 2>&1 | tee gpurun_out/prove_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_prove.csv \
    python scripts/prove_bench.py 20 > gpurun_out/prove_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_prove.csv
