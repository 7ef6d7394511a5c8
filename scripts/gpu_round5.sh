#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_prover_gpu.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_prover.txt
timeout 600 python scripts/msm_variant.py 2>&1 | grep '^{' | tee gpurun_out/msm_variant.txt
PB200_MSM_ACC_3CTA=1 timeout 600 python scripts/msm_variant.py 2>&1 | grep '^{' | tee -a gpurun_out/msm_variant.txt
timeout 600 python scripts/prove_bench.py 20 2>&1 | tee gpurun_out/prove_bench3.txt
PB200_MSM_ACC_3CTA=1 timeout 600 python scripts/prove_bench.py 20 2>&1 | tee -a gpurun_out/prove_bench3.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:quotient_kernel -c 1 -f -o gpurun_out/quotient_r01 \
    python scripts/prove_bench.py 20 > gpurun_out/quotient_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/quotient_r01.ncu-rep
