#!/bin/bash
# Prover-layer GPU tests (KZG + prove) and a first prove timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_prover_gpu.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_prover.txt
timeout 600 python scripts/prove_bench.py 16 18 20 2>&1 | tee gpurun_out/prove_bench.txt
