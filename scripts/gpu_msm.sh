#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_msm_gpu.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_msm.txt
