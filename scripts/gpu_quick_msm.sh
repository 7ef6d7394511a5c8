#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msm_gpu.py -x -q -m gpu 2>&1 | tail -3
SIZES=20,26 python scripts/sweep.py 2>&1 >/dev/null | grep mpts | cut -c1-330
PRECOMP=1 SIZES=20,22 python scripts/sweep.py 2>&1 >/dev/null | grep mpts | cut -c1-330
