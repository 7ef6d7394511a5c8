#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msm_gpu.py -x -q -m gpu 2>&1 | tail -3
PRECOMP=1 LOG=20 NTT_LOG=20 REPS=3 python scripts/prof_driver.py > gpurun_out/prof_small_plain.log 2>&1 &&
PRECOMP=1 LOG=20 NTT_LOG=20 REPS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_small.csv python scripts/prof_driver.py > gpurun_out/prof_small_ncu.log 2>&1
echo rc=$?
