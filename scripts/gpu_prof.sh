#!/bin/bash
# ncu captures: (1) full-set profile of the dominant kernels at the bench sizes, (2) launch list of the bench command.
mkdir -p gpurun_out
LOG=26 NTT_LOG=24 python scripts/prof_driver.py > gpurun_out/prof_plain.log 2>&1 &&
LOG=26 NTT_LOG=24 ncu --set full --clock-control none --import-source on -k regex:'msm_accumulate_kernel|ntt_pass_kernel' -c 4 \
    -f -o gpurun_out/prof_r01 python scripts/prof_driver.py > gpurun_out/prof_ncu.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/prof_ncu.log
python bench.py --steps 2 --warmup 3 --cpu-sample-log 16 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --cpu-sample-log 16 > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
ncu -i gpurun_out/prof_r01.ncu-rep --page raw --csv > gpurun_out/prof_r01_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
