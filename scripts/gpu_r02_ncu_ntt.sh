#!/bin/bash
# ncu --set full of the three passes of a 2^24 forward NTT with the final round-2 kernel (after a plain run of the same command)
mkdir -p gpurun_out
LOG=16 NTT_LOG=24 python scripts/prof_driver.py > gpurun_out/ncu_ntt_plain.log 2>&1 || exit 1
LOG=16 NTT_LOG=24 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass_tma_kernel' -c 3 \
    -f -o gpurun_out/prof_ntt_final python scripts/prof_driver.py > gpurun_out/ncu_ntt_run.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ntt_run.log
ncu -i gpurun_out/prof_ntt_final.ncu-rep --page raw --csv > gpurun_out/prof_ntt_final_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | grep ntt_final
