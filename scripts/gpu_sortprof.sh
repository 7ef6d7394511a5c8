#!/bin/bash
mkdir -p gpurun_out
LOG=26 NTT_LOG=16 python scripts/prof_driver.py > gpurun_out/sortprof_plain.log 2>&1 &&
LOG=26 NTT_LOG=16 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'msm_count|msm_scatter|msm_fine|msm_coarse' --csv --log-file gpurun_out/sortprof.csv python scripts/prof_driver.py > gpurun_out/sortprof_ncu.log 2>&1
echo rc=$?
