#!/bin/bash
# round 2 ncu captures (one GPU): (1) --set full of the NTT passes (all three of a 2^24 transform) + one MSM accumulate + the
# scatter kernels; (2) --set full of the accumulate / reduce / quotient kernels inside one 2^20-gate prove;
# (3) launch list of the bench command (headline prove).  Each ncu run follows a plain run of the same command.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LOG=24 NTT_LOG=24 python scripts/prof_driver.py > gpurun_out/r02_prof_plain.log 2>&1 &&
LOG=24 NTT_LOG=24 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass_tma_kernel|msm_accumulate_kernel|msm_scatter_kernel|msm_fine_scatter_kernel|msm_reduce_chunks_kernel' -c 8 \
    -f -o gpurun_out/prof_r02 python scripts/prof_driver.py > gpurun_out/r02_prof_ncu.log 2>&1
echo "ncu full (driver) rc=$?"; tail -2 gpurun_out/r02_prof_ncu.log
python scripts/prove_bench.py 20 > gpurun_out/r02_prove_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'msm_accumulate_kernel|msm_reduce_chunks_kernel|quotient_kernel' -s 12 -c 6 \
    -f -o gpurun_out/prof_prove_r02 python scripts/prove_bench.py 20 > gpurun_out/r02_prove_ncu.log 2>&1
echo "ncu full (prove) rc=$?"; tail -2 gpurun_out/r02_prove_ncu.log
python bench.py --steps 2 --warmup 3 --skip-large --skip-cpu > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --skip-large --skip-cpu > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
ncu -i gpurun_out/prof_r02.ncu-rep --page raw --csv > gpurun_out/prof_r02_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_prove_r02.ncu-rep --page raw --csv > gpurun_out/prof_prove_r02_raw.csv 2>/dev/null
# per-instruction stall samples of the NTT passes (source page), compressed; the 50 MB reports themselves stay on the box
ncu -i gpurun_out/prof_r02.ncu-rep --page source --csv -k regex:ntt_pass_tma_kernel 2>/dev/null | gzip -9 > gpurun_out/prof_r02_ntt_source.csv.gz
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | grep r02
