#!/bin/bash
# round 2, first GPU call: NTT tests on the new TMA path, kernel sweep, whole GPU suite, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r02a_gpu.txt 2>&1
nproc >> gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests/test_ntt_gpu.py -m gpu -x -q > gpurun_out/r02a_pytest_ntt.txt 2>&1
NTT_RC=$?
echo "ntt tests rc=$NTT_RC"; tail -5 gpurun_out/r02a_pytest_ntt.txt
if [ $NTT_RC -ne 0 ]; then
  echo "TMA path failed: rerunning the NTT tests on the legacy kernel"
  PB200_NTT_KERNEL=legacy timeout 900 python -m pytest tests/test_ntt_gpu.py -m gpu -x -q > gpurun_out/r02a_pytest_ntt_legacy.txt 2>&1
  tail -3 gpurun_out/r02a_pytest_ntt_legacy.txt
  export PB200_NTT_KERNEL=legacy
fi
timeout 900 python scripts/ntt_sweep.py --out gpurun_out/r02a_ntt_sweep.json > gpurun_out/r02a_ntt_sweep.log 2>&1
tail -6 gpurun_out/r02a_ntt_sweep.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_ntt_gpu.py > gpurun_out/r02a_pytest_rest.txt 2>&1
echo "rest rc=$?"; tail -8 gpurun_out/r02a_pytest_rest.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02a_bench_n1.err; head -c 3000 gpurun_out/r02a_bench_n1.json
