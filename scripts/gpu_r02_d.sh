#!/bin/bash
# round 2: whole GPU suite + NTT sweep + bench on one GPU (after the NTT write-back / prefetch tweak, the reduction chunk model, the sharded evaluation)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.txt 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02d_pytest.txt
timeout 600 python scripts/ntt_sweep.py --configs legacy,tma128,tma168 --out gpurun_out/r02d_ntt_sweep.json > gpurun_out/r02d_ntt_sweep.log 2>&1
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02d_ntt_sweep.json'))
for c in r:
    print(c, [(x['log_n'], round(x['fft_ms'],4)) for x in r[c]] if isinstance(r[c], list) else r[c])
PY
timeout 600 python scripts/msm_tail_scaling.py > gpurun_out/r02d_msm_tail_scaling.json 2>&1; cat gpurun_out/r02d_msm_tail_scaling.json
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r02d_bench_n1.err; head -c 2500 gpurun_out/r02d_bench_n1.json
