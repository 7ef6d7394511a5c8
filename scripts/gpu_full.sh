#!/bin/bash
# One-GPU validation as the driver runs it at round end: GPU test-suite, smoke(), default bench.py; then the prove size sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_all.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke_plain.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err | cut -c1-400
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","unit","ms_per_step","result_verified","gpu_launches")})
p=d.get("prove"); print({k:p[k] for k in ("value","min_ms","rounds_ms","gpu_launches_per_prove")}, p["cpu_baseline"]["value"], p["cpu_baseline"]["cores"])
PY
timeout 900 python scripts/prove_bench.py 12 14 16 18 20 22 $EXTRA_SIZES 2>&1 | tee gpurun_out/prove_sizes.txt
