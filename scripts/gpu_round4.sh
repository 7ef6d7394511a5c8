#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_msm_gpu.py tests/test_prover_gpu.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_msm_prover.txt
timeout 600 python scripts/prove_bench.py 16 20 2>&1 | tee gpurun_out/prove_bench2.txt
timeout 900 python bench.py --steps 3 > gpurun_out/bench_r4.json 2> gpurun_out/bench_r4.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_r4.err | cut -c1-400
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r4.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","unit","ms_per_step","result_verified","gpu_launches")}, d["roofline"]["phases_ms"])
p=d.get("prove"); print({k:p[k] for k in ("value","min_ms","rounds_ms","gpu_launches_per_prove")})
PY
