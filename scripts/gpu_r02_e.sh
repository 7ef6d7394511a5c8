#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.txt 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02e_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02e_smoke.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r02e_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['rounds_ms'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'], d['parity'])
PY
