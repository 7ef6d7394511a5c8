#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_prover_gpu.py tests/test_ntt_gpu.py -m gpu -x -q > gpurun_out/r02g_pytest.txt 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02g_pytest.txt
for mode in overlap nooverlap; do
  if [ $mode = nooverlap ]; then export PB200_PROVE_NO_OVERLAP=1; else unset PB200_PROVE_NO_OVERLAP; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --skip-large --skip-cpu > gpurun_out/r02g_bench_$mode.json 2> gpurun_out/r02g_bench_$mode.err
  echo "bench $mode rc=$?"; tail -c 300 gpurun_out/r02g_bench_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode', d['value'], d['e2e']['value'], d['rounds_ms'], d['parity'])
PY
done
unset PB200_PROVE_NO_OVERLAP
timeout 300 python scripts/prove_bench.py 12 16 18 20 22 > gpurun_out/r02g_prove_sizes.txt 2>&1; cat gpurun_out/r02g_prove_sizes.txt | cut -c1-400
