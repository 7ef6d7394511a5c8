#!/bin/bash
# Full GPU regression + bench + sweep (round-end dress rehearsal)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_all.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "ref rc=$?"
PRECOMP=1 python scripts/sweep.py > gpurun_out/sweep_pre.json 2> gpurun_out/sweep_pre.err; echo "sweep rc=$?"
python scripts/sweep.py > gpurun_out/sweep.json 2> gpurun_out/sweep.err; echo "sweep rc=$?"
