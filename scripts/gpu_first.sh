#!/bin/bash
# First GPU contact: NTT parity tests + integer-pipe microbenchmark.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep -E 'Model name|Socket|Core|Thread' >> gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
python - <<'PY' > gpurun_out/imad.txt 2>&1
import plonk_prototype_b200 as pb
c = pb.Context(0)
for i in range(3):
    ops, mhz = c.imad_peak()
    print("IMAD.WIDE lane-ops/s = %.4e  (per SM per clk at 1965MHz: %.2f)  cta0 cycles-derived MHz %.1f" % (ops, ops/148/1.965e9, mhz))
PY
cat gpurun_out/imad.txt
timeout 900 python -m pytest tests/test_ntt_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_ntt.txt
